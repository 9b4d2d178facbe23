"""Phase timeline of the fused SR guidance kernel (experiment build with -DDPS_RSF_TRACE):
    tools/build_variant.sh rsftrace resize_fused.cu -DDPS_RSF_TRACE
    DPSTTC_LIB=dps_ttc_b200/build_variants/libdpsttc_rsftrace.so python tools/rsf_trace.py [n]"""
import ctypes
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dps_ttc_b200 import _lib, tables  # noqa: E402
from dps_ttc_b200.kernels import OperatorPlan  # noqa: E402
from dps_ttc_b200.schedule import Schedule, named_beta_schedule  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 8
dev = torch.device("cuda:0")
k = Schedule(named_beta_schedule("linear", 1000)).consts(500)
(fh, wh), (fw, ww), _ = tables.resizer_tables((1, 3, 256, 256), 0.25)
plan = OperatorPlan.resize(fh, wh, fw, ww, 3, 256, 256, dev)
S = 24   # rotate over more data than L2 holds
x = [torch.randn(n, 3, 256, 256, device=dev) / k.c1 for _ in range(S)]
o6 = [torch.randn(n, 6, 256, 256, device=dev) * 0.3 / k.c2 for _ in range(S)]
g6 = torch.zeros(n, 6, 256, 256, device=dev)
y = torch.randn(1, 3, 64, 64, device=dev)
for i in range(S):
    plan.guidance(x[i], o6[i][:, :3], k, True, y, out=g6[:, :3])
torch.cuda.synchronize()
buf = np.zeros((4096, 16), dtype=np.int64)
rc = _lib.lib().dps_debug_rsf_trace(buf.ctypes.data_as(ctypes.c_void_p))
assert rc == 0, rc
nb = min(4096, 3 * n * 8)
t = buf[:nb, :12].astype(np.float64)
d = np.diff(t, axis=1)
names = ["mbar init+TMA issue+tables+y", "pass 0 (TMA wait, x0)", "syncthreads+arrive#1+H own rows", "wait #1", "H halo rows+tile stores+sync",
         "W pass+residual", "sync+arrive#2+sums+own u rows", "wait #2", "remote u rows+arrive#3", "A_h^T + stores", "wait #3"]
print(f"CTAs traced {nb}; lifetime mean {np.mean(t[:, 11] - t[:, 0]):.0f} cycles; first start to last end {t[:, 11].max() - t[:, 0].min():.0f} (clocks of different SMs are not aligned)")
for i, nm in enumerate(names):
    print(f"{nm:34s} mean {d[:, i].mean():8.0f}  median {np.median(d[:, i]):8.0f}  p90 {np.percentile(d[:, i], 90):8.0f}")
