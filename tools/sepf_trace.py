"""Phase timeline of the fused separable-blur guidance kernel (experiment build with -DDPS_SEPF_TRACE):
    tools/build_variant.sh trace blur_fused.cu -DDPS_SEPF_TRACE
    DPSTTC_LIB=dps_ttc_b200/build_variants/libdpsttc_trace.so python tools/sepf_trace.py [n]"""
import ctypes
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dps_ttc_b200 import _lib, tables  # noqa: E402
from dps_ttc_b200.kernels import OperatorPlan  # noqa: E402
from dps_ttc_b200.schedule import Schedule, named_beta_schedule  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 128
dev = torch.device("cuda:0")
k = Schedule(named_beta_schedule("linear", 1000)).consts(500)
plan = OperatorPlan.blur(tables.gaussian_kernel(61, 3.0).astype(np.float32), 3, 256, 256, dev)
x = torch.randn(n, 3, 256, 256, device=dev) / k.c1
o6 = torch.randn(n, 6, 256, 256, device=dev) * 0.3 / k.c2
g6 = torch.zeros(n, 6, 256, 256, device=dev)
y = torch.randn(1, 3, 256, 256, device=dev)
for _ in range(3):
    plan.guidance(x, o6[:, :3], k, True, y, out=g6[:, :3])
torch.cuda.synchronize()
buf = np.zeros((4096, 16), dtype=np.int64)
lib = _lib.lib()
rc = lib.dps_debug_sepf_trace(buf.ctypes.data_as(ctypes.c_void_p))
assert rc == 0, rc
nb = min(4096, 3 * n * 8)
t = buf[:nb, :13].astype(np.float64)
sm = buf[:nb, 15]
d = np.diff(t, axis=1)
names = ["init+TMA+pass0", "arrive#1 + V own rows", "wait #1", "V halo rows+T2 stores", "syncthreads", "H pass FMAs", "y+residual+Z2+sync",
         "Hᵀ FMAs", "wait#2+s stores+fold", "arrive#3+Vᵀ own rows", "wait #3", "Vᵀ halo rows+stores+wait#4"]
print(f"CTAs traced {nb}; lifetime mean {np.mean(t[:, 12] - t[:, 0]):.0f} cycles, median {np.median(t[:, 12] - t[:, 0]):.0f}")
for i, nm in enumerate(names):
    print(f"{nm:24s} mean {d[:, i].mean():8.0f}  median {np.median(d[:, i]):8.0f}  p90 {np.percentile(d[:, i], 90):8.0f}")
# per-SM: span between first start and last end, sum of lifetimes
for s_ in np.unique(sm)[:3]:
    m = sm == s_
    print(f"SM {s_}: {m.sum()} CTAs, span {t[m, 12].max() - t[m, 0].min():.0f} cycles, sum of lifetimes {np.sum(t[m, 12] - t[m, 0]):.0f}")
