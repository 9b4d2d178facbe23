"""Race check of the fused separable-blur guidance kernel: many planes per launch, repeated launches must be bit-identical
and within the parity tolerance of the two-kernel path.    python tools/sepf_stress.py [n] [reps]"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dps_ttc_b200 import tables  # noqa: E402
from dps_ttc_b200.kernels import OperatorPlan  # noqa: E402
from dps_ttc_b200.schedule import Schedule, named_beta_schedule  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 96
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 30
dev = torch.device("cuda:0")
k = Schedule(named_beta_schedule("linear", 1000)).consts(500)
ok = True
for name, kern in (("gauss61", tables.gaussian_kernel(61, 3.0)), ("gauss9", tables.gaussian_kernel(9, 1.0))):
    plan = OperatorPlan.blur(kern.astype(np.float32), 3, 256, 256, dev)
    gen = torch.Generator(dev).manual_seed(5)
    x = torch.randn(n, 3, 256, 256, device=dev, generator=gen) / k.c1
    o6 = torch.randn(n, 6, 256, 256, device=dev, generator=gen) * 0.3 / k.c2
    y = torch.randn(1, 3, 256, 256, device=dev, generator=gen)
    r2, p2, _ = plan.forward(x, o6[:, :3], k, True, y, want_partials=True)
    g2 = torch.zeros(n, 3, 256, 256, device=dev)
    plan.adjoint(r2, None, x, o6[:, :3], k, True, None, out=g2)
    g0 = torch.zeros(n, 3, 256, 256, device=dev)
    p0, _, _ = plan.guidance(x, o6[:, :3], k, True, y, out=g0)
    err = float((g0 - g2).abs().max()) / max(1.0, float(g2.abs().max()))
    same = True
    for _ in range(reps):
        g1 = torch.full((n, 3, 256, 256), float("nan"), device=dev)
        p1, _, _ = plan.guidance(x, o6[:, :3], k, True, y, out=g1)
        same = same and torch.equal(g1, g0) and torch.equal(p1, p0)
    print(f"{name}: n={n} fused vs two kernels rel err {err:.2e}; {reps} repeats bit-identical: {same}")
    ok = ok and same and err <= 5e-6
print("PASS" if ok else "FAIL")
sys.exit(0 if ok else 1)
