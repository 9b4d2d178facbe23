"""Race check of the fused cluster guidance kernels (SR ×4 / ×8, separable blur): many planes per launch, repeated launches
must be bit-identical and agree with the two-kernel path (SR: bit for bit; blur: parity tolerance).
    python tools/fused_stress.py [n] [reps]"""
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


def plans():
    for f in (4, 8):
        (fh, wh), (fw, ww), _ = tables.resizer_tables((1, 3, 256, 256), 1.0 / f)
        yield f"sr{f}", OperatorPlan.resize(fh, wh, fw, ww, 3, 256, 256, dev), (1, 3, 256 // f, 256 // f), 1e-6
    yield "gauss61", OperatorPlan.blur(tables.gaussian_kernel(61, 3.0).astype(np.float32), 3, 256, 256, dev), (1, 3, 256, 256), 5e-6
    yield "gauss9", OperatorPlan.blur(tables.gaussian_kernel(9, 1.0).astype(np.float32), 3, 256, 256, dev), (1, 3, 256, 256), 5e-6


ok = True
for name, plan, yshape, tol in plans():
    gen = torch.Generator(dev).manual_seed(5)
    x = torch.randn(n, 3, 256, 256, device=dev, generator=gen) / k.c1
    o6 = torch.randn(n, 6, 256, 256, device=dev, generator=gen) * 0.3 / k.c2
    y = torch.randn(*yshape, device=dev, generator=gen)
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
    print(f"{name}: n={n} fused vs two kernels rel err {err:.2e} (tol {tol:.0e}); {reps} repeats bit-identical: {same}")
    ok = ok and same and err <= tol
print("PASS" if ok else "FAIL")
sys.exit(0 if ok else 1)
