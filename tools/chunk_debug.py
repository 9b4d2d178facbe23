"""Which kernel's per-particle result depends on the particles it shares a launch with?  Runs residual / coefficient /
cotangent / update on n particles and on slices of them, and reports the first stage whose outputs differ bit-wise."""
import os
import sys

import numpy as np
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from dps_ttc_b200 import kernels, tables  # noqa: E402
from dps_ttc_b200.kernels import OperatorPlan  # noqa: E402
from dps_ttc_b200.schedule import Schedule, named_beta_schedule  # noqa: E402

dev = torch.device("cuda:0")
k = Schedule(named_beta_schedule("linear", 1000)).consts(999)
g = torch.Generator(dev).manual_seed(11)
rnd = lambda *s: torch.randn(*s, device=dev, generator=g)  # noqa: E731
n = 6
x = rnd(n, 3, 256, 256)
o6 = rnd(n, 6, 256, 256) * 0.3
z = rnd(n, 3, 256, 256)
vjp = rnd(n, 3, 256, 256) * 1e-3
(fh, wh), (fw, ww), _ = tables.resizer_tables((1, 3, 256, 256), 0.25)
plans = {"sr4": OperatorPlan.resize(fh, wh, fw, ww, 3, 256, 256, dev), "phase": OperatorPlan.phase(64, 3, 256, 256, dev),
         "gauss": OperatorPlan.blur(tables.gaussian_kernel(61, 3.0).astype(np.float32), 3, 256, 256, dev)}
np.random.seed(8)
plans["motion"] = OperatorPlan.blur(tables.motion_kernel(61, 0.5).astype(np.float32), 3, 256, 256, dev)
plans["inpaint"] = OperatorPlan.inpainting((np.random.rand(256, 256) > 0.5).astype(np.float32), 3, 256, 256, dev)


def step(plan, sl, mode):
    xs, os_, zs, vs = x[sl].contiguous(), o6[sl], z[sl], vjp[sl]
    y = torch.rand((1,) + tuple(plan.out_shape), device=dev, generator=torch.Generator(dev).manual_seed(3))
    r, partials, aux = plan.forward(xs, os_[:, :3], k, True, y, want_partials=True)
    dist, coef = kernels.guidance_coef(partials, mode, 0.3)
    g6 = torch.zeros(xs.shape[0], 6, 256, 256, device=dev)
    plan.adjoint(r, coef, xs, os_[:, :3], k, True, None, out=g6[:, :3], aux=aux)
    xn, _, _ = kernels.posterior_update("ddpm", xs, os_[:, :3], os_[:, 3:], zs, k, g=g6[:, :3], vjp=vs)
    return {"r": r, "partials": partials, "dist": dist, "coef": coef, "g": g6[:, :3].contiguous(), "x_next": xn}


for name, plan in plans.items():
    for mode in (1, 2):
        full = step(plan, slice(0, n), mode)
        for c in (1, 2, 3):
            bad = []
            for a in range(0, n, c):
                part = step(plan, slice(a, a + c), mode)
                for key, v in part.items():
                    if not torch.equal(v, full[key][a:a + c]):
                        bad.append((key, a, float((v - full[key][a:a + c]).abs().max())))
            print(f"{name:8s} coef_mode={mode} chunk={c}: " + ("bit-identical" if not bad else f"DIFFERS {bad[:6]}"), flush=True)
