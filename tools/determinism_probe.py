"""Run-to-run determinism and uninitialised-read probe of the operator kernels (and of the stand-in / real ε-model):
every forward / adjoint / update launch is repeated on the SAME inputs with the allocator's free blocks poisoned with NaN
in between; any bit that changes, or any NaN in an output, is reported."""
import os
import sys

import numpy as np
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, "tests"))
from dps_ttc_b200 import kernels, tables  # noqa: E402
from dps_ttc_b200.kernels import OperatorPlan  # noqa: E402
from dps_ttc_b200.schedule import Schedule, named_beta_schedule  # noqa: E402

dev = torch.device("cuda:0")
k = Schedule(named_beta_schedule("linear", 1000)).consts(999)


def poison():
    for mb in (1, 8, 64, 256):
        t = torch.full((mb * 262144,), float("nan"), device=dev)
        del t


def plans_for(size):
    out = {}
    if size % 4 == 0:
        (fh, wh), (fw, ww), _ = tables.resizer_tables((1, 3, size, size), 0.25)
        out["sr4"] = OperatorPlan.resize(fh, wh, fw, ww, 3, size, size, dev)
    out["gauss"] = OperatorPlan.blur(tables.gaussian_kernel(61, 3.0).astype(np.float32), 3, size, size, dev)
    if size >= 64:
        np.random.seed(8)
        out["motion"] = OperatorPlan.blur(tables.motion_kernel(61, 0.5).astype(np.float32), 3, size, size, dev)
    out["inpaint"] = OperatorPlan.inpainting((np.random.rand(size, size) > 0.5).astype(np.float32), 3, size, size, dev)
    if size == 256:
        out["phase"] = OperatorPlan.phase(64, 3, 256, 256, dev)
    return out


def run(plan, x, o6, y, with_src):
    poison()
    if with_src:
        r, partials, aux = plan.forward(x, o6[:, :3], k, True, y, want_partials=True)
    else:
        r, partials, aux = plan.forward(x, None, None, False, y, want_partials=True)
    poison()
    norm = kernels.particle_norms(partials)
    coef = torch.where(norm > 0, -0.3 / norm, torch.zeros_like(norm)).contiguous()
    if with_src:
        g = plan.adjoint(r, coef, x, o6[:, :3], k, True, None, aux=aux)
    else:
        g = plan.adjoint(r, coef=coef, aux=aux)
    return {"r": r.clone(), "partials": partials.clone(), "g": g.clone()}


bad = 0
for size in (32, 64, 256):
    for n in (1, 2, 3, 8):
        gen = torch.Generator(dev).manual_seed(size * 100 + n)
        x = torch.randn(n, 3, size, size, device=dev, generator=gen)
        o6 = torch.randn(n, 6, size, size, device=dev, generator=gen) * 0.3
        for name, plan in plans_for(size).items():
            y = torch.rand((1,) + tuple(plan.out_shape), device=dev, generator=gen)
            for with_src in (True, False):
                first = run(plan, x, o6, y, with_src)
                msgs = []
                for rep in range(6):
                    again = run(plan, x, o6, y, with_src)
                    for key in first:
                        if not torch.isfinite(again[key]).all():
                            msgs.append(f"{key}: non-finite")
                        elif not torch.equal(first[key], again[key]):
                            msgs.append(f"{key}: differs by {float((first[key] - again[key]).abs().max()):.2e}")
                if msgs:
                    bad += 1
                    print(f"NONDETERMINISTIC {name} {size}x{size} n={n} src={with_src}: {sorted(set(msgs))}", flush=True)
print(f"operator kernels: {bad} non-deterministic / poisoned configurations", flush=True)

# ---- ε-models: repeat forward + input-VJP on the same input ----
from helpers import TinyEps  # noqa: E402
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False


def model_probe(model, x, tag):
    outs = []
    for rep in range(4):
        xi = x.clone().requires_grad_(True)
        out = model(xi, torch.full((1,), 999.0, device=dev))
        cot = torch.ones_like(out) * 1e-3
        (gx,) = torch.autograd.grad(out, xi, cot)
        outs.append((out.detach().clone(), gx.clone()))
    same = all(torch.equal(outs[0][0], o[0]) and torch.equal(outs[0][1], o[1]) for o in outs[1:])
    print(f"{tag}: forward+VJP repeatable bit-for-bit: {same}", flush=True)


for det in (False, True):
    torch.backends.cudnn.deterministic = det
    gen = torch.Generator(dev).manual_seed(1)
    model_probe(TinyEps(seed=3).to(dev), torch.randn(2, 3, 256, 256, device=dev, generator=gen), f"TinyEps cudnn.deterministic={det}")
    from dps_ttc_b200 import _ref
    if _ref.reference_root() is not None:
        unet = _ref.create_unet("model_config.yaml", reinit_zero_seed=0, device=dev)
        for tf32 in (False, True):
            torch.backends.cudnn.allow_tf32 = tf32
            model_probe(unet, torch.randn(2, 3, 256, 256, device=dev, generator=gen), f"FFHQ UNet cudnn.deterministic={det} tf32={tf32}")
        torch.backends.cudnn.allow_tf32 = False
        del unet
