"""Fused project / ortho_project (dps_operator_project) against the reference's composition on the GPU: equality over repeated
launches (the cluster kernel's DSMEM hand-over under load) and time per call.
    python tools/project_bench.py [--n 8]"""
import argparse
import json
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from dps_ttc_b200.registry import get_operator  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=8)
ap.add_argument("--iters", type=int, default=50)
args = ap.parse_args()
dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(5)


def timeit(fn, iters):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) * 1e3 / iters


ok = True
for name, cfg in (("super_resolution", dict(in_shape=(1, 3, 256, 256), scale_factor=4)),
                  ("super_resolution", dict(in_shape=(1, 3, 256, 256), scale_factor=8)),
                  ("gaussian_blur", dict(kernel_size=61, intensity=3.0)), ("motion_blur", dict(kernel_size=61, intensity=0.5))):
    np.random.seed(2)
    op = get_operator(name, device=dev, **cfg)
    x = (torch.randn((args.n, 3, 256, 256), generator=g) * 1.2).to(dev)
    m = tuple(op.forward(x[:1]).shape[1:])
    y = torch.randn((1,) + m, generator=g).to(dev)
    if name == "super_resolution":
        up = lambda u: torch.nn.functional.interpolate(u, scale_factor=op.scale_factor)
        comp = lambda: x - up(op.forward(x)) + up(y)
    else:
        comp = lambda: (y - op.forward(y)) - op.forward(x)
    want = comp()
    first = op.project(x, y)
    same = all(torch.equal(op.project(x, y), first) for _ in range(40))
    err = float((first - want).abs().max() / want.abs().max())
    ok &= same and err <= 1e-6
    print(json.dumps({"operator": name + (f"x{cfg['scale_factor']}" if "scale_factor" in cfg else ""), "n": args.n,
                      "repeats_bit_identical": same, "equals_composition": bool(torch.equal(first, want)), "rel_err": err, "fused_us": round(timeit(lambda: op.project(x, y), args.iters), 2),
                      "composition_us": round(timeit(comp, args.iters), 2)}))
print("PASS" if ok else "FAIL")
sys.exit(0 if ok else 1)
