"""Bitwise A/B check of kernel variants that are selected by environment variables read once per process
(DPSTTC_RESIZE_VARIANT, DPSTTC_RESIZE_FWD, DPSTTC_LIB, ...): runs the operator's forward and adjoint on the same
seeded inputs in one child process per variant and compares the outputs bit for bit.

    python tools/variant_check.py --op sr4 --n 40 --env DPSTTC_RESIZE_VARIANT=big --env DPSTTC_RESIZE_VARIANT=stream
"""
from __future__ import annotations

import argparse
import os
import subprocess
import sys

import numpy as np
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)


def build_plan(op, dev):
    from dps_ttc_b200 import tables
    from dps_ttc_b200.kernels import OperatorPlan
    if op in ("sr4", "sr8"):
        (fh, wh), (fw, ww), _ = tables.resizer_tables((1, 3, 256, 256), 0.25 if op == "sr4" else 0.125)
        return OperatorPlan.resize(fh, wh, fw, ww, 3, 256, 256, dev)
    if op == "gauss":
        return OperatorPlan.blur(tables.gaussian_kernel(61, 3.0).astype(np.float32), 3, 256, 256, dev)
    if op == "motion":
        np.random.seed(8)
        return OperatorPlan.blur(tables.motion_kernel(61, 0.5).astype(np.float32), 3, 256, 256, dev)
    if op == "phase":
        return OperatorPlan.phase(64, 3, 256, 256, dev)
    raise SystemExit(f"unknown op {op}")


def child(op, n, out_path):
    from dps_ttc_b200.schedule import Schedule, named_beta_schedule
    dev = torch.device("cuda:0")
    plan = build_plan(op, dev)
    k = Schedule(named_beta_schedule("linear", 1000)).consts(500)
    g = torch.Generator(dev).manual_seed(11)
    rnd = lambda *s: torch.randn(*s, device=dev, generator=g)  # noqa: E731
    x = rnd(n, 3, 256, 256) / k.c1
    o6 = rnd(n, 6, 256, 256) * 0.3 / k.c2
    oc, oh, ow = plan.out_shape
    y = rnd(1, oc, oh, ow)
    aux = plan.new_aux(n)
    r, partials, aux = plan.forward(x, o6[:, :3], k, True, y, want_partials=True, aux=aux)
    coef = -0.01 * (1.0 + torch.arange(n, device=dev, dtype=torch.float32) / n)
    gbuf = torch.zeros(n, 6, 256, 256, device=dev)
    plan.adjoint(r, coef, x, o6[:, :3], k, True, None, out=gbuf[:, :3], aux=aux)
    extra = rnd(n, 3, 256, 256) * 1e-3
    g2 = torch.zeros(n, 3, 256, 256, device=dev)
    plan.adjoint(r, coef, x, o6[:, :3], k, True, extra, out=g2, aux=aux)
    # the other argument shapes a launch can have: x̂₀ = x (no ε), no measurement, no partial sums, no clamp mask
    ax, _, _ = plan.forward(x * 0.01)
    r_plain, p_plain, _ = plan.forward(x * 0.01, None, None, False, y, want_partials=True, aux=plan.new_aux(n))
    g3 = torch.zeros(n, 3, 256, 256, device=dev)
    plan.adjoint(r, None, None, None, None, False, None, out=g3, aux=aux)
    torch.cuda.synchronize()
    torch.save({"r": r.cpu(), "partials": partials.cpu(), "g": gbuf[:, :3].cpu(), "g_extra": g2.cpu(), "Ax": ax.cpu(),
                "r_plain": r_plain.cpu(), "partials_plain": p_plain.cpu(), "g_nomask": g3.cpu()}, out_path)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--op", default="sr4")
    ap.add_argument("--n", type=int, default=40)
    ap.add_argument("--env", action="append", default=[], help="KEY=VALUE[,KEY=VALUE] of one variant (repeatable)")
    ap.add_argument("--child", default="")
    a = ap.parse_args()
    if a.child:
        child(a.op, a.n, a.child)
        return
    os.makedirs(os.path.join(REPO, "gpurun_out"), exist_ok=True)
    outs = []
    for i, spec in enumerate(a.env):
        env = dict(os.environ)
        for kv in filter(None, spec.split(",")):
            key, val = kv.split("=", 1)
            env[key] = val
        path = os.path.join(REPO, "gpurun_out", f"_variant_{a.op}_{i}.pt")
        subprocess.run([sys.executable, os.path.abspath(__file__), "--op", a.op, "--n", str(a.n), "--child", path],
                       env=env, check=True)
        outs.append((spec, torch.load(path)))
        os.remove(path)
    base_spec, base = outs[0]
    ok = True
    for spec, o in outs[1:]:
        for key in base:
            same = torch.equal(base[key], o[key])
            diff = (base[key] - o[key]).abs().max().item()
            scale = base[key].abs().max().item()
            print(f"[variant_check] {a.op} n={a.n} {key}: [{base_spec}] vs [{spec}] bit-identical={same} "
                  f"max|diff|={diff:.3e} (scale {scale:.3e})", flush=True)
            ok &= same
    print("[variant_check]", "PASS" if ok else "MISMATCH")
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
