"""A/B of the kernel variants of the fused phase-retrieval guidance: register-resident butterflies (phase_colsreg.cuh,
phase_rowsreg.cuh) against the shared-memory kernels they replace (DPSTTC_PHASE_COLS_REG / DPSTTC_PHASE_ROWS_REG = 0 / 1 for
the fused guidance, DPSTTC_PHASE_FWD_REG / DPSTTC_PHASE_ADJ_REG = 0 / 1 for the forward and adjoint pass of the two-kernel path).
One child process per variant (the switches are read once per process) runs dps_operator_guidance on the same seeded
inputs for 256², 128² and 64² images; the parent compares residual, per-particle norms and cotangent of every variant
with the all-shared-memory one.  The kernels do the same arithmetic per bin but accumulate the partial sums in a
different order, so the comparison is to rounding, not bit for bit.

    python tools/phase_reg_check.py [--n 3]
"""
from __future__ import annotations

import argparse
import os
import subprocess
import sys

import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
SIZES = (256, 128, 64)


def child(n, out_path):
    from dps_ttc_b200 import kernels
    from dps_ttc_b200.kernels import OperatorPlan
    from dps_ttc_b200.schedule import Schedule, named_beta_schedule
    dev = torch.device("cuda:0")
    k = Schedule(named_beta_schedule("linear", 1000)).consts(500)
    res = {}
    for size in SIZES:
        plan = OperatorPlan.phase(64, 3, size, size, dev)
        g = torch.Generator(dev).manual_seed(100 + size)
        x = torch.randn(n, 3, size, size, device=dev, generator=g) / k.c1
        o6 = torch.randn(n, 6, size, size, device=dev, generator=g) * 0.3 / k.c2
        y = torch.rand(1, 3, size + 128, size + 128, device=dev, generator=g) * 1.5
        out = torch.full((n, 3, size, size), float("nan"), device=dev)
        p, r, _ = plan.guidance(x, o6[:, :3], k, True, y, out=out, want_r=True)
        out2 = torch.full((n, 3, size, size), float("nan"), device=dev)
        p2, r2, _ = plan.guidance(x, o6[:, :3], k, True, y, out=out2)          # residual kept on chip
        norms = kernels.particle_norms(p, want_l1=True)
        # two-kernel path: forward (residual + partial sums + unit phase), adjoint from it; A(x) alone without ε / y
        rf, pf, aux = plan.forward(x, o6[:, :3], k, True, y, want_partials=True)
        nf = kernels.particle_norms(pf, want_l1=True)
        gf = torch.full((n, 3, size, size), float("nan"), device=dev)
        plan.adjoint(rf, None, x, o6[:, :3], k, True, None, out=gf, aux=aux)
        ax, _, _ = plan.forward(x * 0.01)
        torch.cuda.synchronize()
        res[size] = {"r": r.cpu(), "g": out.cpu(), "l2": norms[0].cpu(), "l1": norms[1].cpu(),
                     "fwd_r": rf.cpu(), "fwd_l2": nf[0].cpu(), "fwd_l1": nf[1].cpu(), "fwd_adj_g": gf.cpu(), "Ax": ax.cpu(),
                     "same_without_r": bool(r2 is None and torch.equal(out, out2) and torch.equal(p, p2))}
    torch.save(res, out_path)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=3)
    ap.add_argument("--child", default="")
    a = ap.parse_args()
    if a.child:
        child(a.n, a.child)
        return
    os.makedirs(os.path.join(REPO, "gpurun_out"), exist_ok=True)
    outs = {}
    variants = ("0000", "1010", "0101", "1111")   # (fused columns, fused rows, two-kernel forward, adjoint) in registers
    for v in variants:
        path = os.path.join(REPO, "gpurun_out", f"_phase_reg_{v}.pt")
        subprocess.run([sys.executable, os.path.abspath(__file__), "--n", str(a.n), "--child", path],
                       env=dict(os.environ, DPSTTC_PHASE_COLS_REG=v[0], DPSTTC_PHASE_ROWS_REG=v[1], DPSTTC_PHASE_FWD_REG=v[2],
                                DPSTTC_PHASE_ADJ_REG=v[3]),
                       check=True)
        outs[v] = torch.load(path)
        os.remove(path)
    ok = True
    for size in SIZES:
        a0 = outs["0000"][size]
        ok &= a0["same_without_r"]
        for v in variants[1:]:
            a1 = outs[v][size]
            for key, tol in (("r", 2e-6), ("g", 2e-5), ("l2", 2e-6), ("l1", 2e-6), ("fwd_r", 2e-6), ("fwd_l2", 2e-6),
                             ("fwd_l1", 2e-6), ("fwd_adj_g", 2e-5), ("Ax", 2e-6)):
                diff = (a0[key] - a1[key]).abs().max().item()
                scale = max(1.0, a0[key].abs().max().item()) if key in ("r", "g", "fwd_r", "fwd_adj_g", "Ax") else a0[key].abs().max().item()
                good = bool(torch.isfinite(a1[key]).all()) and diff <= tol * scale
                ok &= good
                print(f"[phase_reg_check] {size}x{size} n={a.n} cols/rows/fwd/adj in registers = {v} {key}: max|smem - reg| = {diff:.3e} "
                      f"(scale {scale:.3e}, tol {tol:g}) {'ok' if good else 'FAIL'}", flush=True)
            ok &= a1["same_without_r"]
            print(f"[phase_reg_check] {size}x{size} {v}: with / without r_out bit-identical: {a1['same_without_r']}")
    print("[phase_reg_check]", "PASS" if ok else "MISMATCH")
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
