"""Kernel variants that a launch picks by grid size must agree bit for bit: the sharded and the unsharded run of a
resampling sampler launch different variants (8 particles per rank vs 64 on one GPU) and still have to produce the same
per-particle norms, hence the same ancestor indices.  tools/variant_check.py runs one process per variant."""
import os
import subprocess
import sys

import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
@pytest.mark.parametrize("op,n", [("sr4", 12), ("sr8", 5)])
def test_resize_variants_bit_identical(op, n):
    cmd = [sys.executable, os.path.join(REPO, "tools", "variant_check.py"), "--op", op, "--n", str(n),
           "--env", "DPSTTC_RESIZE_VARIANT=big", "--env", "DPSTTC_RESIZE_VARIANT=small", "--env", "DPSTTC_RESIZE_VARIANT=stream"]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-3000:]
    assert "PASS" in res.stdout


@pytest.mark.gpu
@pytest.mark.parametrize("op,n", [("sr4", 8), ("sr8", 3)])
def test_resize_forward_ring_depths_bit_identical(op, n):
    """3-stage (4 CTAs/SM) and 6-stage (2 CTAs/SM, small grids) TMA rings of the strip forward: same chunks, same order."""
    cmd = [sys.executable, os.path.join(REPO, "tools", "variant_check.py"), "--op", op, "--n", str(n),
           "--env", "DPSTTC_RESIZE_VARIANT=big,DPSTTC_RESIZE_FWD_STAGES=3",
           "--env", "DPSTTC_RESIZE_VARIANT=big,DPSTTC_RESIZE_FWD_STAGES=6"]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-3000:]
    assert "PASS" in res.stdout


@pytest.mark.gpu
@pytest.mark.parametrize("op,n,envs", [
    ("sr4", 8, ("DPSTTC_RESIZE_FWD_LEAN=0,DPSTTC_RESIZE_ADJ_LEAN=0", "DPSTTC_RESIZE_FWD_LEAN=1,DPSTTC_RESIZE_ADJ_LEAN=1")),
    ("phase", 3, ("DPSTTC_PHASE_LEAN=0", "DPSTTC_PHASE_LEAN=1")),
])
def test_lean_kernels_bit_identical_to_round1_kernels(op, n, envs):
    """The lean strip forward / short-strip adjoint / phase epilogue (default since round 2) against the round-1 kernels."""
    cmd = [sys.executable, os.path.join(REPO, "tools", "variant_check.py"), "--op", op, "--n", str(n)]
    for e in envs:
        cmd += ["--env", e]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-3000:]
    assert "PASS" in res.stdout


@pytest.mark.gpu
def test_phase_register_kernels_agree_with_shared_memory_kernels():
    """The register-resident column / row kernels of the phase guidance (phase_colsreg.cuh, phase_rowsreg.cuh) against the
    shared-memory kernels, in every combination, on the same inputs at 256², 128² and 64²: residual, per-particle norms and
    cotangent to rounding."""
    cmd = [sys.executable, os.path.join(REPO, "tools", "phase_reg_check.py"), "--n", "3"]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-3000:]
    assert "PASS" in res.stdout


@pytest.mark.gpu
def test_motion_blur_row_variants_bit_identical():
    """blur_sparse.cu launches 16-row CTAs while the grid is small (N ≲ 12) and 32-row CTAs otherwise; a particle's residual,
    partial sums and cotangent must not depend on which one ran (sharded runs put 8 particles per launch where the unsharded
    run has 64).  Same particles alone (n = 8) and inside a larger batch (n = 16), bit for bit."""
    import numpy as np
    import torch
    from dps_ttc_b200 import tables
    from dps_ttc_b200.kernels import OperatorPlan
    from dps_ttc_b200.schedule import Schedule, named_beta_schedule
    dev = torch.device("cuda:0")
    np.random.seed(8)
    plan = OperatorPlan.blur(tables.motion_kernel(61, 0.5).astype(np.float32), 3, 256, 256, dev)
    k = Schedule(named_beta_schedule("linear", 1000)).consts(600)
    gen = torch.Generator(dev).manual_seed(3)
    x = torch.randn(16, 3, 256, 256, device=dev, generator=gen) / k.c1
    eps = torch.randn(16, 3, 256, 256, device=dev, generator=gen) * 0.3 / k.c2
    y = torch.randn(1, 3, 256, 256, device=dev, generator=gen)
    r_big, p_big, _ = plan.forward(x, eps, k, True, y, want_partials=True)
    r_small, p_small, _ = plan.forward(x[:8], eps[:8], k, True, y, want_partials=True)
    assert torch.equal(r_big[:8], r_small) and torch.equal(p_big[:8], p_small)
    g_big = torch.zeros(16, 3, 256, 256, device=dev)
    plan.adjoint(r_big, None, x, eps, k, True, None, out=g_big)
    g_small = torch.zeros(8, 3, 256, 256, device=dev)
    plan.adjoint(r_small, None, x[:8], eps[:8], k, True, None, out=g_small)
    assert torch.equal(g_big[:8], g_small)
    assert float(r_small.abs().max()) > 0 and float(g_small.abs().max()) > 0
