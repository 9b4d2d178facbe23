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
