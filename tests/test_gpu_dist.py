"""Sharded == unsharded on real GPUs: runs tools/dist_check.py under torchrun with one rank per GPU (2 ranks; needs a box
with at least 2 GPUs — skipped on a single-GPU box, where the world_size-2 gloo test covers the host logic).  Every
transport of dps_ttc_b200.dist (fused P2P exchange kernel with in-kernel rendezvous, P2P gather between symmetric-memory
barriers, NCCL all-gather, NCCL all_to_all) must reproduce the unsharded run: ancestor indices bit-identical on every rank
at every resampling step, particles and distances equal; same for the greedy search broadcast."""
import os
import subprocess
import sys

import pytest
import torch

from helpers import REPO

pytestmark = pytest.mark.gpu


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs (NCCL does not put two ranks on one device)")
def test_sharded_samplers_match_unsharded_on_two_gpus():
    port = 29400 + os.getpid() % 500
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(REPO, "tools", "dist_check.py")]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    assert res.returncode == 0, res.stdout[-4000:] + res.stderr[-4000:]
    assert "FAIL" not in res.stdout
