"""Multi-GPU check of the particle-sharded samplers (run under torchrun, one rank per GPU):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tools/dist_check.py

1. ttc_ddim + ps (Gaussian deblur, 64×64, N = 4 per rank, CPU-bridged stand-in model): the sharded run must
   reproduce the unsharded N-particle run — ancestor indices bit-identical on every rank and at every resampling
   step, particles equal — for every transport (fused P2P exchange kernel with in-kernel rendezvous; P2P gather between
   symmetric-memory barriers; NCCL all-gather + gather kernel; NCCL all_to_all_single).
2. search_ddpm greedy broadcast, same comparison.
3. Timing of the particle exchange at 256×256, 8 particles per rank, per transport."""
from __future__ import annotations

import os
import sys

import numpy as np
import torch
import torch.distributed as dist

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, "tests"))

from helpers import CpuBridge, TinyEps  # noqa: E402
from dps_ttc_b200 import kernels  # noqa: E402
from dps_ttc_b200.dist import TRANSPORTS, ParticleShards, shared_uniforms  # noqa: E402
from dps_ttc_b200.registry import get_conditioning_method, get_noise, get_operator  # noqa: E402
from dps_ttc_b200.sampler import NoiseTape, create_sampler  # noqa: E402

DIFF = dict(steps=1000, noise_schedule="linear", model_mean_type="epsilon", model_var_type="learned_range",
            dynamic_threshold=False, clip_denoised=True, rescale_timesteps=True)


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device(f"cuda:{local}")
    dist.init_process_group("nccl", device_id=dev)
    torch.backends.cudnn.allow_tf32 = False
    ok = True
    n_local, size, steps = 4, 64, 12
    N = n_local * world
    g = torch.Generator().manual_seed(5)
    x_true = torch.rand(1, 3, size, size, generator=g) * 2 - 1
    x_start = torch.randn(N, 3, size, size, generator=g)
    zs = {i: torch.randn(N, 3, size, size, generator=g) for i in range(steps)}

    def run(sampler_name, shards, lo, hi):
        op = get_operator("gaussian_blur", kernel_size=61, intensity=3.0, device=dev)
        cond = get_conditioning_method("ps", op, get_noise("gaussian", sigma=0.05), scale=0.3)
        s = create_sampler(sampler=sampler_name, timestep_respacing=str(steps), **DIFF)
        # the unsharded run replays the uniforms the sharded ranks derive from (seed, step)
        uni = {i: shared_uniforms(0, i, N, "cpu") for i in range(steps)} if shards is None else None
        s.noise, s.parity_rng = NoiseTape(z={i: z[lo:hi].contiguous() for i, z in zs.items()}, uniforms=uni), False
        y = op.forward(x_true.to(dev)).detach()
        model = CpuBridge(TinyEps(seed=21))
        kw = dict(model=model, x_start=x_start[lo:hi].to(dev), measurement=y, measurement_cond_fn=cond.conditioning,
                  record=False, save_root=None, shards=shards)
        if sampler_name == "search_ddpm":
            return s.p_sample_loop(operator=op, **kw), None, s
        img, d = s.p_sample_loop(**kw)
        return img, d, s

    for sampler_name in ("ttc_ddim", "search_ddpm"):
        full_img, full_d, s_full = run(sampler_name, None, 0, N)          # unsharded, every rank computes it
        for transport in TRANSPORTS:
            shards = ParticleShards(n_local, transport=transport)
            img, d, s = run(sampler_name, shards, shards.offset, shards.offset + n_local)
            want = full_img[shards.offset:shards.offset + n_local]
            err = float((img - want).abs().max())
            same_ids = True
            if sampler_name == "ttc_ddim":
                a, b = s_full.last_stats["ancestors"], s.last_stats["ancestors"]
                same_ids = a.keys() == b.keys() and all(torch.equal(a[k], b[k]) for k in a)
                err = max(err, float((d - full_d[shards.offset:shards.offset + n_local]).abs().max()))
            good = same_ids and err <= 1e-5
            ok &= good
            print(f"[rank {rank}] {sampler_name:12s} transport={shards.transport:9s} ancestors_equal={same_ids} "
                  f"max|Δparticles|={err:.2e} exchanged={shards.bytes_exchanged / 1e6:.2f} MB {'PASS' if good else 'FAIL'}",
                  flush=True)

    # ---- exchange timing at the BASELINE particle size ----
    n_local = 8
    x = torch.randn(n_local, 3, 256, 256, device=dev)
    d = torch.rand(n_local, device=dev)
    ids = torch.randint(0, n_local * world, (n_local * world,), generator=torch.Generator().manual_seed(3)).to(dev)
    for transport in TRANSPORTS:
        sh = ParticleShards(n_local, transport=transport)
        for _ in range(3):
            sh.exchange(x, d, ids)
        torch.cuda.synchronize(); dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            out, _ = sh.exchange(x, d, ids)
        e1.record(); torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) / 20], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ref = sh.all_gather_particles(x)[ids[sh.offset:sh.offset + n_local]]
        good = torch.equal(out, ref)
        ok &= good
        if rank == 0:
            print(f"exchange of {n_local} particles/rank (256x256) transport={sh.transport:11s}: {float(t) * 1e3:8.1f} us  "
                  f"bit-equal to all-gather reference: {good}", flush=True)
    flag = torch.tensor([int(ok)], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if int(flag) else 1)


if __name__ == "__main__":
    main()
