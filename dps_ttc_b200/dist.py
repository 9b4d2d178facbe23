"""Particle sharding across the GPUs of one box — new design, the reference is
single-process / single-GPU.

Rank r of W owns the global particles [r·n, (r+1)·n).  Between resampling steps everything is local.
At a resampling step
  1. all_gather of the per-particle log-weights (N·4 bytes),
  2. every rank runs the SAME weights → CDF → ancestors kernels on the full vector with the same
     uniforms (drawn from an identically seeded CPU generator), so ancestor indices are bit-identical
     on all ranks and identical to the single-GPU run — no broadcast needed,
  3. particles move: all_gather of the particle tensor + local gather kernel (v1; N·T bytes over NVSwitch).
Greedy search broadcasts the single best particle from its owner.

Backend: torch.distributed (NCCL on GPUs; the same code runs under gloo on CPU for the scalar paths,
which is how the world_size-2 tests exercise the index logic without GPUs).
"""
from __future__ import annotations

import torch
import torch.distributed as dist


class ParticleShards:
    def __init__(self, n_local: int, group=None):
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.n_local = int(n_local)
        self.total = self.n_local * self.world
        self.bytes_exchanged = 0

    @property
    def offset(self):
        return self.rank * self.n_local

    def local_slice(self, full):
        return full[self.offset:self.offset + self.n_local]

    def all_gather_scalars(self, local: torch.Tensor) -> torch.Tensor:
        """(n_local,) → (N,) in global particle order."""
        if self.world == 1:
            return local
        out = torch.empty(self.total, device=local.device, dtype=local.dtype)
        dist.all_gather_into_tensor(out, local.contiguous(), group=self.group)
        self.bytes_exchanged += out.numel() * out.element_size()
        return out

    def all_gather_particles(self, local: torch.Tensor) -> torch.Tensor:
        if self.world == 1:
            return local
        out = torch.empty((self.total,) + tuple(local.shape[1:]), device=local.device, dtype=local.dtype)
        dist.all_gather_into_tensor(out, local.contiguous(), group=self.group)
        self.bytes_exchanged += out.numel() * out.element_size()
        return out

    def exchange(self, img_local, dist_local, ancestors):
        """New local particles = global[ancestors[offset : offset+n_local]] (and their distances)."""
        mine = ancestors[self.offset:self.offset + self.n_local].contiguous()
        if img_local.is_cuda:
            from . import kernels
            full = self.all_gather_particles(img_local)
            new_img = kernels.gather_particles(full, mine)
        else:  # gloo / CPU path of the tests: index logic only
            new_img = self.all_gather_particles(img_local)[mine]
        new_dist = self.all_gather_scalars(dist_local)[mine]
        return new_img, new_dist

    def greedy_broadcast(self, img_local, costs_local):
        """img[argmin costs] replicated to every particle of every rank (first minimum, global order)."""
        costs = self.all_gather_scalars(costs_local)
        if img_local.is_cuda:
            from . import kernels
            best, _ = kernels.argmin(costs)
            if self.world == 1:
                return kernels.broadcast_particle(img_local, best, self.n_local)
            b = int(best.item())  # owner must be known on the host to pick the broadcast root
        else:
            b = int(torch.argmin(costs).item())
        owner, local_idx = divmod(b, self.n_local)
        buf = img_local[local_idx].clone() if self.rank == owner else torch.empty_like(img_local[0])
        if self.world > 1:
            dist.broadcast(buf, src=owner, group=self.group)
            self.bytes_exchanged += buf.numel() * buf.element_size()
        return buf.unsqueeze(0).expand(self.n_local, *buf.shape).contiguous()


def shared_uniforms(seed: int, idx: int, n: int, device):
    """fp64 uniforms every rank can reproduce: a CPU generator keyed by (seed, step)."""
    g = torch.Generator().manual_seed((int(seed) * 1_000_003 + int(idx)) & 0x7FFFFFFFFFFF)
    return torch.rand(n, dtype=torch.float64, generator=g).to(device, non_blocking=True)
