"""Particle sharding across the GPUs of one box — new design, the reference is
single-process / single-GPU.

Rank r of W owns the global particles [r·n, (r+1)·n).  Between resampling steps everything is local.
At a resampling step
  1. all_gather of the per-particle log-weights (N·4 bytes, NCCL),
  2. every rank runs the SAME weights → CDF → ancestors kernels on the full vector with the same uniforms
     (drawn from an identically seeded CPU generator), so ancestor indices are bit-identical on all ranks and
     identical to the single-GPU run — no broadcast needed,
  3. particles move.  Two transports:
       * p2p  (default on GPUs when symmetric memory is available): every rank keeps its particles in a
         symmetric-memory buffer; ONE kernel (dps_gather_particles_p2p) reads each needed ancestor straight
         from its owner's HBM over NVLink/NVSwitch and writes the new local particles — exchange and gather
         fused, n·T bytes per rank on the fabric;
       * allgather (fallback, and the gloo/CPU test path): NCCL all-gather of all N particles + local gather
         kernel — W× more fabric traffic.
Greedy search broadcasts the single best particle from its owner.
"""
from __future__ import annotations

import torch
import torch.distributed as dist


class ParticleShards:
    def __init__(self, n_local: int, group=None, p2p: bool | None = None):
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.n_local = int(n_local)
        self.total = self.n_local * self.world
        self.bytes_exchanged = 0
        self.want_p2p = p2p
        self._symm = None  # (buffer tensor, handle)
        self.transport = "local" if self.world == 1 else "allgather"

    @property
    def offset(self):
        return self.rank * self.n_local

    def local_slice(self, full):
        return full[self.offset:self.offset + self.n_local]

    # -- symmetric memory ---------------------------------------------------------------------------
    def _symmetric(self, like: torch.Tensor):
        """Lazily allocate the symmetric particle buffer (same shape on every rank) and exchange peer pointers."""
        if self._symm is not None and self._symm[0].shape == like.shape:
            return self._symm
        import torch.distributed._symmetric_memory as symm_mem
        buf = symm_mem.empty(tuple(like.shape), dtype=torch.float32, device=like.device)
        hdl = symm_mem.rendezvous(buf, self.group if self.group is not None else dist.group.WORLD)
        self._symm = (buf, hdl)
        return self._symm

    def _use_p2p(self, t: torch.Tensor) -> bool:
        if self.world == 1 or not t.is_cuda or self.want_p2p is False:
            return False
        try:
            self._symmetric(t)
            self.transport = "p2p"
            return True
        except Exception:  # noqa: BLE001  (no symmetric memory on this platform / backend)
            if self.want_p2p:
                raise
            self.want_p2p = False
            return False

    # -- collectives --------------------------------------------------------------------------------
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
        new_dist = self.all_gather_scalars(dist_local)[mine]
        if not img_local.is_cuda:  # gloo / CPU path of the tests: index logic only
            return self.all_gather_particles(img_local)[mine], new_dist
        from . import kernels
        if self._use_p2p(img_local):
            buf, hdl = self._symm
            buf.copy_(img_local)                       # publish my particles (local HBM copy)
            hdl.barrier()                              # everyone's buffer is written
            new_img = kernels.gather_particles_p2p(hdl.buffer_ptrs_dev, self.n_local, mine, img_local)
            hdl.barrier()                              # everyone has read: buffers may be overwritten
            self.bytes_exchanged += new_img.numel() * 4
            return new_img, new_dist
        full = self.all_gather_particles(img_local)
        return kernels.gather_particles(full, mine), new_dist

    def greedy_broadcast(self, img_local, costs_local):
        """img[argmin costs] replicated to every particle of every rank (first minimum, global order)."""
        costs = self.all_gather_scalars(costs_local)
        if img_local.is_cuda:
            from . import kernels
            best, _ = kernels.argmin(costs)
            if self.world == 1:
                return kernels.broadcast_particle(img_local, best, self.n_local)
            if self._use_p2p(img_local):
                buf, hdl = self._symm
                buf.copy_(img_local)
                hdl.barrier()
                ids = best.expand(self.n_local).contiguous()
                out = kernels.gather_particles_p2p(hdl.buffer_ptrs_dev, self.n_local, ids, img_local)
                hdl.barrier()
                return out
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
