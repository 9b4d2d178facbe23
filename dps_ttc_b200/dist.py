"""Particle sharding across the GPUs of one box — new design, the reference is
single-process / single-GPU.

Rank r of W owns the global particles [r·n, (r+1)·n).  Between resampling steps everything is local.
At a resampling step
  1. ONE all-gather of the per-particle (log-weight, distance) pairs (N·8 bytes, NCCL),
  2. every rank runs the SAME weights → CDF → ancestors kernels on the full vector with the same uniforms
     (drawn from an identically seeded CPU generator), so ancestor indices are bit-identical on all ranks and
     identical to the single-GPU run — no broadcast needed,
  3. particles move.  Transports:
       * p2p  (default on GPUs when symmetric memory is available): every rank keeps a double-buffered
         symmetric-memory particle buffer; on a resampling step the posterior-update kernel writes x_{t-1}
         STRAIGHT into it (`publish_target`, no staging copy) and ONE kernel (dps_exchange_particles_p2p) does the
         inter-GPU rendezvous (release/acquire flags in a symmetric signal pad), reads each needed ancestor from its
         owner's HBM over NVLink/NVSwitch and writes the new local particles — rendezvous, exchange and gather
         fused, at most n·T bytes per rank on the fabric, no barrier launches (resample.cu explains why the
         double buffer makes a trailing barrier unnecessary);
       * p2p_barrier: the round-1 form — staging copy + torch's symmetric-memory barrier kernels around the
         un-synchronised gather kernel (kept for A/B measurements);
       * allgather / all_to_all (NCCL; also the gloo/CPU test path): all-gather of all N particles + local gather
         kernel (W× the fabric traffic), or all_to_all_single of exactly the needed particles (needs the ancestor
         counts on the host: one sync).
Greedy search replicates the single best particle on every rank through the same transports.
"""
from __future__ import annotations

import torch
import torch.distributed as dist

TRANSPORTS = ("p2p", "p2p_barrier", "allgather", "all_to_all")


class ParticleShards:
    def __init__(self, n_local: int, group=None, p2p: bool | None = None, transport: str | None = None, timing: bool = False):
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.n_local = int(n_local)
        self.total = self.n_local * self.world
        self.bytes_exchanged = 0          # scalars + particle bytes that crossed (or could cross) the fabric, this rank
        self.exchanges = 0
        if transport is not None and transport not in TRANSPORTS:
            raise ValueError(f"transport must be one of {TRANSPORTS}")
        if transport is None:
            transport = None if p2p is None else ("p2p" if p2p else "allgather")
        self.want = transport             # None: p2p when possible, else allgather
        self.transport = "local" if self.world == 1 else (transport or "allgather")
        self._symm = None                 # (buffer (2, n, C, H, W), handle, signal pad, signal handle)
        self._epoch = 0
        self.timing = timing
        self.spans = []                   # (start event, end event) per exchange when timing=True

    @property
    def offset(self):
        return self.rank * self.n_local

    def local_slice(self, full):
        return full[self.offset:self.offset + self.n_local]

    # -- symmetric memory ---------------------------------------------------------------------------
    def _symmetric(self, like: torch.Tensor):
        """Lazily allocate the double-buffered symmetric particle buffer + the signal pad and exchange peer pointers."""
        shape = (2,) + tuple(like.shape)
        if self._symm is not None and tuple(self._symm[0].shape) == shape:
            return self._symm
        import torch.distributed._symmetric_memory as symm_mem
        grp = self.group if self.group is not None else dist.group.WORLD
        buf = symm_mem.empty(shape, dtype=torch.float32, device=like.device)
        hdl = symm_mem.rendezvous(buf, grp)
        sig = symm_mem.empty((max(64, self.world),), dtype=torch.int32, device=like.device)
        sig.zero_()
        sig_hdl = symm_mem.rendezvous(sig, grp)
        torch.cuda.synchronize(like.device)
        sig_hdl.barrier()                 # every pad is zero before anybody may signal
        torch.cuda.synchronize(like.device)
        self._symm = (buf, hdl, sig, sig_hdl)
        self._epoch = 0
        return self._symm

    def _p2p_ready(self, t: torch.Tensor) -> bool:
        if self.world == 1 or not t.is_cuda or self.want in ("allgather", "all_to_all"):
            return False
        try:
            self._symmetric(t)
        except Exception:  # noqa: BLE001  (no symmetric memory on this platform / backend)
            if self.want in ("p2p", "p2p_barrier"):
                raise
            self.want = "allgather"
            self.transport = "allgather"
            return False
        self.transport = self.want or "p2p"
        return True

    def publish_target(self, like: torch.Tensor):
        """Where the producer of the next exchange's particles (the posterior-update kernel) should write them: the half
        of the symmetric buffer the next exchange reads, or None when the transport needs no such buffer."""
        if not self._p2p_ready(like) or self.transport != "p2p":
            return None
        return self._symm[0][(self._epoch + 1) & 1]

    # -- collectives --------------------------------------------------------------------------------
    def all_gather_scalars(self, local: torch.Tensor) -> torch.Tensor:
        """(n_local, …) → (N, …) in global particle order."""
        if self.world == 1:
            return local
        local = local.contiguous()
        out = torch.empty((self.total,) + tuple(local.shape[1:]), device=local.device, dtype=local.dtype)
        dist.all_gather_into_tensor(out, local, group=self.group)
        self.bytes_exchanged += out.numel() * out.element_size()
        return out

    def all_gather_particles(self, local: torch.Tensor) -> torch.Tensor:
        return self.all_gather_scalars(local)

    def _move(self, img_local, mine):
        """new local particles = global[mine] (mine: this rank's n_local ancestor indices)."""
        from . import kernels
        tgt = None
        if self._p2p_ready(img_local):
            buf, hdl, sig, sig_hdl = self._symm
            if self.transport == "p2p":
                self._epoch += 1
                slot = self._epoch & 1
                tgt = buf[slot]
                if img_local.data_ptr() != tgt.data_ptr():      # the producer did not write into publish_target()
                    tgt.copy_(img_local)
                out = kernels.exchange_particles_p2p(hdl.buffer_ptrs_dev, sig_hdl.buffer_ptrs_dev, self.rank, self.world,
                                                     self._epoch, slot * tgt.numel(), self.n_local, mine, tgt)
            else:                                               # p2p_barrier
                buf[0].copy_(img_local)
                hdl.barrier()
                out = kernels.gather_particles_p2p(hdl.buffer_ptrs_dev, self.n_local, mine, img_local)
                hdl.barrier()
            self.bytes_exchanged += out.numel() * 4
            return out
        if self.transport == "all_to_all":
            return self._all_to_all(img_local, mine)
        full = self.all_gather_particles(img_local)
        return kernels.gather_particles(full, mine)

    def _all_to_all(self, img_local, mine_dev):
        """NCCL all_to_all_single of exactly the needed particles.  Every rank knows all ancestors, so the send and
        receive lists are computed locally — but the split sizes must be on the host (one device→host sync)."""
        anc = self._last_ancestors.cpu()
        n = self.n_local
        owners = anc // n
        send_idx, in_splits, out_splits = [], [], []
        for q in range(self.world):                              # what rank q wants from me, in q's destination order
            want = anc[q * n:(q + 1) * n]
            sel = want[(want // n) == self.rank] - self.rank * n
            send_idx.append(sel)
            in_splits.append(int(sel.numel()))
        mine = anc[self.rank * n:(self.rank + 1) * n]
        for q in range(self.world):
            out_splits.append(int(((mine // n) == q).sum()))
        send = img_local[torch.cat(send_idx).to(img_local.device)] if sum(in_splits) else img_local[:0]
        recv = torch.empty((sum(out_splits),) + tuple(img_local.shape[1:]), device=img_local.device, dtype=img_local.dtype)
        dist.all_to_all_single(recv, send.contiguous(), out_splits, in_splits, group=self.group)
        self.bytes_exchanged += recv.numel() * 4
        # received blocks are grouped by owner, each in my destination order: undo the grouping
        order = torch.argsort(owners[self.rank * n:(self.rank + 1) * n], stable=True)
        inv = torch.empty_like(order)
        inv[order] = torch.arange(n)
        return recv[inv.to(img_local.device)]

    def exchange(self, img_local, dist_local, ancestors, dist_all=None):
        """New local particles = global[ancestors[offset : offset+n_local]] (and their distances).  `dist_all`: the
        already all-gathered distances (the samplers gather them together with the log-weights)."""
        mine = ancestors[self.offset:self.offset + self.n_local].contiguous()
        if dist_all is None:
            dist_all = self.all_gather_scalars(dist_local)
        new_dist = dist_all[mine]
        self.exchanges += 1
        if not img_local.is_cuda:  # gloo / CPU path of the tests: index logic only
            if self.want == "all_to_all" and self.world > 1:
                self._last_ancestors = ancestors
                self.transport = "all_to_all"
                return self._all_to_all(img_local, mine), new_dist
            return self.all_gather_particles(img_local)[mine], new_dist
        self._last_ancestors = ancestors
        ev = None
        if self.timing:
            ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
            ev[0].record()
        out = self._move(img_local, mine)
        if ev is not None:
            ev[1].record()
            self.spans.append(ev)
        return out, new_dist

    def greedy_broadcast(self, img_local, costs_local):
        """img[argmin costs] replicated to every particle of every rank (first minimum, global order)."""
        costs = self.all_gather_scalars(costs_local)
        if img_local.is_cuda:
            from . import kernels
            best, _ = kernels.argmin(costs)
            if self.world == 1:
                return kernels.broadcast_particle(img_local, best, self.n_local)
            if self._p2p_ready(img_local):
                ids = best.expand(self.n_local).contiguous()
                return self._move(img_local, ids)
            b = int(best.item())  # owner must be known on the host to pick the broadcast root
        else:
            b = int(torch.argmin(costs).item())
        owner, local_idx = divmod(b, self.n_local)
        buf = img_local[local_idx].clone() if self.rank == owner else torch.empty_like(img_local[0])
        if self.world > 1:
            dist.broadcast(buf, src=owner, group=self.group)
            self.bytes_exchanged += buf.numel() * buf.element_size()
        return buf.unsqueeze(0).expand(self.n_local, *buf.shape).contiguous()

    def exchange_us(self):
        """Mean µs per exchange (CUDA events on the sampling stream); call after a synchronize."""
        if not self.spans:
            return None
        return 1e3 * sum(a.elapsed_time(b) for a, b in self.spans) / len(self.spans)


def shared_uniforms(seed: int, idx: int, n: int, device):
    """fp64 uniforms every rank can reproduce: a CPU generator keyed by (seed, step)."""
    g = torch.Generator().manual_seed((int(seed) * 1_000_003 + int(idx)) & 0x7FFFFFFFFFFF)
    return torch.rand(n, dtype=torch.float64, generator=g).to(device, non_blocking=True)
