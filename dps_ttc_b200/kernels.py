"""Tensor-level wrappers of the C ABI: one Python function per entry point of include/dpsttc.h.
All outputs are torch tensors allocated by the caller or here with torch.empty (the library never
allocates in a launch); all launches go to torch's current stream.
"""
from __future__ import annotations

import contextlib
import ctypes as C

import numpy as np
import torch

from . import _lib
from ._lib import DpsError, check, lib, make_consts, make_source, particle_view, ptr, require_cuda_f32, stream_ptr


class KernelTimer:
    """CUDA-event brackets around libdpsttc launches on the launching (current) stream.  Off by default;
    bench.py installs one (`kernels.TIMER = KernelTimer()`) to measure per-kernel durations live."""

    def __init__(self):
        self.spans = {}

    def start(self, name):
        e = torch.cuda.Event(enable_timing=True)
        e.record()
        return name, e

    def stop(self, tok):
        e = torch.cuda.Event(enable_timing=True)
        e.record()
        self.spans.setdefault(tok[0], []).append((tok[1], e))

    def summary(self):
        """name → (launches, mean ms); call after a synchronize."""
        return {n: (len(v), sum(a.elapsed_time(b) for a, b in v) / len(v)) for n, v in self.spans.items()}


TIMER = None
_NULL = contextlib.nullcontext()


def _on(device):
    """Context that makes `device` the current CUDA device when it is not already (a launch goes to the current device;
    the operator tables and the per-device shared-memory opt-in belong to the plan's device)."""
    return _NULL if torch.cuda.current_device() == device.index else torch.cuda.device(device)


def _chw(x: torch.Tensor) -> int:
    n = 1
    for s in x.shape[1:]:
        n *= s
    return n


# ------------------------------------------------------------------------------------------------
# posterior update
# ------------------------------------------------------------------------------------------------
def x0_from_eps(x, eps, k, clip=True, out=None):
    """x̂₀ = clamp(c1·x − c2·ε)  — EpsilonXMeanProcessor.predict_xstart + process_xstart."""
    src = make_source(x, eps, k.c1, k.c2, clip)
    out = torch.empty(x.shape, device=x.device, dtype=torch.float32) if out is None else out
    check(lib().dps_x0_from_eps(C.byref(src), out.data_ptr(), x.shape[0], _chw(x), stream_ptr(x.device)),
          "dps_x0_from_eps")
    return out


def posterior_update(sampler: str, x, eps, v, z, k, *, clip=True, g=None, vjp=None, var_mode=0, max_log=None,
                     want_sample=False, want_x0=False, out=None, deferred=None, philox=None):
    """Fused x̂₀ / posterior mean / log-variance / σ·z / guidance step.  Returns (x_next, sample, x0);
    the last two are None unless requested.  `g` may be a channel-slice view (e.g. of an (N,6,H,W)
    cotangent buffer).
    deferred=(partials, coef_mode, scale[, l2_out]): g and vjp are UNSCALED (g = mask ⊙ Aᵀr); the kernel derives the
    per-particle coefficient from the residual kernel's partial sums (and writes ‖r‖ to l2_out) — see dps_update_ext.
    philox=(seed, step, particle_offset): with z=None the noise is generated in the kernel (Philox4x32-10 + Box–Muller)."""
    if deferred is not None or philox is not None:
        if want_sample or want_x0:
            raise DpsError("the extended update writes x_next only")
        return _posterior_update_ext(sampler, x, eps, v, z, k, clip, g, vjp, var_mode, max_log, out, deferred, philox), None, None
    n, chw = x.shape[0], _chw(x)
    src = make_source(x, eps, k.c1, k.c2, clip)
    kc = make_consts(k, var_mode, max_log)
    dev = x.device
    x_next = torch.empty(x.shape, device=dev, dtype=torch.float32) if out is None else out
    sample = torch.empty(x.shape, device=dev, dtype=torch.float32) if want_sample else None
    x0 = torch.empty(x.shape, device=dev, dtype=torch.float32) if want_x0 else None
    gp, gs = (None, 0) if g is None else particle_view(g, "g")
    if vjp is not None:
        vjp = _lib.dense(vjp, "vjp")
    if z is not None:
        z = _lib.dense(z, "z")
    tok = TIMER.start(f"posterior_update_{sampler}") if TIMER else None
    if sampler == "ddpm":
        vp, vs = (None, 0) if v is None else particle_view(v, "v")
        rc = lib().dps_posterior_update_ddpm(C.byref(src), vp, vs, ptr(z), gp, gs, ptr(vjp), C.byref(kc),
                                             x_next.data_ptr(), ptr(sample), ptr(x0), n, chw, stream_ptr(dev))
    elif sampler == "ddim":
        rc = lib().dps_posterior_update_ddim(C.byref(src), ptr(z), gp, gs, ptr(vjp), C.byref(kc), x_next.data_ptr(),
                                             ptr(sample), ptr(x0), n, chw, stream_ptr(dev))
    else:
        raise DpsError(f"unknown sampler kind {sampler!r}")
    check(rc, f"dps_posterior_update_{sampler}")
    if tok:
        TIMER.stop(tok)
    return x_next, sample, x0


def _posterior_update_ext(sampler, x, eps, v, z, k, clip, g, vjp, var_mode, max_log, out, deferred, philox):
    n, chw = x.shape[0], _chw(x)
    src = make_source(x, eps, k.c1, k.c2, clip)
    kc = make_consts(k, var_mode, max_log)
    dev = x.device
    x_next = torch.empty(x.shape, device=dev, dtype=torch.float32) if out is None else out
    ext = _lib.UpdateExt()
    keep = []
    if deferred is not None:
        partials, mode, scale = deferred[:3]
        l2_out = deferred[3] if len(deferred) > 3 else None
        require_cuda_f32(partials, "partials")
        if partials.shape[0] != n or partials.shape[-1] != 2 or not partials.is_contiguous():
            raise DpsError(f"partials must be a contiguous (N, P, 2) tensor, got {tuple(partials.shape)}")
        ext.partials, ext.P, ext.coef_mode, ext.scale = partials.data_ptr(), partials.shape[1], int(mode), float(scale)
        ext.l2_out = ptr(l2_out)
        keep.append(partials)
    if philox is not None:
        seed, step, offset = philox
        ext.use_philox, ext.philox_seed, ext.philox_step, ext.particle_offset = 1, int(seed) & (2**64 - 1), int(step), int(offset)
    gp, gs = (None, 0) if g is None else particle_view(g, "g")
    if vjp is not None:
        vjp = _lib.dense(vjp, "vjp")
    if z is not None:
        z = _lib.dense(z, "z")
    tok = TIMER.start(f"posterior_update_{sampler}") if TIMER else None
    if sampler == "ddpm":
        vp, vs = (None, 0) if v is None else particle_view(v, "v")
        rc = lib().dps_posterior_update_ddpm_ext(C.byref(src), vp, vs, ptr(z), gp, gs, ptr(vjp), C.byref(kc), C.byref(ext),
                                                 x_next.data_ptr(), n, chw, stream_ptr(dev))
    elif sampler == "ddim":
        rc = lib().dps_posterior_update_ddim_ext(C.byref(src), ptr(z), gp, gs, ptr(vjp), C.byref(kc), C.byref(ext),
                                                 x_next.data_ptr(), n, chw, stream_ptr(dev))
    else:
        raise DpsError(f"unknown sampler kind {sampler!r}")
    check(rc, f"dps_posterior_update_{sampler}_ext")
    if tok:
        TIMER.stop(tok)
    return x_next


def guidance_grad(g, vjp, k, out=None):
    """grad = c1·g − c2·vjp as a tensor (the chain rule the update kernel applies internally); `g` may be a
    channel-slice view.  Used on DiffStateGrad projection steps (gaussian_diffusion.py:240-253)."""
    n, chw = g.shape[0], _chw(g)
    gp, gs = particle_view(g, "g")
    if vjp is not None:
        vjp = _lib.dense(vjp, "vjp")
    out = torch.empty((n,) + tuple(g.shape[1:]), device=g.device, dtype=torch.float32) if out is None else out
    check(lib().dps_guidance_grad(gp, gs, ptr(vjp), k.c1, k.c2, out.data_ptr(), n, chw, stream_ptr(g.device)),
          "dps_guidance_grad")
    return out


def apply_gradient(sample, grad, out=None):
    """x' = sample − grad; a gradient of batch 1 is applied to every particle (what the reference's projected
    gradient does, gaussian_diffusion.py:255)."""
    sample, grad = _lib.dense(sample, "sample"), _lib.dense(grad, "grad")
    n, chw = sample.shape[0], _chw(sample)
    if grad.shape[0] not in (1, n) or _chw(grad) != chw:
        raise DpsError(f"grad {tuple(grad.shape)} does not match sample {tuple(sample.shape)}")
    out = torch.empty_like(sample) if out is None else out
    stride = 0 if (grad.shape[0] == 1 and n > 1) else chw
    check(lib().dps_apply_gradient(sample.data_ptr(), grad.data_ptr(), stride, out.data_ptr(), n, chw,
                                   stream_ptr(sample.device)), "dps_apply_gradient")
    return out


def q_sample(y, noise, a: float, b: float, out=None):
    y = _lib.dense(y, "y")
    noise = _lib.dense(noise, "noise")
    out = torch.empty_like(y) if out is None else out
    check(lib().dps_q_sample(y.data_ptr(), noise.data_ptr(), float(a), float(b), out.data_ptr(), y.numel(),
                             stream_ptr(y.device)), "dps_q_sample")
    return out


# ------------------------------------------------------------------------------------------------
# operator plans
# ------------------------------------------------------------------------------------------------
class OperatorPlan:
    """Owns a dps_operator handle (immutable device tables) for one CUDA device."""

    def __init__(self, handle, device):
        self._h = handle
        self.device = torch.device(device)
        info = _lib.OperatorInfo()
        check(lib().dps_operator_get_info(self._h, C.byref(info)), "dps_operator_get_info")
        self.kind = _lib.OP_KINDS.get(info.kind, str(info.kind))
        self.in_shape = (info.C, info.H, info.W)
        self.out_shape = (info.out_C, info.out_H, info.out_W)
        self.partials_per_particle = info.partials_per_particle
        self.aux_floats = info.aux_floats_per_particle
        self.taps = info.taps
        self.guidance_partials = info.guidance_partials      # > 0: dps_operator_guidance is ONE fused kernel

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            try:
                lib().dps_operator_destroy(h)
            except Exception:
                pass

    # -- constructors -----------------------------------------------------------------------------
    @staticmethod
    def _create(fn_name, device, *args):
        device = torch.device(device)
        if device.type != "cuda":
            raise DpsError("dps_ttc_b200 operators need a CUDA device (no CPU path)")
        if device.index is None:
            device = torch.device("cuda", torch.cuda.current_device())
        handle = C.c_void_p()
        with torch.cuda.device(device):
            check(getattr(lib(), fn_name)(*args, C.byref(handle)), fn_name)
        return OperatorPlan(handle, device)

    @classmethod
    def inpainting(cls, mask_hw: np.ndarray, C_, H, W, device):
        m = np.ascontiguousarray(mask_hw, dtype=np.float32).reshape(H, W)
        return cls._create("dps_operator_create_inpainting", device, m.ctypes.data, C_, H, W)

    @classmethod
    def blur(cls, kernel: np.ndarray, C_, H, W, device, mode=0):
        k = np.ascontiguousarray(kernel, dtype=np.float32)
        if k.ndim != 2 or k.shape[0] != k.shape[1]:
            raise DpsError(f"blur kernel must be (k,k), got {k.shape}")
        return cls._create("dps_operator_create_blur", device, k.ctypes.data, k.shape[0], C_, H, W, mode)

    @classmethod
    def resize(cls, fov_h, w_h, fov_w, w_w, C_, H, W, device):
        fh = np.ascontiguousarray(fov_h, dtype=np.int32)
        wh = np.ascontiguousarray(w_h, dtype=np.float32)
        fw = np.ascontiguousarray(fov_w, dtype=np.int32)
        ww = np.ascontiguousarray(w_w, dtype=np.float32)
        if fh.shape != wh.shape or fw.shape != ww.shape or fh.ndim != 2 or fw.ndim != 2:
            raise DpsError("resize tables must be (taps, out_len) pairs")
        return cls._create("dps_operator_create_resize", device, fh.ctypes.data, wh.ctypes.data, fh.shape[0],
                           fh.shape[1], fw.ctypes.data, ww.ctypes.data, fw.shape[0], fw.shape[1], C_, H, W)

    @classmethod
    def phase(cls, pad, C_, H, W, device):
        return cls._create("dps_operator_create_phase", device, int(pad), C_, H, W)

    # -- launches ---------------------------------------------------------------------------------
    def _check_in(self, x, name):
        require_cuda_f32(x, name)
        if x.device != self.device:
            raise DpsError(f"{name} is on {x.device} but this operator plan's tables live on {self.device}")
        if tuple(x.shape[1:]) != self.in_shape:
            raise DpsError(f"{name}: expected (N,{self.in_shape}), got {tuple(x.shape)}")

    def new_aux(self, n):
        return torch.empty((n, self.aux_floats), device=self.device, dtype=torch.float32) if self.aux_floats else None

    def forward(self, x, eps=None, k=None, clip=False, y=None, want_partials=False, aux=None, out=None):
        """out = A(x̂₀) (y None) or y − A(x̂₀); x̂₀ = x when eps is None.  Returns (out, partials)."""
        self._check_in(x, "x")
        n = x.shape[0]
        src = make_source(x, eps, k.c1 if k else 1.0, k.c2 if k else 0.0, clip)
        if out is None:
            out = torch.empty((n,) + self.out_shape, device=x.device, dtype=torch.float32)
        yp, ys = None, 0
        if y is not None:
            require_cuda_f32(y, "y")
            if tuple(y.shape[-3:]) != self.out_shape:
                raise DpsError(f"measurement shape {tuple(y.shape)} does not match operator output {self.out_shape}")
            y = y.reshape((-1,) + self.out_shape)
            if y.shape[0] not in (1, n):
                raise DpsError(f"measurement batch {y.shape[0]} must be 1 or {n}")
            y = _lib.dense(y, "y")
            yp, ys = y.data_ptr(), (0 if y.shape[0] == 1 else y[0].numel())
        partials = (torch.empty((n, self.partials_per_particle, 2), device=x.device, dtype=torch.float32)
                    if want_partials else None)
        if self.aux_floats and aux is None:
            aux = self.new_aux(n)
        tok = TIMER.start(f"{self.kind}_forward") if TIMER else None
        with _on(self.device):
            check(lib().dps_operator_forward(self._h, C.byref(src), yp, ys, out.data_ptr(), ptr(partials), ptr(aux), n,
                                             stream_ptr(x.device)), f"dps_operator_forward[{self.kind}]")
        if tok:
            TIMER.stop(tok)
        return out, partials, aux

    def guidance(self, x, eps, k, clip, y, out, aux=None, want_r=False):
        """The operator part of a guided step with the per-particle coefficient DEFERRED to the update kernel:
        out ← 1[|pre| ≤ 1] ⊙ Aᵀ(y − A x̂₀) (unscaled), returns (partials, r or None, aux).  One fused cluster kernel where
        the plan has one (`guidance_partials` > 0: r stays on chip unless want_r), else forward + adjoint launches."""
        self._check_in(x, "x")
        if eps is None or k is None:
            raise DpsError("guidance(): x̂₀ is formed from x and ε — pass eps and the step constants")
        n = x.shape[0]
        src = make_source(x, eps, k.c1, k.c2, clip)
        require_cuda_f32(y, "y")
        if tuple(y.shape[-3:]) != self.out_shape:
            raise DpsError(f"measurement shape {tuple(y.shape)} does not match operator output {self.out_shape}")
        y = _lib.dense(y.reshape((-1,) + self.out_shape), "y")
        if y.shape[0] not in (1, n):
            raise DpsError(f"measurement batch {y.shape[0]} must be 1 or {n}")
        ys = 0 if y.shape[0] == 1 else y[0].numel()
        fused = self.guidance_partials > 0
        P = self.guidance_partials if fused else self.partials_per_particle
        partials = torch.empty((n, P, 2), device=x.device, dtype=torch.float32)
        r = None
        if want_r or not fused:
            r = torch.empty((n,) + self.out_shape, device=x.device, dtype=torch.float32)
        if self.aux_floats and aux is None:
            aux = self.new_aux(n)
        gp, gs = particle_view(out, "g")
        tok = TIMER.start(f"{self.kind}_guidance" if fused else f"{self.kind}_forward+adjoint") if TIMER else None
        with _on(self.device):
            check(lib().dps_operator_guidance(self._h, C.byref(src), y.data_ptr(), ys, ptr(r), gp, gs, partials.data_ptr(),
                                              ptr(aux), n, stream_ptr(x.device)), f"dps_operator_guidance[{self.kind}]")
        if tok:
            TIMER.stop(tok)
        return partials, r, aux

    def project(self, data, y=None, out=None):
        """Operator.project(data, y) — or ortho_project(data) when y is None — with the elementwise part in the operator
        kernels' epilogues (dps_operator_project; measurements.py:48-54, :90-91).  No autograd."""
        self._check_in(data, "data")
        n = data.shape[0]
        dp, ds = particle_view(data, "data")
        if out is None:
            out = torch.empty((n,) + self.in_shape, device=data.device, dtype=torch.float32)
        op_, os_ = particle_view(out, "out")
        yp, ys, ny = None, 0, 0
        if y is not None:
            require_cuda_f32(y, "y")
            if tuple(y.shape[-3:]) != self.out_shape:
                raise DpsError(f"measurement shape {tuple(y.shape)} does not match operator output {self.out_shape}")
            y = _lib.dense(y.reshape((-1,) + self.out_shape), "y")
            ny = y.shape[0]
            if ny not in (1, n):
                raise DpsError(f"measurement batch {ny} must be 1 or {n}")
            yp, ys = y.data_ptr(), y[0].numel()
        scratch = None
        if self.kind == "resize":
            if self.guidance_partials == 0:
                scratch = torch.empty((n,) + self.out_shape, device=data.device, dtype=torch.float32)
        elif y is not None:
            scratch = torch.empty((ny,) + self.out_shape, device=data.device, dtype=torch.float32)
        tok = TIMER.start(f"{self.kind}_project") if TIMER else None
        with _on(self.device):
            check(lib().dps_operator_project(self._h, dp, ds, yp, ys, ny, op_, os_, ptr(scratch), n, stream_ptr(data.device)),
                  f"dps_operator_project[{self.kind}]")
        if tok:
            TIMER.stop(tok)
        return out

    def adjoint(self, r, coef=None, mask_x=None, mask_eps=None, k=None, clip=True, extra=None, out=None, aux=None):
        """g = 1[−1 ≤ c1·x − c2·ε ≤ 1] ⊙ (coef_n·Aᵀr + extra); mask only when mask_x/mask_eps given.
        `out` may be a channel-slice view of a larger buffer."""
        require_cuda_f32(r, "r")
        if r.device != self.device:
            raise DpsError(f"r is on {r.device} but this operator plan's tables live on {self.device}")
        n = r.shape[0]
        if tuple(r.shape[1:]) != self.out_shape:
            raise DpsError(f"r: expected (N,{self.out_shape}), got {tuple(r.shape)}")
        r = _lib.dense(r, "r")
        if out is None:
            out = torch.empty((n,) + self.in_shape, device=r.device, dtype=torch.float32)
        gp, gs = particle_view(out, "g")
        msrc = None
        if mask_x is not None and mask_eps is not None:
            msrc = C.byref(make_source(mask_x, mask_eps, k.c1, k.c2, clip))
        ep, es = (None, 0) if extra is None else particle_view(extra, "extra")
        if coef is not None:
            require_cuda_f32(coef, "coef")
        if self.aux_floats and aux is None:
            if self.kind == "phase":
                raise DpsError("phase retrieval adjoint needs the aux tensor its forward pass returned (the phase)")
            aux = self.new_aux(n)  # pure scratch (padded t of the sparse-blur adjoint)
        tok = TIMER.start(f"{self.kind}_adjoint") if TIMER else None
        with _on(self.device):
            check(lib().dps_operator_adjoint(self._h, r.data_ptr(), ptr(coef), msrc, ep, es, gp, gs, ptr(aux), n,
                                             stream_ptr(r.device)), f"dps_operator_adjoint[{self.kind}]")
        if tok:
            TIMER.stop(tok)
        return out


# ------------------------------------------------------------------------------------------------
# reductions / reweighting / resampling
# ------------------------------------------------------------------------------------------------
def particle_norms(partials, want_l1=False):
    n, P, _ = partials.shape
    l2 = torch.empty(n, device=partials.device, dtype=torch.float32)
    l1 = torch.empty(n, device=partials.device, dtype=torch.float32) if want_l1 else None
    check(lib().dps_particle_norms(partials.data_ptr(), P, n, l2.data_ptr(), ptr(l1), stream_ptr(partials.device)),
          "dps_particle_norms")
    return (l2, l1) if want_l1 else l2


def particle_sqdiff(a, ref, P: int = 32):
    """Per-particle (‖a − ref‖₂, ‖a − ref‖₁); ref is (1, …) (broadcast) or (n, …).  Two launches: partial sums, finish."""
    a = _lib.dense(a, "a")
    ref = _lib.dense(ref, "ref")
    n = a.shape[0]
    chw = a[0].numel()
    if ref.numel() not in (chw, n * chw):
        raise DpsError(f"reference of {ref.numel()} elements does not match particles of {chw}")
    ref_stride = 0 if ref.numel() == chw else chw
    partials = torch.empty((n, P, 2), device=a.device, dtype=torch.float32)
    check(lib().dps_particle_sqdiff(a.data_ptr(), chw, ref.data_ptr(), ref_stride, n, chw, partials.data_ptr(), P,
                                    stream_ptr(a.device)), "dps_particle_sqdiff")
    return particle_norms(partials, want_l1=True)


def guidance_coef(partials, mode: int, scale: float):
    """(‖r‖ per particle, coefficient folded into the adjoint): −scale/‖r‖ (mode 1) or −2·scale (mode 2)."""
    n, P, _ = partials.shape
    l2 = torch.empty(n, device=partials.device, dtype=torch.float32)
    coef = torch.empty(n, device=partials.device, dtype=torch.float32)
    tok = TIMER.start("guidance_coef") if TIMER else None
    check(lib().dps_guidance_coef(partials.data_ptr(), P, n, int(mode), float(scale), l2.data_ptr(), coef.data_ptr(),
                                  stream_ptr(partials.device)), "dps_guidance_coef")
    if tok:
        TIMER.stop(tok)
    return l2, coef


def particle_logweights(meas, sem=None, tau=1.0, meas_scale=1.0, meas_pow=1, sem_scale=0.0, sem_pow=1):
    meas = _lib.dense(meas, "meas")
    if sem is not None:
        sem = _lib.dense(sem, "sem")
    logw = torch.empty_like(meas)
    check(lib().dps_particle_logweights(meas.data_ptr(), ptr(sem), meas.numel(), float(tau), float(meas_scale),
                                        int(meas_pow), float(sem_scale), int(sem_pow), logw.data_ptr(),
                                        stream_ptr(meas.device)), "dps_particle_logweights")
    return logw


def weights_cdf(logw, linear_mode=False):
    """→ (normalised weights fp32, cdf fp32, lse fp32[1], degenerate int32[1])"""
    logw = _lib.dense(logw, "logw")
    n, dev = logw.numel(), logw.device
    w = torch.empty(n, device=dev, dtype=torch.float32)
    cdf = torch.empty(n, device=dev, dtype=torch.float32)
    lse = torch.empty(1, device=dev, dtype=torch.float32)
    deg = torch.empty(1, device=dev, dtype=torch.int32)
    check(lib().dps_weights_cdf(logw.data_ptr(), n, int(bool(linear_mode)), w.data_ptr(), cdf.data_ptr(),
                                lse.data_ptr(), deg.data_ptr(), stream_ptr(dev)), "dps_weights_cdf")
    return w, cdf, lse, deg


def ancestors(cdf, uniforms, n_draws, systematic=False, degenerate=None):
    """Ancestor indices (int64).  uniforms: fp64 CUDA tensor, n_draws values (multinomial) or 1 (systematic)."""
    if uniforms.dtype != torch.float64 or not uniforms.is_cuda:
        raise DpsError("uniforms must be a float64 CUDA tensor")
    need = 1 if systematic else n_draws
    if uniforms.numel() < need:
        raise DpsError(f"need {need} uniforms, got {uniforms.numel()}")
    out = torch.empty(n_draws, device=cdf.device, dtype=torch.int64)
    fn = lib().dps_ancestors_systematic if systematic else lib().dps_ancestors_multinomial
    check(fn(cdf.data_ptr(), cdf.numel(), uniforms.data_ptr(), n_draws, ptr(degenerate), out.data_ptr(),
             stream_ptr(cdf.device)), "dps_ancestors")
    return out


def gather_particles(src, ancestors_idx, out=None):
    """out[i] = src[ancestors[i]]  (img[ids], gaussian_diffusion.py:550, :697)"""
    src = _lib.dense(src, "src")
    n_dst = ancestors_idx.numel()
    elems = src[0].numel()
    out = torch.empty((n_dst,) + tuple(src.shape[1:]), device=src.device, dtype=torch.float32) if out is None else out
    check(lib().dps_gather_particles(src.data_ptr(), ancestors_idx.data_ptr(), out.data_ptr(), n_dst, elems,
                                     stream_ptr(src.device)), "dps_gather_particles")
    return out


def gather_particles_p2p(peer_ptrs_dev: int, n_per_rank: int, ancestors_idx, like, out=None):
    """out[i] = particle ancestors[i] read from its owner rank's symmetric buffer over NVLink.
    peer_ptrs_dev: device address of the array of per-rank buffer pointers (symmetric memory `buffer_ptrs_dev`)."""
    n_dst = ancestors_idx.numel()
    elems = like[0].numel()
    out = torch.empty((n_dst,) + tuple(like.shape[1:]), device=like.device, dtype=torch.float32) if out is None else out
    check(lib().dps_gather_particles_p2p(peer_ptrs_dev, int(n_per_rank), ancestors_idx.data_ptr(), out.data_ptr(), n_dst,
                                         elems, stream_ptr(like.device)), "dps_gather_particles_p2p")
    return out


def exchange_particles_p2p(peer_ptrs_dev: int, signal_pads_dev: int, rank: int, world: int, epoch: int, slot_elems: int,
                           n_per_rank: int, ancestors_idx, like, out=None):
    """gather_particles_p2p with the inter-GPU rendezvous inside the kernel (dps_exchange_particles_p2p): no barrier
    launches around it; reads the half of the double-buffered symmetric particle buffers that starts at `slot_elems`."""
    n_dst = ancestors_idx.numel()
    elems = like[0].numel()
    out = torch.empty((n_dst,) + tuple(like.shape[1:]), device=like.device, dtype=torch.float32) if out is None else out
    check(lib().dps_exchange_particles_p2p(peer_ptrs_dev, signal_pads_dev, int(rank), int(world), int(epoch) & 0xFFFFFFFF,
                                           int(slot_elems), int(n_per_rank), ancestors_idx.data_ptr(), out.data_ptr(),
                                           n_dst, elems, stream_ptr(like.device)), "dps_exchange_particles_p2p")
    return out


def argmin(costs):
    costs = _lib.dense(costs, "costs")
    best = torch.empty(1, device=costs.device, dtype=torch.int64)
    best_cost = torch.empty(1, device=costs.device, dtype=torch.float32)
    check(lib().dps_argmin(costs.data_ptr(), costs.numel(), best.data_ptr(), best_cost.data_ptr(),
                           stream_ptr(costs.device)), "dps_argmin")
    return best, best_cost


def broadcast_particle(src, index, n_dst, out=None):
    """out[i] = src[index] for all i  (img[best_path.repeat(n_paths)], gaussian_diffusion.py:633)"""
    src = _lib.dense(src, "src")
    elems = src[0].numel()
    out = torch.empty((n_dst,) + tuple(src.shape[1:]), device=src.device, dtype=torch.float32) if out is None else out
    check(lib().dps_broadcast_particle(src.data_ptr(), index.data_ptr(), out.data_ptr(), n_dst, elems,
                                       stream_ptr(src.device)), "dps_broadcast_particle")
    return out
