"""Measurement operators and noise models with the reference's names, constructor arguments and
duck type (guided_diffusion/measurements.py), running on the libdpsttc kernels.

Two ways in:
  * the reference's interface — `op.forward(data, **kwargs)` is differentiable under torch.autograd
    (forward kernel / adjoint kernel), ignores unknown kwargs, and `transpose`, `ortho_project`,
    `project`, `get_kernel`, `set_kernel` behave as in measurements.py:35-54, :76-189;
  * the fused interface the B200 samplers use — `residual()` (y − A(x̂₀) with x̂₀ formed on the fly
    from x and ε, plus the per-particle partial sums) and `cotangent()` (coef·Aᵀr ⊙ clamp mask).
"""
from __future__ import annotations

import zlib

import numpy as np
import torch

from . import kernels, tables
from ._lib import DPS_COEF_NORM, DPS_COEF_NORM_SQ, DpsError
from .kernels import OperatorPlan
from .registry import register_noise, register_operator


class _ForwardFn(torch.autograd.Function):
    """y = A(x): forward kernel; backward = adjoint kernel (Jᵀ for phase retrieval)."""

    @staticmethod
    def forward(ctx, data, plan):
        out, _, aux = plan.forward(data.detach())
        ctx.plan, ctx.aux = plan, aux
        return out

    @staticmethod
    def backward(ctx, grad_out):
        return ctx.plan.adjoint(grad_out.contiguous(), aux=ctx.aux), None


class _ResidualNormFn(torch.autograd.Function):
    """‖y − A(x)‖₂ per particle in one forward launch; backward is one adjoint launch with the
    per-particle factor −ḡ/‖r‖ folded in (torch.linalg.norm backward, 0 at r = 0)."""

    @staticmethod
    def forward(ctx, data, y, plan):
        r, partials, aux = plan.forward(data.detach(), y=y, want_partials=True)
        norm = kernels.particle_norms(partials)
        ctx.plan, ctx.aux = plan, aux
        ctx.save_for_backward(r, norm)
        return norm

    @staticmethod
    def backward(ctx, grad_norm):
        r, norm = ctx.saved_tensors
        coef = torch.where(norm > 0, -grad_norm / norm, torch.zeros_like(norm)).contiguous()
        return ctx.plan.adjoint(r, coef=coef, aux=ctx.aux), None, None


class B200Operator:
    """Common machinery: lazily built plan per input shape, autograd bridge, fused entry points."""
    name = "b200"
    linear = True

    def __init__(self, device):
        self.device = torch.device(device)
        self._plans = {}

    # -- subclass hook ----------------------------------------------------------------------------
    def _build_plan(self, C, H, W, **kwargs) -> OperatorPlan:
        raise NotImplementedError

    def _plan_key(self, shape, kwargs):
        return tuple(shape[1:])

    def plan_for(self, data, **kwargs) -> OperatorPlan:
        if not data.is_cuda:
            raise DpsError(f"{type(self).__name__}: data is on {data.device}; the B200 operators have no CPU path")
        key = (data.device.index, self._plan_key(data.shape, kwargs))    # plans own device tables: one per device
        plan = self._plans.get(key)
        if plan is None:
            _, C, H, W = data.shape
            plan = self._plans[key] = self._build_plan(C, H, W, device=data.device, **kwargs)
        return plan

    # -- reference interface ----------------------------------------------------------------------
    def forward(self, data, **kwargs):
        plan = self.plan_for(data, **kwargs)
        data = data if data.dtype == torch.float32 else data.float()
        if not data.is_contiguous():
            data = data.contiguous()
        return _ForwardFn.apply(data, plan)

    def transpose(self, data, **kwargs):  # the reference's blur/inpainting "transpose" is the identity
        return data

    @staticmethod
    def _no_grad_needed(*tensors):  # the project kernels carry no autograd graph
        return not (torch.is_grad_enabled() and any(t.requires_grad for t in tensors))

    def _fused_project(self, data, measurement, **kwargs):
        """dps_operator_project: the elementwise part of project / ortho_project in the operator kernels' epilogues."""
        plan = self.plan_for(data, **kwargs)
        d = data.detach().float().contiguous()
        y = None if measurement is None else measurement.detach().to(d.device, torch.float32)
        return plan.project(d, y)

    def ortho_project(self, data, **kwargs):  # (I − AᵀA)x with the reference's transpose   measurements.py:48-50
        if self.linear and self._no_grad_needed(data):
            return self._fused_project(data, None, **kwargs)
        return data - self.transpose(self.forward(data, **kwargs), **kwargs)

    def project(self, data, measurement, **kwargs):  # measurements.py:52-54
        if self.linear and self._no_grad_needed(data, measurement):
            return self._fused_project(data, measurement, **kwargs)
        return self.ortho_project(measurement, **kwargs) - self.forward(data, **kwargs)

    def residual_norm(self, data, measurement, **kwargs):
        """Differentiable per-particle ‖measurement − A(data)‖₂ (fused; condition_methods.py:36-39)."""
        plan = self.plan_for(data, **kwargs)
        data = data.float().contiguous()
        return _ResidualNormFn.apply(data, measurement.to(data.device, torch.float32), plan)

    # -- fused interface --------------------------------------------------------------------------
    def residual(self, x, eps=None, k=None, clip=True, y=None, want_partials=True, out=None, aux=None, **kwargs):
        """(r = y − A(x̂₀), partials, aux); x̂₀ = clamp(c1·x − c2·ε) on the fly (x̂₀ = x when eps is None)."""
        return self.plan_for(x, **kwargs).forward(x, eps, k, clip, y, want_partials, aux, out)

    def guidance(self, x, eps, k, clip, y, out, want_r=False, **kwargs):
        """(partials, r or None, aux) with out ← clamp-mask ⊙ Aᵀ(y − A x̂₀), UNSCALED: the per-particle coefficient is applied
        by the posterior-update kernel (kernels.posterior_update(deferred=…)).  One fused kernel where the plan has one."""
        return self.plan_for(x, **kwargs).guidance(x, eps, k, clip, y, out, want_r=want_r)

    def cotangent(self, r, coef, x, eps=None, k=None, clip=True, extra=None, out=None, aux=None, **kwargs):
        """clamp-mask ⊙ (coef_n·Aᵀ r + extra) — the cotangent w.r.t. the pre-clamp x̂₀ (App. A.4)."""
        if x is None:
            raise DpsError("cotangent(): pass the particle tensor x (plan lookup and clamp mask)")
        return self.plan_for(x, **kwargs).adjoint(r, coef, x, eps, k, clip, extra, out, aux)


@register_operator(name="noise")
class DenoiseOperator:
    """Identity (measurements.py:57-73).  No kernel: nothing to compute."""
    linear = True

    def __init__(self, device):
        self.name = "noise"
        self.device = device

    def forward(self, data, **kwargs):
        return data

    def transpose(self, data, **kwargs):
        return data

    def ortho_project(self, data, **kwargs):
        return data

    def project(self, data, *args, **kwargs):
        return data


@register_operator(name="super_resolution")
class SuperResolutionOperator(B200Operator):
    """measurements.py:76-91: antialiased bicubic ↓scale_factor via the Resizer bands."""

    def __init__(self, in_shape, scale_factor, device):
        super().__init__(device)
        self.name = "super_resolution"
        self.in_shape = tuple(in_shape)
        self.scale_factor = scale_factor
        (self.fov_h, self.w_h), (self.fov_w, self.w_w), self.out_hw = tables.resizer_tables(self.in_shape, 1 / scale_factor)

    def _build_plan(self, C, H, W, device, **kwargs):
        if (H, W) != self.in_shape[-2:]:
            raise DpsError(f"super_resolution built for {self.in_shape[-2:]}, got {(H, W)}")
        return OperatorPlan.resize(self.fov_h, self.w_h, self.fov_w, self.w_w, C, H, W, device)

    def transpose(self, data, **kwargs):  # nearest up-sampling, as the reference (F.interpolate default)
        return torch.nn.functional.interpolate(data, scale_factor=self.scale_factor)

    def project(self, data, measurement, **kwargs):  # measurements.py:90-91
        if self._no_grad_needed(data, measurement):
            return self._fused_project(data, measurement)       # one cluster-kernel launch at 256², ×4 / ×8
        return data - self.transpose(self.forward(data)) + self.transpose(measurement)


class _BlurOperator(B200Operator):
    kernel_size = 0

    def _kernel_fp32(self) -> np.ndarray:
        raise NotImplementedError

    def _build_plan(self, C, H, W, device, **kwargs):
        return OperatorPlan.blur(self._kernel_fp32(), C, H, W, device)


@register_operator(name="gaussian_blur")
class GaussialBlurOperator(_BlurOperator):
    """measurements.py:129-149 (class name spelled as in the reference)."""

    def __init__(self, kernel_size, intensity, device):
        super().__init__(device)
        self.name = "gaussian_blur"
        self.kernel_size = kernel_size
        self.kernel = torch.from_numpy(tables.gaussian_kernel(kernel_size, intensity))  # fp64, like Blurkernel.k

    def _kernel_fp32(self):
        return self.kernel.to(torch.float32).numpy()

    def get_kernel(self):
        return self.kernel.view(1, 1, self.kernel_size, self.kernel_size)


@register_operator(name="motion_blur")
class MotionBlurOperator(_BlurOperator):
    """measurements.py:93-126.  The kernel generator of the reference is the external `motionblur`
    package; if it is importable it is used, else tables.MotionKernel (same interface)."""

    def __init__(self, kernel_size, intensity, device):
        super().__init__(device)
        self.name = "motion_blur"
        self.kernel_size = kernel_size
        try:
            from motionblur.motionblur import Kernel
        except Exception:  # noqa: BLE001
            Kernel = tables.MotionKernel
        self.kernel = Kernel(size=(kernel_size, kernel_size), intensity=intensity)
        self._weights = np.asarray(self.kernel.kernelMatrix, dtype=np.float32)

    def _kernel_fp32(self):
        return self._weights

    def get_kernel(self):
        k = torch.from_numpy(np.asarray(self.kernel.kernelMatrix)).type(torch.float32).to(self.device)
        return k.view(1, 1, self.kernel_size, self.kernel_size)

    def set_kernel(self, kernel):
        """measurements.py:119-126 — note the transpose the reference applies."""
        self._weights = np.ascontiguousarray(np.asarray(kernel, dtype=np.float32).T)
        self._plans.clear()


@register_operator(name="inpainting")
class InpaintingOperator(B200Operator):
    """measurements.py:151-168: data * mask; the mask arrives as a kwarg on every call."""

    _MAX_MASKS = 4     # plans kept alive (one cudaMalloc'd mask each); the drivers use one mask per image

    def __init__(self, device):
        super().__init__(device)
        self.name = "inpainting"
        self._mask_cache = {}   # key -> (strong reference to the caller's mask, its _version / content hash, host copy)

    def _mask_host(self, mask, H, W):
        """(cache key, host fp32 (H, W) copy) of a mask.  A cache entry keeps a STRONG reference to the caller's mask
        object, so its id / address cannot be recycled for another mask while the entry lives (the drivers build a new
        mask per image: partial(cond.conditioning, mask=mask), sample_condition_*.py); a hit additionally requires the
        same object and an unchanged tensor version (numpy: an unchanged crc32 of the bytes).  The cache is a small LRU —
        evicted entries drop their plan and its device mask."""
        if mask is None:
            raise ValueError("Require mask")
        is_t = torch.is_tensor(mask)
        key = (id(mask), tuple(mask.shape), H, W)
        stamp = mask._version if is_t else zlib.crc32(np.ascontiguousarray(mask).view(np.uint8).reshape(-1))
        hit = self._mask_cache.get(key)
        if hit is not None and hit[0] is mask and hit[1] == stamp:
            self._mask_cache[key] = self._mask_cache.pop(key)          # most recently used last
            return key, hit[2]
        m = mask.detach().to("cpu", torch.float32).numpy() if is_t else np.asarray(mask, np.float32)
        if m.size != H * W:
            m = np.broadcast_to(m, (1, m.shape[-3] if m.ndim >= 3 else 1, H, W))
            if not (m == m[:, :1]).all():
                raise DpsError("inpainting: per-channel / per-particle masks are not supported by the kernel")
            m = m[0, 0]
        host = np.ascontiguousarray(m.reshape(H, W)).copy()
        self._mask_cache.pop(key, None)
        self._drop_plans(key)                                            # same object, new contents: its plan is stale
        self._mask_cache[key] = (mask, stamp, host)
        while len(self._mask_cache) > self._MAX_MASKS:
            old = next(iter(self._mask_cache))
            del self._mask_cache[old]
            self._drop_plans(old)
        return key, host

    def _drop_plans(self, mask_key):
        for k in [k for k in self._plans if k[1][-1] == mask_key]:
            del self._plans[k]

    def _plan_key(self, shape, kwargs):
        mask = kwargs.get("mask")
        if mask is None:
            raise ValueError("Require mask")
        key, _ = self._mask_host(mask, shape[-2], shape[-1])
        return (tuple(shape[1:]), key)

    def _build_plan(self, C, H, W, device, **kwargs):
        _, m = self._mask_host(kwargs.get("mask"), H, W)
        return OperatorPlan.inpainting(m, C, H, W, device)

    def forward(self, data, **kwargs):
        if kwargs.get("mask", None) is None:
            raise ValueError("Require mask")  # measurements.py:159-162
        return super().forward(data, **kwargs)

    def ortho_project(self, data, **kwargs):  # measurements.py:167-168
        if self._no_grad_needed(data):
            return self._fused_project(data, None, **kwargs)
        return data - self.forward(data, **kwargs)


@register_operator(name="phase_retrieval")
class PhaseRetrievalOperator(B200Operator):
    """measurements.py:179-189: |centred ortho FFT2 of the zero-padded image|.  Nonlinear."""
    linear = False

    def __init__(self, oversample, device):
        super().__init__(device)
        self.pad = int((oversample / 8.0) * 256)
        self.name = "phase_retrieval"

    def _build_plan(self, C, H, W, device, **kwargs):
        return OperatorPlan.phase(self.pad, C, H, W, device)

    def project(self, data, measurement, **kwargs):  # NonLinearOperator.project, measurements.py:175-177
        return data + measurement - self.forward(data)


@register_operator(name="nonlinear_blur")
class NonlinearBlurOperator:
    """measurements.py:191-218: the bkse KernelWizard CNN — an un-vendored external network with a fresh
    random latent per call.  It cannot be reduced to a convolution; it stays a PyTorch module behind the
    same interface (boundary only)."""
    linear = False

    def __init__(self, opt_yml_path, device):
        self.name = "nonlinear_blur"
        self.device = device
        try:
            import yaml
            from bkse.models.kernel_encoding.kernel_wizard import KernelWizard
        except Exception as e:  # noqa: BLE001
            raise RuntimeError("nonlinear_blur needs the external `bkse` package and its pretrained weights") from e
        with open(opt_yml_path) as f:
            opt = yaml.safe_load(f)["KernelWizard"]
        self.blur_model = KernelWizard(opt)
        self.blur_model.eval()
        self.blur_model.load_state_dict(torch.load(opt["pretrained"]))
        self.blur_model = self.blur_model.to(device)

    def forward(self, data, **kwargs):
        random_kernel = torch.randn(1, 512, 2, 2).to(self.device) * 1.2
        blurred = self.blur_model.adaptKernel((data + 1.0) / 2.0, kernel=random_kernel)
        return (blurred * 2.0 - 1.0).clamp(-1, 1)

    def project(self, data, measurement, **kwargs):
        return data + measurement - self.forward(data)


# ---------------------------------------------------------------------------------------------
# noise models (measurements.py:243-285) — set-up only, plain torch/numpy like the reference
# ---------------------------------------------------------------------------------------------
class _Noise:
    def __call__(self, data):
        return self.forward(data)


@register_noise(name="clean")
class Clean(_Noise):
    def forward(self, data):
        return data


@register_noise(name="gaussian")
class GaussianNoise(_Noise):
    def __init__(self, sigma):
        self.sigma = sigma

    def forward(self, data):
        return data + torch.randn_like(data, device=data.device) * self.sigma


@register_noise(name="poisson")
class PoissonNoise(_Noise):
    def __init__(self, rate):
        self.rate = rate

    def forward(self, data):
        x = ((data + 1.0) / 2.0).clamp(0, 1).detach().cpu()
        x = torch.from_numpy(np.random.poisson(x * 255.0 * self.rate) / 255.0 / self.rate)
        return (x * 2.0 - 1.0).clamp(-1, 1).to(data.device)


GUIDANCE_COEF = {"norm": DPS_COEF_NORM, "norm_sq": DPS_COEF_NORM_SQ}
