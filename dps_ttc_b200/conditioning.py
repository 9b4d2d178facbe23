"""Conditioning methods with the reference's names, constructor arguments and return arities
(guided_diffusion/condition_methods.py).  Each class works two ways:

  * `conditioning(x_prev, x_t, x_0_hat, measurement, **kw)` — the reference's call, differentiated by
    torch.autograd through whatever produced x_0_hat (the UNet); the residual norm and its backward run
    as ONE forward and ONE adjoint kernel (operators._ResidualNormFn) instead of the ATen chain;
  * `guidance(step)` — a declarative GuidanceSpec the fused samplers (sampler.py) consume, so that the
    whole step runs as three kernels around the UNet forward/VJP.

HEAD of the reference is internally inconsistent about return arities; the classes
return exactly what HEAD returns, and the fused samplers implement the intended update
(x' = sample − ζ·∇) for all of them.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Callable, Optional

import torch

from ._lib import DPS_COEF_NORM, DPS_COEF_NORM_SQ
from .operators import B200Operator
from .registry import register_conditioning_method
from .schedule import semantic_scale


@dataclass
class GuidanceSpec:
    """What the fused step needs to know about a conditioning method."""
    kind: str                     # 'none' | 'ps' | 'ps_anneal' | 'ps_semantic' | 'mcg'
    coef_mode: int = DPS_COEF_NORM  # ∇‖r‖ or ∇‖r‖²
    scale: float = 0.0            # ζ multiplying the measurement term at this step
    semantic: Optional[Callable] = None  # x0 (N,C,H,W, requires grad) -> (sem_loss (N,), sem_dist (N,)) or None
    project: bool = False         # mcg: follow with operator.project(x_t, noisy_measurement)
    returns: str = "x_t"          # what HEAD's conditioning() returns first: 'x_t' or 'grad'


class ConditioningMethod:
    def __init__(self, operator, noiser, **kwargs):
        self.operator = operator
        self.noiser = noiser
        self.l1 = kwargs.get("l1", 0.0)

    def project(self, data, noisy_measurement, **kwargs):
        return self.operator.project(data=data, measurement=noisy_measurement, **kwargs)

    # ‖y − A(x̂₀)‖₂ per particle, differentiable w.r.t. x_0_hat
    def _distance(self, x_0_hat, measurement, **kwargs):
        if isinstance(self.operator, B200Operator) and x_0_hat.is_cuda:
            return self.operator.residual_norm(x_0_hat, measurement, **kwargs)
        diff = measurement - self.operator.forward(x_0_hat, **kwargs)
        return torch.linalg.norm(diff.reshape(diff.shape[0], -1), dim=-1)

    def grad_and_value(self, x_prev, x_0_hat, measurement, **kwargs):
        """condition_methods.py:33-60."""
        name = self.noiser.__name__
        if name == "gaussian":
            norm = self._distance(x_0_hat, measurement, **kwargs)
            power = norm ** 2 if kwargs.get("norm_exp", 1) == 2 else norm
            norm_grad = torch.autograd.grad(outputs=power.sum(), inputs=x_prev)[0]
        elif name == "poisson":
            diff = measurement - self.operator.forward(x_0_hat, **kwargs)
            norm = (torch.linalg.norm(diff) / measurement.abs()).mean()
            norm_grad = torch.autograd.grad(outputs=norm, inputs=x_prev)[0]
        else:
            raise NotImplementedError
        return norm_grad, norm

    def conditioning(self, x_t, measurement, noisy_measurement=None, **kwargs):
        raise NotImplementedError

    def guidance(self, beta_scale: float = 0.0, t: float = 1.0, anneal: float = 1.0) -> GuidanceSpec:
        return GuidanceSpec(kind="none")


@register_conditioning_method(name="vanilla")
class Identity(ConditioningMethod):
    def conditioning(self, x_t, *args, **kwargs):
        return x_t


@register_conditioning_method(name="projection")
class Projection(ConditioningMethod):
    def conditioning(self, x_t, noisy_measurement, **kwargs):
        return self.project(data=x_t, noisy_measurement=noisy_measurement)


@register_conditioning_method(name="mcg")
class ManifoldConstraintGradient(ConditioningMethod):
    def __init__(self, operator, noiser, **kwargs):
        super().__init__(operator, noiser)
        self.scale = kwargs.get("scale", 1.0)

    def conditioning(self, x_prev, x_t, x_0_hat, measurement, noisy_measurement, **kwargs):
        norm_grad, norm = self.grad_and_value(x_prev=x_prev, x_0_hat=x_0_hat, measurement=measurement, **kwargs)
        x_t -= norm_grad * self.scale
        x_t = self.project(data=x_t, noisy_measurement=noisy_measurement, **kwargs)
        return x_t, norm

    def guidance(self, beta_scale=0.0, t=1.0, anneal=1.0):
        return GuidanceSpec(kind="mcg", scale=self.scale, project=True)


@register_conditioning_method(name="ps")
class PosteriorSampling(ConditioningMethod):
    def __init__(self, operator, noiser, **kwargs):
        super().__init__(operator, noiser)
        self.scale = kwargs.get("scale", 0.3)
        self.operator_name = operator.name

    def conditioning(self, x_prev, x_t, x_0_hat, measurement, **kwargs):
        norm_grad, norm = self.grad_and_value(x_prev=x_prev, x_0_hat=x_0_hat, measurement=measurement, **kwargs)
        x_t -= norm_grad * self.scale
        return x_t, norm, self.scale / 2 / norm

    def guidance(self, beta_scale=0.0, t=1.0, anneal=1.0):
        return GuidanceSpec(kind="ps", scale=self.scale)


@register_conditioning_method(name="ps_semantic")
class PosteriorSamplingSemanticGuid(ConditioningMethod):
    """condition_methods.py:110-195.  The face-embedding network is external (facenet_pytorch,
    pretrained weights from the network): pass `embedder` (any module mapping (N,3,H,W) → (N,D)) and
    `guid_emb` ((1,n_guid,D) or (n_guid,D)), or `guid_images` when facenet_pytorch is installed."""

    def __init__(self, operator, noiser, **kwargs):
        super().__init__(operator, noiser)
        self.operator_name = operator.name
        self.scale = kwargs.get("scale", 0.3)
        self.sem_guid_scale = kwargs.get("sem_guid_scale", 0.5)
        self.anneal_factor = kwargs.get("anneal_factor", 1.0)
        self.norm_exp = kwargs.get("norm_exp", 1)
        self.guid_images = kwargs.get("guid_images", None)
        self.resnet = kwargs.get("embedder", None)
        guid_emb = kwargs.get("guid_emb", None)
        self.n_guid_images = 1
        self.guid_image_emb = None
        if self.sem_guid_scale == 0 or (self.guid_images is None and guid_emb is None):
            self.resnet = None  # plain DPS that returns its gradient (condition_methods.py:147-150)
        elif guid_emb is not None:
            if self.resnet is None:
                raise ValueError("guid_emb given without an embedder")
            self.guid_image_emb = guid_emb if guid_emb.ndim == 3 else guid_emb.unsqueeze(0)
            self.n_guid_images = self.guid_image_emb.shape[1]
        else:
            from facenet_pytorch import MTCNN, InceptionResnetV1  # external, see class docstring
            device = "cuda:0"
            self.n_guid_images = len(self.guid_images)
            mtcnn = MTCNN(image_size=256, margin=10, min_face_size=20, device=device)
            self.resnet = InceptionResnetV1(pretrained="vggface2", device=device).eval()
            with torch.no_grad():
                crops = torch.stack(mtcnn(self.guid_images)).to(device)
                self.guid_image_emb = self.resnet(crops).unsqueeze(0)

    @property
    def semantic_enabled(self):
        return self.resnet is not None and self.sem_guid_scale != 0

    def _semantic(self, x_0_hat):
        """(Σ-able loss term before scaling, distance) per particle — :158-173."""
        emb = self.resnet(x_0_hat).unsqueeze(1)
        diff = (emb - self.guid_image_emb.to(emb.device)).reshape(emb.shape[0], -1)
        dist = torch.norm(diff, dim=-1) / self.n_guid_images
        return (dist ** 2 if self.norm_exp == 2 else dist), dist

    def measurement_semantic_guidance(self, x_prev, x_0_hat, measurement, **kwargs):
        if self.semantic_enabled:
            s_t = semantic_scale(kwargs.get("t", 1), self.sem_guid_scale, self.anneal_factor)
            sem_loss, sem_dist = self._semantic(x_0_hat)
        else:
            s_t, sem_dist = 0, torch.tensor(0.0).to(x_0_hat.device)
            sem_loss = sem_dist
        if self.noiser.__name__ != "gaussian":
            raise NotImplementedError
        meas = self._distance(x_0_hat, measurement, **kwargs)
        net_loss = self.scale * meas + s_t * sem_loss
        norm_grad = torch.autograd.grad(outputs=net_loss.sum(), inputs=x_prev)[0]
        return norm_grad, meas, sem_dist

    def conditioning(self, x_prev, x_t, x_0_hat, measurement, **kwargs):
        return self.measurement_semantic_guidance(x_prev=x_prev, x_0_hat=x_0_hat, measurement=measurement, **kwargs)

    def guidance(self, beta_scale=0.0, t=1.0, anneal=1.0):
        sem = None
        if self.semantic_enabled:
            s_t = semantic_scale(t, self.sem_guid_scale, self.anneal_factor)

            def sem(x0, s_t=s_t):
                loss, dist = self._semantic(x0)
                return s_t * loss, dist
        return GuidanceSpec(kind="ps_semantic", scale=self.scale, semantic=sem, returns="grad")


@register_conditioning_method(name="ps_anneal")
class PosterorSamplingAnnealing(ConditioningMethod):
    """condition_methods.py:198-212 (class name spelled as in the reference)."""

    def __init__(self, operator, noiser, **kwargs):
        super().__init__(operator, noiser)
        self.noise_sigma = max(noiser.sigma, 0.05)
        self.scale = kwargs.get("scale", 0.3)
        self.operator_name = operator.name

    def net_scaling(self, beta_scale, anneal):
        return beta_scale / (anneal * self.noise_sigma ** 2)

    def conditioning(self, x_prev, x_t, x_0_hat, measurement, **kwargs):
        beta_scale = kwargs.pop("beta_scale", self.scale)
        anneal = kwargs.pop("anneal", 1.0)
        net_scaling = torch.tensor(self.net_scaling(beta_scale, anneal)).to(x_t.device)
        kwargs.pop("norm_exp", None)
        grad, norm = self.grad_and_value(x_prev=x_prev, x_0_hat=x_0_hat, measurement=measurement, norm_exp=2, **kwargs)
        x_t -= net_scaling * grad
        return x_t, norm, net_scaling

    def guidance(self, beta_scale=None, t=1.0, anneal=1.0):
        beta_scale = self.scale if beta_scale is None else beta_scale
        return GuidanceSpec(kind="ps_anneal", coef_mode=DPS_COEF_NORM_SQ, scale=float(self.net_scaling(beta_scale, anneal)))


@register_conditioning_method(name="ps+")
class PosteriorSamplingPlus(ConditioningMethod):
    """condition_methods.py:215-232 (global norm over several noisy copies of x̂₀)."""

    def __init__(self, operator, noiser, **kwargs):
        super().__init__(operator, noiser)
        self.num_sampling = kwargs.get("num_sampling", 5)
        self.scale = kwargs.get("scale", 1.0)

    def conditioning(self, x_prev, x_t, x_0_hat, measurement, **kwargs):
        norm = 0
        for _ in range(self.num_sampling):
            noisy = x_0_hat + 0.05 * torch.rand_like(x_0_hat)
            norm += torch.linalg.norm(measurement - self.operator.forward(noisy)) / self.num_sampling
        norm_grad = torch.autograd.grad(outputs=norm, inputs=x_prev)[0]
        x_t -= norm_grad * self.scale
        return x_t, norm
