"""Samplers with the reference's names, constructor and `p_sample_loop` signatures
(guided_diffusion/gaussian_diffusion.py), re-designed around the fused kernels.

One guided step (ddpm/ddim + ps / ps_anneal / ps_semantic / mcg) is

    UNet forward (reference module, under autograd)                                   [torch]
    residual      r = y − A(clamp(c1·x − c2·ε)), per-CTA Σr²                         [kernel 1]
    coefficients  ‖r‖ₙ, −ζ/‖r‖ₙ                                                      [kernel 2, O(N)]
    cotangent     g = 1[|pre| ≤ 1] ⊙ (coefₙ·Aᵀr + semantic term) → ε-channels of G6  [kernel 3]
    UNet VJP      vjp = autograd.grad(model_out, x, G6)                               [torch]
    update        x' = μ(x̂₀,x) + σ(v)·z − (c1·g − c2·vjp)                            [kernel 4]

with no host synchronisation, no table upload and no intermediate particle tensor other than r, g and
vjp.  The reference does the same work in ~430-650 ATen calls with ≥3 device→host syncs per step
(SURVEY §3, App. A.6).

Per-step RNG draw order on the particle device is the reference's: z ~ randn_like(x), then
randn_like(measurement) for q_sample, then the resampling uniforms.
"""
from __future__ import annotations

import functools

import torch

from . import diffstategrad, kernels
from ._lib import DPS_COEF_GLOBAL_NORM, DPS_COEF_NORM, DPS_COEF_NORM_SQ, DpsError
from .conditioning import ConditioningMethod, GuidanceSpec
from .graphed import GraphedEps
from .operators import B200Operator
from .registry import get_sampler, register_sampler
from .schedule import Schedule, anneal_factor, named_beta_schedule, space_timesteps


def create_sampler(sampler, steps, noise_schedule, model_mean_type, model_var_type, dynamic_threshold,
                   clip_denoised, rescale_timesteps, timestep_respacing=""):
    """gaussian_diffusion.py:34-56, same arguments (the YAML of configs/diffusion_config.yaml)."""
    cls = get_sampler(name=sampler)
    betas = named_beta_schedule(noise_schedule, steps)
    respacing = timestep_respacing if timestep_respacing else [steps]
    return cls(use_timesteps=space_timesteps(steps, respacing), betas=betas, model_mean_type=model_mean_type,
               model_var_type=model_var_type, dynamic_threshold=dynamic_threshold, clip_denoised=clip_denoised,
               rescale_timesteps=rescale_timesteps)


# ------------------------------------------------------------------------------------------------
# noise sources
# ------------------------------------------------------------------------------------------------
class TorchNoise:
    """Draws from torch's global generators exactly where the reference does."""

    def z(self, idx, like):
        return torch.randn_like(like)

    def q(self, idx, like):
        return torch.randn_like(like)

    def uniforms(self, idx, n, device):
        # torch.multinomial on CPU consumes n fp64 uniforms of the global CPU generator (pinned in tests)
        return torch.rand(n, dtype=torch.float64).to(device, non_blocking=True)


class PhiloxNoise:
    """Throughput mode: σ·z is generated INSIDE the posterior-update kernel from the counter-based Philox4x32-10 generator
    keyed by (seed, step, global particle index, element) — no z tensor is written or read (7T → 6T bytes for the DDPM
    update) and a sharded run draws the same noise as an unsharded one.  It does not reproduce torch's RNG stream: parity
    runs use TorchNoise or a NoiseTape.  Draws that are not the update's own z (q_sample, resampling uniforms) come from
    torch generators seeded with the same seed."""

    def __init__(self, seed=0):
        self.seed = int(seed)
        self._gen = {}

    def _g(self, device):
        key = str(device)
        if key not in self._gen:
            self._gen[key] = torch.Generator(device).manual_seed(self.seed)
        return self._gen[key]

    def z(self, idx, like):
        return None                      # made in the kernel

    def q(self, idx, like):
        return torch.randn(like.shape, device=like.device, dtype=like.dtype, generator=self._g(like.device))

    def uniforms(self, idx, n, device):
        g = torch.Generator().manual_seed((self.seed * 1_000_003 + int(idx)) & 0x7FFFFFFFFFFF)
        return torch.rand(n, dtype=torch.float64, generator=g).to(device, non_blocking=True)


class NoiseTape:
    """Pre-recorded draws (e.g. captured from a reference run, or produced on the host): per step a
    pinned host tensor that is copied to the device inside the step — the e2e / parity path."""

    def __init__(self, z=None, q=None, uniforms=None, pin=True):
        def prep(d):
            if d is None:
                return {}
            out = {}
            for k, t in d.items():
                t = t.contiguous()
                out[k] = t.pin_memory() if (pin and torch.cuda.is_available() and not t.is_cuda) else t
            return out
        self._z, self._q, self._u = prep(z), prep(q), prep(uniforms)
        self.h2d_bytes = 0

    def _get(self, table, idx, like_shape, device, dtype):
        if idx not in table:
            raise DpsError(f"noise tape has no entry for step {idx}")
        t = table[idx]
        self.h2d_bytes += t.numel() * t.element_size()
        return t.to(device, non_blocking=True).to(dtype).reshape(like_shape)

    def z(self, idx, like):
        return self._get(self._z, idx, like.shape, like.device, like.dtype)

    def q(self, idx, like):
        if idx not in self._q:
            return None
        return self._get(self._q, idx, like.shape, like.device, like.dtype)

    def uniforms(self, idx, n, device):
        return self._get(self._u, idx, (n,), device, torch.float64)


def _returns_gradient(method) -> bool:
    """Does this conditioning object's conditioning() return the GRADIENT first (HEAD's ps_semantic, condition_methods.py:
    145-195) rather than the updated x_t?  Ours say so in their GuidanceSpec; the reference's class is recognised by
    identity when its module is loaded (by name only as a last resort, e.g. a subclass defined by a driver)."""
    import sys
    if isinstance(method, ConditioningMethod):
        return method.guidance().returns == "grad"
    ref = sys.modules.get("guided_diffusion.condition_methods")
    ref_cls = getattr(ref, "PosteriorSamplingSemanticGuid", None) if ref is not None else None
    if ref_cls is not None and isinstance(method, ref_cls):
        return True
    return any(c.__name__ == "PosteriorSamplingSemanticGuid" for c in type(method).__mro__)


def _resolve_cond_fn(fn):
    """Unwrap functools.partial layers (the drivers bind mask=…/l1=… this way) down to the bound method."""
    bound = {}
    while isinstance(fn, functools.partial):
        bound = {**fn.keywords, **bound}
        fn = fn.func
    return getattr(fn, "__self__", None), fn, bound


# ------------------------------------------------------------------------------------------------
# base chain
# ------------------------------------------------------------------------------------------------
class SpacedSampler:
    """GaussianDiffusion + SpacedDiffusion (gaussian_diffusion.py:59-117, :395-463) on fp64 host tables."""
    kind = "ddpm"
    _VAR_MODES = {"learned_range": 0, "fixed_small": 1, "fixed_large": 1, "learned": 2}

    def __init__(self, use_timesteps, betas, model_mean_type, model_var_type, dynamic_threshold, clip_denoised,
                 rescale_timesteps):
        if model_mean_type not in ("epsilon", "start_x", "previous_x"):
            raise NameError(f"Name {model_mean_type} is not defined.")      # get_mean_processor's error
        if model_mean_type != "epsilon" and self.kind == "ddim":
            raise NotImplementedError("DDIM recovers ε with the epsilon processor's tables (gaussian_diffusion.py:506-509); "
                                      f"model_mean_type={model_mean_type!r} is covered for the DDPM family only")
        self.model_mean_type = model_mean_type
        if model_var_type not in self._VAR_MODES:
            raise NameError(f"Name {model_var_type} is not defined.")
        if dynamic_threshold and model_mean_type != "epsilon":
            raise NotImplementedError("dynamic_threshold is covered for model_mean_type='epsilon' (the configured processor)")
        self.dynamic_threshold = bool(dynamic_threshold)
        self.schedule = Schedule(betas, use_timesteps, rescale_timesteps)
        s = self.schedule
        self.betas, self.num_timesteps, self.timestep_map = s.betas, s.num_timesteps, s.timestep_map
        self.original_num_steps, self.rescale_timesteps = s.original_num_steps, s.rescale_timesteps
        for name in ("alphas_cumprod", "alphas_cumprod_prev", "sqrt_alphas_cumprod", "sqrt_one_minus_alphas_cumprod",
                     "sqrt_recip_alphas_cumprod", "sqrt_recipm1_alphas_cumprod", "posterior_variance",
                     "posterior_log_variance_clipped", "posterior_mean_coef1", "posterior_mean_coef2"):
            setattr(self, name, getattr(s, name))
        self.model_var_type = model_var_type
        self.var_mode = self._VAR_MODES[model_var_type]
        self.clip_denoised = bool(clip_denoised)
        self.eta = 0.0
        self.noise = TorchNoise()
        self.parity_rng = True      # also draw the (unused) q_sample noise, keeping the reference's RNG stream
        self.last_stats = {}

    # -- helpers ----------------------------------------------------------------------------------
    def _idx(self, t) -> int:
        return int(t.reshape(-1)[0]) if torch.is_tensor(t) else int(t)

    def _consts(self, idx: int):
        return self.schedule.consts(idx, self.eta, self.model_mean_type)

    def _max_log(self, k):
        if self.model_var_type == "fixed_small":
            return k.fixed_small_log
        if self.model_var_type == "fixed_large":
            return k.fixed_large_log
        return None

    def _model_out(self, model, x, k):
        t = torch.full((1,), k.model_t, device=x.device, dtype=torch.float32)
        out = model(x, t)
        if not out.is_contiguous():  # e.g. a channels_last module: the kernels read dense NCHW particle planes
            out = out.contiguous()
        C = x.shape[1]
        if out.shape[1] == 2 * C:
            return out, out[:, :C], out[:, C:]
        if self.var_mode in (0, 2):
            raise DpsError(f"model_var_type={self.model_var_type} needs a model with 2·C output channels")
        return out, out, None

    # -- dynamic thresholding (process_xstart, posterior_mean_variance.py:40-45; util/img_utils.py:237-249) -----------
    # x̂₀ = clip(pre·quantile(|pre|, 0.95)) over the WHOLE particle batch, differentiated through the quantile.  At t ≈ T the
    # product pre·s is ~10⁴ before clipping, so x̂₀ only matches the reference to 1e-4 if it is rounded exactly where the
    # reference rounds it (folding s into c1, c2 — which the fused kernels would need — is off by ~3e-3).  Samplers with
    # dynamic_threshold=True therefore take the generic autograd path: this torch restatement for x̂₀ / mean / sample, the
    # B200 operator kernels (forward + adjoint through autograd) for the guidance.
    def _process_xstart(self, pre):
        """torch restatement (differentiable, like the reference: autograd flows through the quantile too)."""
        if self.dynamic_threshold:
            pre = torch.clip(pre * torch.quantile(pre.abs(), 0.95), -1.0, 1.0)
        return pre.clamp(-1, 1) if self.clip_denoised else pre

    def _needs_z(self, k):
        return bool(k.noise_on and (self.kind == "ddpm" or k.ddim_sigma != 0.0))

    def _draws(self, idx, img, y, k, need_q=False):
        """z then q — the reference's per-step draw order (p_sample :472/:494, then q_sample :145)."""
        z = self.noise.z(idx, img) if (self.parity_rng or self._needs_z(k)) else None
        q = self.noise.q(idx, y) if (self.parity_rng or need_q) else None
        return z, q

    # -- reference API ----------------------------------------------------------------------------
    def q_sample(self, x_start, t):
        """gaussian_diffusion.py:134-151 (draws randn_like(x_start))."""
        k = self._consts(self._idx(t))
        noise = torch.randn_like(x_start)
        if x_start.is_cuda:
            return kernels.q_sample(x_start, noise, k.sqrt_acp, k.sqrt_1macp)
        return k.sqrt_acp * x_start + k.sqrt_1macp * noise

    def p_mean_variance(self, model, x, t):
        """gaussian_diffusion.py:308-330 — differentiable w.r.t. x through the model like the original."""
        k = self._consts(self._idx(t))
        _, eps, v = self._model_out(model, x, k)
        pre = k.c1 * x - k.c2 * eps
        x0 = self._process_xstart(pre)
        mean = eps if k.mean_mode else k.p1 * x0 + k.p2 * x      # previous_x: the model predicts the mean itself
        if self.var_mode == 0:
            frac = (v + 1.0) / 2.0
            logvar = frac * k.max_log + (1 - frac) * k.min_log
        elif self.var_mode == 1:
            logvar = torch.full_like(x, self._max_log(k))
        else:
            logvar = v
        return {"mean": mean, "variance": torch.exp(logvar), "log_variance": logvar, "pred_xstart": x0}

    def p_sample(self, model, x, t):
        """DDPM.p_sample / DDIM.p_sample (:468-509): returns {'sample', 'pred_xstart'}; pred_xstart stays
        connected to x for autograd-based conditioning, the sample comes from the fused kernel."""
        idx = self._idx(t)
        k = self._consts(idx)
        out, eps, v = self._model_out(model, x, k)
        pre = k.c1 * x - k.c2 * eps
        x0 = self._process_xstart(pre)
        z = self.noise.z(idx, x)
        xd, ed = x.detach(), eps.detach()
        if self.dynamic_threshold:
            return {"sample": self._torch_sample(x, eps, v, x0, z, k), "pred_xstart": x0}
        sample, _, _ = kernels.posterior_update(self.kind, xd, ed, None if v is None else v.detach(), z, k,
                                                clip=self.clip_denoised, var_mode=self.var_mode,
                                                max_log=self._max_log(k))
        return {"sample": sample, "pred_xstart": x0}

    def _torch_sample(self, x, eps, v, x0, z, k):
        """DDPM.p_sample / DDIM.p_sample (:468-509) in torch, for the dynamic-threshold path (x̂₀ comes from
        _process_xstart and cannot be recomputed inside a kernel)."""
        if self.kind == "ddim":
            eps2 = (k.c1 * x - x0) / k.c2
            sample = x0 * k.ddim_sa + k.ddim_sb * eps2
            return sample + k.ddim_sigma * z if (k.noise_on and k.ddim_sigma != 0.0) else sample
        mean = k.p1 * x0 + k.p2 * x
        if self.var_mode == 0:
            frac = (v + 1.0) / 2.0
            logvar = frac * k.max_log + (1 - frac) * k.min_log
        elif self.var_mode == 1:
            logvar = torch.full_like(x, self._max_log(k))
        else:
            logvar = v
        return mean + torch.exp(0.5 * logvar) * z if k.noise_on else mean

    # -- the fused guided step --------------------------------------------------------------------
    def _buffers(self, x):
        key = (tuple(x.shape), x.device)
        if getattr(self, "_buf_key", None) != key:
            n, C, H, W = x.shape
            self._buf_key = key
            self._g6 = torch.zeros((n, 2 * C, H, W), device=x.device, dtype=torch.float32)  # ε | v cotangent, v half stays 0
            self._g3 = torch.zeros((n, C, H, W), device=x.device, dtype=torch.float32)
        return self._g6, self._g3

    @staticmethod
    def _inv_abs_mean(y) -> float:
        """mean(1/|y|) of the measurement (Poisson likelihood, condition_methods.py:52).  y is fixed during a run: the loops
        compute it ONCE per p_sample_loop call (in _prepare) and hand the value down — never cached by address, a sampler
        object outlives its measurements."""
        return float((1.0 / y.abs()).mean())

    def _graphed(self, model, x):
        """The model's forward + input-VJP captured in CUDA graphs for this particle batch (graphed.GraphedEps)."""
        cache = self.__dict__.setdefault("_graphs", {})
        key = (tuple(x.shape), x.device)
        hit = cache.get(key)
        if hit is None or hit.model is not model:       # GraphedEps holds the module: its id cannot be recycled while cached
            if len(cache) >= 4:
                cache.pop(next(iter(cache)))
            hit = cache[key] = GraphedEps(model, tuple(x.shape), x.device)
        return hit

    # -- micro-batching of the ε-model ---------------------------------------------------------------
    # The UNet's saved activations, not the graft, bound the particles per GPU (1.9 GB / particle for the FFHQ model,
    # 5.5 GB for ImageNet-256: SURVEY §7.2).  Every quantity of a guided step is per particle (residual, its norm, the
    # cotangent, the VJP, the update), so a step over n particles can run as ⌈n/chunk⌉ passes of "forward → residual →
    # cotangent → VJP → update" over slices: the graph of a slice is freed before the next one is built, the results are
    # those of the one-pass step bit for bit (kernels are particle-major: no value depends on which particles share a
    # launch) and nothing is recomputed.  `unet_chunk`: None (one pass), an int, or "auto" (sized from free memory).
    unet_chunk = None

    def _chunk_size(self, model, x) -> int:
        n = x.shape[0]
        c = self.unet_chunk
        if c is None:
            return n
        if c == "auto":
            key = (id(model), tuple(x.shape[1:]), x.device)
            if getattr(self, "_auto_chunk_key", None) != key:
                self._auto_chunk_key, self._auto_chunk = key, self._probe_chunk(model, x)
            c = self._auto_chunk
        c = max(1, min(int(c), n))
        while n % c:             # equal slices: one CUDA-graph shape, no ragged tail
            c -= 1
        return c

    @staticmethod
    def _probe_chunk(model, x, reserve=0.15):
        """Particles per pass that fit: peak bytes of a ONE-particle forward + input-VJP, against the free memory."""
        dev = x.device
        torch.cuda.synchronize(dev)
        torch.cuda.reset_peak_memory_stats(dev)
        base = torch.cuda.memory_allocated(dev)
        xi = x[:1].detach().clone().requires_grad_(True)
        with torch.enable_grad():
            out = model(xi, torch.zeros((1,), device=dev, dtype=torch.float32))
            torch.autograd.grad(out, xi, torch.zeros_like(out))
        torch.cuda.synchronize(dev)
        per = max(1, torch.cuda.max_memory_allocated(dev) - base)
        del out, xi
        torch.cuda.empty_cache()
        free, _ = torch.cuda.mem_get_info(dev)
        return max(1, int(free * (1.0 - reserve)) // per)

    def guided_step(self, model, x, idx, measurement, method, spec: GuidanceSpec, cond_kwargs, noisy_measurement=None,
                    z=None, dsg=False, graph_model=False, inv_abs_mean=None, out=None, particle_offset=0):
        """One reverse step with measurement guidance.  Returns (x_next, meas_dist (N,), sem_dist or None).
        `dsg`: DiffStateGrad projection step (gaussian_diffusion.py:240-255) — the gradient is materialised,
        projected onto the sample's leading singular subspaces and applied to every particle.
        `out`: optional destination of x_next (e.g. a symmetric-memory buffer other GPUs read from)."""
        n = x.shape[0]
        poisson = getattr(method.noiser, "__name__", "gaussian") == "poisson" and spec.kind != "ps_semantic"
        chunk = self._chunk_size(model, x)
        if chunk >= n or dsg or poisson:
            # DiffStateGrad applies particle 0's projected gradient to all particles and the Poisson likelihood has ONE norm
            # over all particles: both need the whole batch in one pass
            if chunk < n:
                raise DpsError("unet_chunk: DiffStateGrad steps and the Poisson likelihood couple the particles of a step; "
                               "run them with unet_chunk=None")
            return self._guided_pass(model, x, idx, measurement, method, spec, cond_kwargs, noisy_measurement, z, dsg,
                                     graph_model, inv_abs_mean, out, particle_offset)
        if z is None and self._needs_z(self._consts(idx)):
            z = self.noise.z(idx, x)                         # one draw for the whole batch: the RNG stream does not see the slicing
        x_next = torch.empty_like(x) if out is None else out
        dists, sems = [], []
        for a in range(0, n, chunk):
            sl = slice(a, a + chunk)
            _, d, sd = self._guided_pass(model, x[sl], idx, measurement, method, spec, cond_kwargs,
                                         None if noisy_measurement is None else noisy_measurement, None if z is None else z[sl],
                                         False, graph_model, inv_abs_mean, x_next[sl], particle_offset + a)
            dists.append(d)
            sems.append(sd)
        dist = torch.cat(dists)
        sem = None if sems[0] is None else torch.cat([t.reshape(-1) for t in sems])
        return x_next, dist, sem

    # The per-particle factor of the guidance gradient (−ζ/‖r‖ or −2ζ) commutes with Aᵀ, the clamp mask and the UNet VJP,
    # so by default it is applied by the update kernel (dps_update_ext) to the UNSCALED cotangent: the coefficient launch
    # disappears and operators with a fused residual+cotangent kernel (super-resolution) run ONE launch before the VJP.
    # False restores the round-1 sequence residual → coefficient → cotangent (A/B aid; also taken whenever a step needs the
    # scaled cotangent itself: semantic term, Poisson likelihood, DiffStateGrad).
    deferred_coef = True

    def _philox(self, k, particle_offset, idx):
        if isinstance(self.noise, PhiloxNoise) and self._needs_z(k):
            return (self.noise.seed, idx, particle_offset)
        return None

    def _guided_pass(self, model, x, idx, measurement, method, spec, cond_kwargs, noisy_measurement, z, dsg, graph_model,
                     inv_abs_mean, out, particle_offset=0):
        k = self._consts(idx)
        op = method.operator
        gm = self._graphed(model, x) if graph_model else None
        if gm is not None:
            xd = x.detach()
            mo = gm.forward(xd, k.model_t)               # replayed forward graph; mo is a static buffer
            if mo.shape[1] != 2 * x.shape[1] and self.var_mode in (0, 2):
                raise DpsError(f"model_var_type={self.model_var_type} needs a model with 2·C output channels")
        else:
            x = x.detach().requires_grad_(True)
            with torch.enable_grad():
                mo, eps, v = self._model_out(model, x, k)
            xd = x.detach()
        out_d = mo.detach()
        C = x.shape[1]
        eps_d = out_d[:, :C] if out_d.shape[1] == 2 * C else out_d
        v_d = out_d[:, C:] if out_d.shape[1] == 2 * C else None
        clip = self.clip_denoised
        two_c = mo.shape[1] == 2 * C
        if gm is not None:
            g = gm.cotangent[:, :C] if two_c else gm.cotangent
        else:
            g6, g3 = self._buffers(xd)
            g = g6[:, :C] if two_c else g3
        poisson = getattr(method.noiser, "__name__", "gaussian") == "poisson" and spec.kind != "ps_semantic"
        deferred = None
        extra, sem_dist = None, None
        if self.deferred_coef and not dsg and not poisson and spec.semantic is None \
                and spec.coef_mode in (DPS_COEF_NORM, DPS_COEF_NORM_SQ):
            # kernels 1-3 as ONE call: residual, partial sums and the unscaled masked cotangent (fused where the operator
            # has such a kernel); the coefficient is applied by the update kernel below
            partials, _, aux = op.guidance(xd, eps_d, k, clip, measurement, out=g, **cond_kwargs)
            dist = torch.empty((x.shape[0],), device=x.device, dtype=torch.float32)
            deferred = (partials, spec.coef_mode, spec.scale, dist)
        else:
            # kernel 1: residual + partial sums, x̂₀ formed on the fly
            r, partials, aux = op.residual(xd, eps_d, k, clip, measurement, **cond_kwargs)
            # kernel 2: ‖r‖ and the per-particle coefficient
            if poisson:
                # Poisson branch of grad_and_value (condition_methods.py:50-55): loss = ‖r‖_F(all particles)·mean(1/|y|);
                # norm_exp is ignored there, so ps_anneal's ζ_t multiplies the same gradient
                inv = self._inv_abs_mean(measurement) if inv_abs_mean is None else inv_abs_mean
                l2, coef = kernels.guidance_coef(partials, DPS_COEF_GLOBAL_NORM, spec.scale * inv)
                dist = torch.linalg.norm(l2) * inv            # the scalar the reference returns as `norm`
            else:
                dist, coef = kernels.guidance_coef(partials, spec.coef_mode, spec.scale)
            # optional semantic term: gradient w.r.t. x̂₀ of s_t·ℓ_sem (external embedder stays PyTorch)
            if spec.semantic is not None:
                x0 = kernels.x0_from_eps(xd, eps_d, k, clip).requires_grad_(True)
                with torch.enable_grad():
                    sem_loss, sem_dist = spec.semantic(x0)
                    extra = torch.autograd.grad(sem_loss.sum(), x0)[0].contiguous()
                sem_dist = sem_dist.detach()
            # kernel 3: cotangent w.r.t. the pre-clamp x̂₀ written into the ε-channels of the cotangent buffer
            op.cotangent(r, coef, xd, eps_d, k, clip, extra, out=g, aux=aux, **cond_kwargs)
        # UNet VJP
        vjp = None
        if gm is not None:
            vjp = gm.vjp()                               # replayed backward graph (input gradient only)
        elif mo.requires_grad:
            vjp = torch.autograd.grad(mo, x, grad_outputs=g6 if two_c else g3)[0]
        # kernel 4: fused posterior update
        if z is None and self._needs_z(k):
            z = self.noise.z(idx, xd)
        if dsg:
            sample, _, _ = kernels.posterior_update(self.kind, xd, eps_d, v_d, z, k, clip=clip,
                                                    var_mode=self.var_mode, max_log=self._max_log(k),
                                                    philox=self._philox(k, particle_offset, idx))
            x_next = diffstategrad.projected_update(sample, kernels.guidance_grad(g, vjp, k))
            if out is not None:
                x_next = out.copy_(x_next)
        else:
            x_next, _, _ = kernels.posterior_update(self.kind, xd, eps_d, v_d, z, k, clip=clip, g=g,
                                                    vjp=vjp, var_mode=self.var_mode, max_log=self._max_log(k), out=out,
                                                    deferred=deferred, philox=self._philox(k, particle_offset, idx))
        if spec.project:  # mcg: x_t = operator.project(x_t, noisy_measurement)
            x_next = method.project(data=x_next, noisy_measurement=noisy_measurement, **cond_kwargs)
            if out is not None and x_next.data_ptr() != out.data_ptr():
                x_next = out.copy_(x_next)
        return x_next, dist, sem_dist

    def _generic_step(self, model, img, idx, measurement, cond_fn, extra_kw, dsg=False):
        """Foreign conditioning function (e.g. the reference's own class): autograd end to end, the
        x̂₀/sample arithmetic still fused.  Handles every return arity of HEAD."""
        img = img.detach().requires_grad_(True)
        with torch.enable_grad():
            out = self.p_sample(model=model, x=img, t=idx)   # draws z
            k = self._consts(idx)
            q_noise = self.noise.q(idx, measurement)            # then the q_sample draw
            noisy_measurement = None if q_noise is None else kernels.q_sample(measurement, q_noise, k.sqrt_acp, k.sqrt_1macp)
            sample = out["sample"].clone()
            res = cond_fn(x_t=sample, measurement=measurement, noisy_measurement=noisy_measurement, x_prev=img,
                          x_0_hat=out["pred_xstart"], **extra_kw)
        if not isinstance(res, tuple):
            return res.detach(), None, None
        first, dist = res[0], (res[1] if len(res) > 1 else None)
        third = res[2] if len(res) > 2 else None
        returns_grad = _returns_gradient(_resolve_cond_fn(cond_fn)[0])
        if returns_grad and dsg:
            x_next = diffstategrad.projected_update(out["sample"].detach(), first.detach().contiguous())
        else:
            x_next = out["sample"] - first if returns_grad else first
        third = third.detach() if (returns_grad and torch.is_tensor(third)) else (third if returns_grad else None)
        return x_next.detach(), (dist.detach() if torch.is_tensor(dist) else dist), third

    def _step_indices(self, kwargs):
        """All steps T−1 … 0 like the reference, or a window of the chain: `start_idx` (default T−1) and
        `num_steps` — extra kwargs the reference's **kwargs signature tolerates; bench.py times K steps."""
        hi = int(kwargs.get("start_idx", self.num_timesteps - 1))
        n = kwargs.get("num_steps")
        lo = 0 if n is None else max(0, hi - int(n) + 1)
        return range(hi, lo - 1, -1)

    def _prepare(self, x_start, measurement, measurement_cond_fn):
        if not x_start.is_cuda:
            raise DpsError("x_start must be a CUDA tensor: the B200 samplers have no CPU path")
        img = x_start.detach().to(torch.float32).contiguous()
        y = measurement.detach().to(img.device, torch.float32).contiguous()
        method, fn, bound = _resolve_cond_fn(measurement_cond_fn)
        fused = isinstance(method, ConditioningMethod) and isinstance(method.operator, B200Operator) \
            and getattr(method.noiser, "__name__", "gaussian") in ("gaussian", "poisson") \
            and method.guidance().kind != "none" and not self.dynamic_threshold
        # Poisson likelihood: mean(1/|y|) once per loop call (one host sync here instead of one per step)
        self._loop_inv = self._inv_abs_mean(y) if (fused and getattr(method.noiser, "__name__", "") == "poisson") else None
        return img, y, method, bound, fused

    # -- base loop (GaussianDiffusion.p_sample_loop, :175-303) --------------------------------------
    def p_sample_loop(self, model, x_start, measurement, measurement_cond_fn, record=False, save_root=None, **kwargs):
        img, y, method, bound, fused = self._prepare(x_start, measurement, measurement_cond_fn)
        fused = fused and kwargs.get("fused", True)   # fused=False forces the generic autograd path (debugging / tests)
        anneal_kw = {k_: kwargs[k_] for k_ in ("anneal_amp", "anneal_scale", "anneal_loc") if k_ in kwargs}
        callback = kwargs.get("callback")
        # DiffStateGrad (gaussian_diffusion.py:203-204, :240-253): `project`, `period` as upstream
        dsg_on, period = bool(kwargs.get("project", False)), int(kwargs.get("period", 20))
        meas_d = sem_d = None
        for idx in self._step_indices(kwargs):
            k = self._consts(idx)
            t = idx / self.num_timesteps
            dsg = dsg_on and period != 0 and idx % period == 0
            if fused:
                anneal = anneal_factor(t, kwargs.get("anneal_amp", 1.0), kwargs.get("anneal_scale", 10.0),
                                       kwargs.get("anneal_loc", 0.5)) if anneal_kw else 1.0
                spec = method.guidance(beta_scale=k.beta, t=t, anneal=anneal)
                z, q_noise = self._draws(idx, img, y, k, need_q=spec.project)
                noisy = kernels.q_sample(y, q_noise, k.sqrt_acp, k.sqrt_1macp) if spec.project else None
                img, meas_d, sem_d = self.guided_step(model, img, idx, y, method, spec, bound, noisy, z, dsg=dsg,
                                                      graph_model=bool(kwargs.get("graph_model", False)),
                                                      inv_abs_mean=self._loop_inv)
            else:
                img, meas_d, sem_d = self._generic_step(model, img, idx, y, measurement_cond_fn,
                                                        {"beta_scale": k.beta, "t": t}, dsg=dsg)
            if callback is not None:
                callback(idx, img, meas_d, sem_d)
        if sem_d is None:
            sem_d = torch.zeros((), device=img.device)
        return img, meas_d, sem_d


@register_sampler(name="ddpm")
class DDPM(SpacedSampler):
    kind = "ddpm"


@register_sampler(name="ddim")
class DDIM(SpacedSampler):
    kind = "ddim"

    def predict_eps_from_x_start(self, x_t, t, pred_xstart):
        k = self._consts(self._idx(t))
        return (k.c1 * x_t - pred_xstart) / k.c2


# ------------------------------------------------------------------------------------------------
# particle search
# ------------------------------------------------------------------------------------------------
@register_sampler(name="search_ddpm")
class SearchDDPM(DDPM):
    """Greedy best-of-N inside the loop (gaussian_diffusion.py:592-641) and the resample_update
    potentials (:516-587)."""

    sync_free = False   # False: host check of max(w) != min(w) before drawing, like the reference (:545) — keeps its RNG stream.
                        # True: no host sync; the ancestors kernel returns the identity when the weights are degenerate
                        # (the uniforms are drawn either way, so the RNG stream differs from the reference's in that case)

    @torch.no_grad()
    def resample_update(self, candidates, denoised_candidates, operator, measurement, resample=True, rs_temp=0.01,
                        prev_costs=None, potential_type="min", steps_done=1, uniforms=None, **op_kwargs):
        n = denoised_candidates.shape[0]
        if resample and prev_costs is not None:
            tau = rs_temp / steps_done if potential_type == "mean" else rs_temp
            logw = kernels.particle_logweights(prev_costs.float().contiguous(), tau=tau)
            w, cdf, _, degenerate = kernels.weights_cdf(logw, linear_mode=True)
            if self.sync_free or not bool(degenerate.item()):
                u = self.noise.uniforms(-1, n, candidates.device) if uniforms is None else uniforms
                ids = kernels.ancestors(cdf, u, n, degenerate=degenerate if self.sync_free else None)
                candidates = kernels.gather_particles(candidates, ids)
                denoised_candidates = kernels.gather_particles(denoised_candidates, ids)
                prev_costs = prev_costs[ids]
        _, partials, _ = operator.residual(denoised_candidates.contiguous(), y=measurement, **op_kwargs)
        _, l1 = kernels.particle_norms(partials, want_l1=True)
        _, C, H, W = denoised_candidates.shape
        curr = l1 ** 2 / (C * H * W)
        if prev_costs is None or potential_type == "curr":
            net = curr
        elif potential_type == "mean":
            net = curr + prev_costs
        elif potential_type == "min":
            net = torch.minimum(curr, prev_costs)
        elif potential_type == "diff":
            net = curr - prev_costs
        else:
            raise NotImplementedError
        return candidates, net

    def p_sample_loop(self, model, x_start, measurement, measurement_cond_fn, record=False, save_root=None,
                      operator=None, potential_type="min", resample_every_steps=10, rs_temp=0.1, **kwargs):
        if operator is None:
            raise TypeError("search_ddpm.p_sample_loop needs operator=")
        img, y, _, bound, _ = self._prepare(x_start, measurement, measurement_cond_fn)
        shards = kwargs.get("shards")
        n = img.shape[0]
        self.last_stats = {"best": []}
        for idx in self._step_indices(kwargs):
            k = self._consts(idx)
            with torch.no_grad():
                _, eps, v = self._model_out(model, img, k)
                z = self.noise.z(idx, img)
                if self.dynamic_threshold:
                    img = self._torch_sample(img, eps, v, self._process_xstart(k.c1 * img - k.c2 * eps), z, k)
                else:
                    img, _, _ = kernels.posterior_update("ddpm", img, eps, v, z, k, clip=self.clip_denoised,
                                                         var_mode=self.var_mode, max_log=self._max_log(k),
                                                         philox=self._philox(k, 0 if shards is None else shards.offset, idx))
                _, partials, _ = operator.residual(img, y=y, **bound)      # ‖y − A(x_{t−1})‖₂, :626-630
                costs = kernels.particle_norms(partials)
                if shards is None:
                    best, _ = kernels.argmin(costs)                            # first minimum, :631
                    img = kernels.broadcast_particle(img, best, n)             # img[best.repeat(n)], :633
                else:
                    img = shards.greedy_broadcast(img, costs)
        return img


@register_sampler(name="ttc_ddim")
class TTC_DDIM(DDIM):
    """DDIM + guidance + multinomial particle resampling every 10th index with w = exp(−d/100)
    (gaussian_diffusion.py:644-707).  `scheme`, `resample_every_steps`, `resample_scale` generalise the
    constants the reference hard-codes (:661, :689); `shards` spreads the particles over ranks."""
    resample_every_steps = 10
    resample_scale = 100.0
    scheme = "multinomial"
    lse_weights = False      # True: w = exp(logw − max) instead of the reference's exp(logw)
    sync_free = True         # False: host check of w.max() != w.min() like the reference (:693)
    resample_seed = 0        # sharded runs: seed of the uniforms all ranks draw identically

    def p_sample_loop(self, model, x_start, measurement, measurement_cond_fn, record=False, save_root=None, **kwargs):
        img, y, method, bound, fused = self._prepare(x_start, measurement, measurement_cond_fn)
        fused = fused and kwargs.get("fused", True)
        shards = kwargs.get("shards")
        sem_weight = kwargs.get("semantic_weight", 0.0)  # config 5: semantic term in the reweighting
        callback = kwargs.get("callback")
        graph_model = bool(kwargs.get("graph_model", False))
        anneal_kw = any(k_ in kwargs for k_ in ("anneal_amp", "anneal_scale", "anneal_loc"))
        distance = None
        self.last_stats = {"ancestors": {}}
        for idx in self._step_indices(kwargs):
            k = self._consts(idx)
            sem_d = None
            n_total = img.shape[0] if shards is None else shards.total
            resampling = n_total > 1 and idx % self.resample_every_steps == 0
            if fused:
                # the reference's TTC loop hands the conditioning function neither beta_scale nor t (:672-676), so
                # ps_anneal runs with its constructor scale and ps_semantic with t = 1 — as the generic path below does
                # (extension: the annealing-schedule kwargs of the base loop, when given, scale the guidance here too)
                spec = method.guidance(anneal=anneal_factor(idx / self.num_timesteps, kwargs.get("anneal_amp", 1.0),
                                                            kwargs.get("anneal_scale", 10.0), kwargs.get("anneal_loc", 0.5))) \
                    if anneal_kw else method.guidance()
                z, q_noise = self._draws(idx, img, y, k, need_q=spec.project)
                noisy = kernels.q_sample(y, q_noise, k.sqrt_acp, k.sqrt_1macp) if spec.project else None
                # sharded resampling step: the update kernel writes x_{t-1} straight into the symmetric buffer the
                # other GPUs read from (no staging copy)
                tgt = shards.publish_target(img) if (resampling and shards is not None and img.is_cuda) else None
                if tgt is not None and tgt.data_ptr() == img.data_ptr():
                    tgt = None
                img, distance, sem_d = self.guided_step(model, img, idx, y, method, spec, bound, noisy, z,
                                                        graph_model=graph_model, inv_abs_mean=self._loop_inv, out=tgt,
                                                        particle_offset=0 if shards is None else shards.offset)
            else:
                img, distance, _ = self._generic_step(model, img, idx, y, measurement_cond_fn, {})
            if resampling:
                logw = kernels.particle_logweights(distance.contiguous(), sem_d if sem_weight else None,
                                                   tau=1.0 / self.resample_scale, sem_scale=sem_weight)
                if shards is None:
                    logw_all, dist_all = logw, None
                else:   # ONE all-gather carries the log-weights and the distances that travel with the particles
                    both = shards.all_gather_scalars(torch.stack((logw, distance.reshape(-1)), dim=1))
                    logw_all, dist_all = both[:, 0].contiguous(), both[:, 1]
                w, cdf, _, degenerate = kernels.weights_cdf(logw_all, linear_mode=not self.lse_weights)
                if not self.sync_free and bool(degenerate.item()):
                    if callback is not None:
                        callback(idx, img, distance, sem_d)
                    continue            # the reference draws nothing when max(w) == min(w) (:693)
                n_u = 1 if self.scheme == "systematic" else n_total
                if shards is None:
                    u = self.noise.uniforms(idx, n_u, img.device)
                else:  # every rank must see the same uniforms: a CPU generator keyed by (seed, step)
                    from .dist import shared_uniforms
                    u = shared_uniforms(self.resample_seed, idx, n_u, img.device)
                ids = kernels.ancestors(cdf, u, n_total, systematic=(self.scheme == "systematic"), degenerate=degenerate)
                self.last_stats["ancestors"][idx] = ids
                if shards is None:
                    img = kernels.gather_particles(img, ids)
                    distance = distance[ids]
                else:
                    img, distance = shards.exchange(img, distance, ids, dist_all=dist_all)
            if callback is not None:
                callback(idx, img, distance, sem_d)
        return img, distance
