"""Generate tests/golden/*.npz by RUNNING THE REFERENCE ITSELF (CPU, this container only).

    python oracle/make_golden.py            # needs /root/reference (read-only) — never runs on the GPU box

The reference has no tests or golden vectors; these fixtures are what pins the oracle
(oracle/dps_oracle.py) and, through it, the CUDA kernels.  Everything is seeded; draws made inside the
reference's loops (torch.randn_like, torch.multinomial) are recorded by wrapping the torch functions, so
the fixtures do not depend on torch's RNG staying bit-stable across versions.
"""
from __future__ import annotations

import os
import sys
from functools import partial

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, "tests"))

from dps_ttc_b200 import _ref  # noqa: E402  (only for locating the reference + import stubs)
from helpers import GOLDEN, SemEmbedder, TinyEps, seeded_randn, seeded_unet, tensor_checksum  # noqa: E402

_ref.ensure_reference()
with _ref.quiet():
    from guided_diffusion.condition_methods import get_conditioning_method
    from guided_diffusion.gaussian_diffusion import create_sampler
    from guided_diffusion.measurements import get_noise, get_operator
    from util.img_utils import mask_generator
    from util.resizer import Resizer

torch.set_num_threads(4)
DIFF = dict(steps=1000, noise_schedule="linear", model_mean_type="epsilon", model_var_type="learned_range",
            dynamic_threshold=False, clip_denoised=True, rescale_timesteps=True)


def save(name, **arrays):
    os.makedirs(GOLDEN, exist_ok=True)
    out = {}
    for k, v in arrays.items():
        if torch.is_tensor(v):
            v = v.detach().cpu().numpy()
        out[k] = np.asarray(v)
    path = os.path.join(GOLDEN, name)
    np.savez_compressed(path, **out)
    print(f"{name}: {os.path.getsize(path) / 1024:.0f} KiB, {len(out)} arrays")


class Recorder:
    """Records every torch.randn_like / torch.multinomial / torch.rand call made while active."""

    def __enter__(self):
        self.randn, self.multinomial = [], []
        self._randn_like, self._multinomial = torch.randn_like, torch.multinomial

        def randn_like(t, *a, **k):
            out = self._randn_like(t, *a, **k)
            self.randn.append(out.detach().clone())
            return out

        def multinomial(w, n, replacement=False, **k):
            state = torch.get_rng_state()
            out = self._multinomial(w, n, replacement, **k)
            after = torch.get_rng_state()
            torch.set_rng_state(state)
            u = torch.rand(n, dtype=torch.float64)          # the uniforms the CPU kernel consumed
            assert torch.equal(torch.get_rng_state(), after), "torch.multinomial RNG consumption changed"
            self.multinomial.append((w.detach().clone(), out.detach().clone(), u))
            return out

        torch.randn_like, torch.multinomial = randn_like, multinomial
        return self

    def __exit__(self, *exc):
        torch.randn_like, torch.multinomial = self._randn_like, self._multinomial


def sampler(name, respacing, **overrides):
    with _ref.quiet():
        return create_sampler(sampler=name, timestep_respacing=respacing, **{**DIFF, **overrides})


# ---------------------------------------------------------------------------------------------
def gen_schedule():
    out = {}
    for tag, resp in (("full", ""), ("r50", "50"), ("r12", "12")):
        s = sampler("ddpm", resp)
        out[f"{tag}_timestep_map"] = np.array(s.timestep_map)
        for attr in ("betas", "alphas_cumprod", "alphas_cumprod_prev", "sqrt_recip_alphas_cumprod",
                     "sqrt_recipm1_alphas_cumprod", "posterior_mean_coef1", "posterior_mean_coef2",
                     "posterior_log_variance_clipped", "sqrt_alphas_cumprod", "sqrt_one_minus_alphas_cumprod"):
            out[f"{tag}_{attr}"] = getattr(s, attr)
    save("schedule.npz", **out)


def gen_resizer():
    out = {}
    for n, s in ((256, 4), (256, 8), (64, 4), (32, 4)):
        r = Resizer((1, 3, n, n), 1 / s)
        out[f"fov_{n}_{s}"] = r.field_of_view[0].numpy()
        out[f"w_{n}_{s}"] = r.weights[0].numpy().reshape(r.weights[0].shape[0], -1)
        out[f"sorted_dims_{n}_{s}"] = np.array(r.sorted_dims)
    save("resizer.npz", **out)


def norm_grad(op, x, y, **kw):
    x = x.clone().requires_grad_(True)
    diff = y - op.forward(x, **kw)
    norm = torch.linalg.norm(diff.reshape(diff.shape[0], -1), dim=-1)
    (g,) = torch.autograd.grad(norm.sum(), x)
    return norm.detach(), g


def gen_operators():
    g = torch.Generator().manual_seed(1234)
    out = {}
    x = torch.rand(2, 3, 64, 64, generator=g) * 2 - 1
    out["x"] = x
    with _ref.quiet():
        ops = {"gaussian_blur": (get_operator("gaussian_blur", kernel_size=61, intensity=3.0, device="cpu"), {}),
               "super_resolution": (get_operator("super_resolution", in_shape=(1, 3, 64, 64), scale_factor=4, device="cpu"), {}),
               "phase_retrieval": (get_operator("phase_retrieval", oversample=2.0, device="cpu"), {})}
        np.random.seed(8)
        ops["motion_blur"] = (get_operator("motion_blur", kernel_size=61, intensity=0.5, device="cpu"), {})
        out["motion_kernel"] = ops["motion_blur"][0].kernel.kernelMatrix
        np.random.seed(8)
        mask = mask_generator("random", mask_prob_range=(0.3, 0.7), image_size=64)(x[:1])[:, 0].unsqueeze(0)
        out["mask"] = mask
        ops["inpainting"] = (get_operator("inpainting", device="cpu"), {"mask": mask})
    out["gaussian_kernel"] = ops["gaussian_blur"][0].kernel.numpy()
    for name, (op, kw) in ops.items():
        with torch.no_grad():
            ax = op.forward(x, **kw)
        y = op.forward(x[:1], **kw).detach() + 0.05 * torch.randn(ax[:1].shape, generator=g)
        norm, grad = norm_grad(op, x, y, **kw)
        out[f"{name}_Ax"], out[f"{name}_y"], out[f"{name}_norm"], out[f"{name}_grad"] = ax, y, norm, grad
    save("operators.npz", **out)


def run_trace(tag, sampler_name, respacing, method, params, op_name, op_cfg, n, size, seed, use_loop=True, mask=None,
              noise_sigma=0.05, anneal=False, loop_kwargs=None, noise=None, diffusion=None):
    """Run the reference's own loop (or an upstream-arity loop assembled from its classes) and record."""
    torch.manual_seed(seed)
    np.random.seed(seed)
    model = TinyEps(seed=seed)
    with _ref.quiet():
        op = get_operator(op_name, device="cpu", **op_cfg)
        noiser = get_noise(**noise) if noise else get_noise("gaussian", sigma=noise_sigma)
        cond = get_conditioning_method(method, op, noiser, **params)
    s = sampler(sampler_name, respacing, **(diffusion or {}))
    kw = {"mask": mask} if mask is not None else {}
    x_true = torch.rand(1, 3, size, size) * 2 - 1
    y = noiser(op.forward(x_true, **kw)).detach()
    x_start = torch.randn(n, 3, size, size)
    steps = []

    def cond_fn(**k):
        res = cond.conditioning(**kw, **k)
        rec = {"x_prev": k["x_prev"].detach().clone(), "x0": k["x_0_hat"].detach().clone()}
        if method == "ps_semantic":
            rec["sample"] = k["x_t"].detach().clone()
            rec["grad"], rec["dist"] = res[0].detach().clone(), res[1].detach().clone()
        else:
            rec["x_t_out"], rec["dist"] = res[0].detach().clone(), res[1].detach().clone()
        steps.append(rec)
        return res

    with Recorder() as rec, _ref.quiet():
        if use_loop:
            extra = {"operator": op} if sampler_name == "search_ddpm" else {}
            extra.update(loop_kwargs or {})
            result = s.p_sample_loop(model=model, x_start=x_start.clone(), measurement=y, measurement_cond_fn=cond_fn,
                                     record=False, save_root=None, **extra)
        else:
            img = x_start.clone()
            for idx in reversed(range(s.num_timesteps)):
                t = torch.tensor([idx])
                img = img.requires_grad_()
                out = s.p_sample(x=img, t=t, model=model)
                extra = {"beta_scale": s.betas[idx], "anneal": 1.0} if anneal else {}
                res = cond_fn(x_t=out["sample"], measurement=y, x_prev=img, x_0_hat=out["pred_xstart"], **extra)
                img = res[0].detach()
            result = (img, res[1].detach())
    final = result if torch.is_tensor(result) else result[0]
    out = {"x_start": x_start, "y": y, "final": final.detach(), "n_steps": np.array(s.num_timesteps)}
    if not torch.is_tensor(result):
        out["final_dist"] = result[1].detach()
    if mask is not None:
        out["mask"] = mask
    if op_name == "motion_blur":
        out["kernel"] = op.kernel.kernelMatrix
    for i, r in enumerate(rec.randn):
        out[f"randn_{i}"] = r
    for i, (w, ids, u) in enumerate(rec.multinomial):
        out[f"mn_w_{i}"], out[f"mn_ids_{i}"], out[f"mn_u_{i}"] = w, ids, u
    for i, st in enumerate(steps):
        for k2, v in st.items():
            out[f"step{i}_{k2}"] = v
    save(f"trace_{tag}.npz", **out)


def gen_multinomial():
    out = {}
    for i, n in enumerate((4, 8, 64, 256)):
        g = torch.Generator().manual_seed(100 + i)
        d = torch.rand(n, generator=g) * 40 + 60
        w = torch.exp(-d / 100)
        torch.manual_seed(200 + i)
        ids = torch.multinomial(w, n, replacement=True)
        torch.manual_seed(200 + i)
        u = torch.rand(n, dtype=torch.float64)
        out[f"d_{i}"], out[f"w_{i}"], out[f"ids_{i}"], out[f"u_{i}"] = d, w, ids, u
    save("multinomial.npz", **out)


def gen_masks():
    out = {}
    img = torch.zeros(1, 3, 256, 256)
    for seed in (0, 8):
        np.random.seed(seed)
        out[f"random_{seed}"] = mask_generator("random", mask_prob_range=(0.3, 0.7), image_size=256)(img)[0, 0].numpy().astype(np.uint8)
        np.random.seed(seed)
        out[f"box_{seed}"] = mask_generator("box", mask_len_range=(128, 129), image_size=256)(img)[0, 0].numpy().astype(np.uint8)
    save("masks.npz", **out)


def gen_var_types():
    """DDPM.p_sample with the other variance processors (posterior_mean_variance.py:159-228)."""
    out = {}
    g = torch.Generator().manual_seed(31)
    x = torch.randn(2, 3, 16, 16, generator=g)
    out["x"] = x
    model = TinyEps(seed=31)
    for var_type in ("fixed_small", "fixed_large", "learned", "learned_range"):
        with _ref.quiet():
            s = create_sampler(sampler="ddpm", timestep_respacing="", **{**DIFF, "model_var_type": var_type})
        for idx in (999, 500, 1, 0):
            with Recorder() as rec, _ref.quiet(), torch.no_grad():
                o = s.p_sample(model=model, x=x, t=torch.tensor([idx]))
            out[f"{var_type}_{idx}_sample"] = o["sample"]
            out[f"{var_type}_{idx}_x0"] = o["pred_xstart"]
            out[f"{var_type}_{idx}_z"] = rec.randn[0]
    save("var_types.npz", **out)


def gen_mean_types():
    """DDPM.p_sample + the `ps` conditioning step with every mean processor (posterior_mean_variance.py:45-129):
    sample, pred_xstart, the guided x_{t-1} and the distance, through the reference's own autograd chain."""
    out = {}
    g = torch.Generator().manual_seed(53)
    x = torch.randn(2, 3, 32, 32, generator=g)
    with _ref.quiet():
        op = get_operator("gaussian_blur", kernel_size=61, intensity=3.0, device="cpu")
        noiser = get_noise("gaussian", sigma=0.05)
        cond = get_conditioning_method("ps", op, noiser, scale=0.3)
    y = op.forward(torch.rand(1, 3, 32, 32, generator=g) * 2 - 1).detach()
    out.update(x=x, y=y)
    model = TinyEps(seed=53)
    for mean_type in ("epsilon", "start_x", "previous_x"):
        with _ref.quiet():
            s = create_sampler(sampler="ddpm", timestep_respacing="", **{**DIFF, "model_mean_type": mean_type})
        for idx in (999, 500, 1, 0):
            tag = f"{mean_type}_{idx}"
            if mean_type == "previous_x":
                # the reference cannot differentiate this processor (p_sample's in-place `sample +=` hits a view of
                # torch.split, gaussian_diffusion.py:474): only the unguided step exists upstream
                with Recorder() as rec, _ref.quiet(), torch.no_grad():
                    o = s.p_sample(model=model, x=x.clone(), t=torch.tensor([idx]))
                out[f"{tag}_sample"], out[f"{tag}_x0"], out[f"{tag}_z"] = o["sample"], o["pred_xstart"], rec.randn[0]
                continue
            xi = x.clone().requires_grad_(True)
            with Recorder() as rec, _ref.quiet():
                o = s.p_sample(model=model, x=xi, t=torch.tensor([idx]))
                sample = o["sample"].detach().clone()        # `ps` updates x_t in place (condition_methods.py:103)
                res = cond.conditioning(x_t=o["sample"], measurement=y, noisy_measurement=None, x_prev=xi,
                                        x_0_hat=o["pred_xstart"])
            out[f"{tag}_sample"], out[f"{tag}_x0"], out[f"{tag}_z"] = sample, o["pred_xstart"], rec.randn[0]
            out[f"{tag}_next"], out[f"{tag}_dist"] = res[0], res[1]
    save("mean_types.npz", **out)


def gen_resample_update():
    """SearchDDPM.resample_update (gaussian_diffusion.py:516-587) for every potential type."""
    out = {}
    torch.manual_seed(41)
    with _ref.quiet():
        s = create_sampler(sampler="search_ddpm", timestep_respacing="", **DIFF)
        op = get_operator("gaussian_blur", kernel_size=61, intensity=3.0, device="cpu")
    cand = torch.randn(6, 3, 64, 64)
    den = torch.rand(6, 3, 64, 64) * 2 - 1
    y = op.forward(torch.rand(1, 3, 64, 64) * 2 - 1).detach()
    prev = torch.rand(6) * 5 + 20
    out.update(cand=cand, den=den, y=y, prev=prev)
    for pt in ("min", "mean", "diff", "curr"):
        with Recorder() as rec, _ref.quiet():
            c2, net = s.resample_update(cand.clone(), den.clone(), op, y, resample=True, rs_temp=0.05,
                                        prev_costs=prev.clone(), potential_type=pt, steps_done=3)
        out[f"{pt}_cand"], out[f"{pt}_net"] = c2, net
        if rec.multinomial:
            out[f"{pt}_ids"], out[f"{pt}_u"] = rec.multinomial[0][1], rec.multinomial[0][2]
    with _ref.quiet():
        c2, net = s.resample_update(cand.clone(), den.clone(), op, y, prev_costs=None, potential_type="min")
    out["first_net"] = net
    save("resample_update.npz", **out)


def gen_psnr():
    """compute_psnr_manual of the reference (compute_metrics.py:93-98).  The module imports torchmetrics / lpips
    (absent here) at the top, so only that function's source is compiled."""
    import ast
    src = open("/root/reference/compute_metrics.py").read()
    fn = [n for n in ast.parse(src).body if isinstance(n, ast.FunctionDef) and n.name == "compute_psnr_manual"][0]
    ns = {"torch": torch}
    exec(compile(ast.Module([fn], []), "compute_metrics.py", "exec"), ns)
    g = torch.Generator().manual_seed(3)
    real = torch.rand(1, 3, 32, 32, generator=g) * 2 - 1
    fake = real + 0.1 * torch.randn(5, 3, 32, 32, generator=g)
    vals = np.array([float(ns["compute_psnr_manual"](real, fake[i:i + 1])) for i in range(5)], dtype=np.float32)
    save("psnr_manual.npz", real=real.numpy(), fake=fake.numpy(), psnr=vals)


class SeededDraws:
    """Replaces torch.randn_like while active: the i-th call returns seeded_randn(base, i, shape) — a tape both the
    reference run here and the GPU test can regenerate, so multi-MB noise tensors need not be committed."""

    def __init__(self, base):
        self.base, self.calls = base, 0

    def __enter__(self):
        self._orig = torch.randn_like

        def randn_like(t, *a, **k):
            out = seeded_randn(self.base, self.calls, t.shape).to(t.device, t.dtype)
            self.calls += 1
            return out
        torch.randn_like = randn_like
        return self

    def __exit__(self, *exc):
        torch.randn_like = self._orig


def gen_phase256():
    """Phase retrieval at the size the CUDA kernels are built for (256² → 384²): operator output, ‖r‖, its gradient and a
    2-step ps_anneal trace from the REFERENCE; inputs are regenerated from seeds (checksums stored), outputs stored."""
    g = torch.Generator().manual_seed(2560)
    x = torch.rand(1, 3, 256, 256, generator=g) * 2 - 1
    y = torch.rand(1, 3, 384, 384, generator=g) * 1.5           # any non-negative array is a valid |FFT| measurement here
    with _ref.quiet():
        op = get_operator("phase_retrieval", oversample=2.0, device="cpu")
        noiser = get_noise("gaussian", sigma=0.05)
        cond = get_conditioning_method("ps_anneal", op, noiser, scale=1.0)
    with torch.no_grad():
        ax = op.forward(x)
    norm, grad = norm_grad(op, x, y)
    out = {"x_sum": tensor_checksum(x), "y_sum": tensor_checksum(y), "Ax_sub": ax[..., ::3, ::3], "norm": norm, "grad": grad}
    # trace: DDPM.p_sample + PosterorSamplingAnnealing.conditioning, respacing "2", one particle, upstream-arity loop
    s = sampler("ddpm", "2")
    model = TinyEps(seed=25)
    x_start = seeded_randn(2600, 0, (1, 3, 256, 256))
    img = x_start.clone()
    with SeededDraws(2700) as draws, _ref.quiet():
        for i, idx in enumerate(reversed(range(s.num_timesteps))):
            img = img.requires_grad_()
            o = s.p_sample(x=img, t=torch.tensor([idx]), model=model)
            res = cond.conditioning(x_t=o["sample"], measurement=y, x_prev=img, x_0_hat=o["pred_xstart"],
                                    beta_scale=s.betas[idx], anneal=1.0)
            img = res[0].detach()
            out[f"step{i}_dist"] = res[1].detach()
            if i == 0:
                out["step0_x_t_out"] = img.clone()
    out["final"], out["n_draws"] = img, np.array(draws.calls)
    save("phase256.npz", **out)


def gen_project():
    """§8f row 2: LinearOperator.ortho_project / project (measurements.py:48-54), the SR and inpainting overrides (:90-91,
    :167-168) and the `projection` conditioning method (condition_methods.py:72-82) from the reference's classes.
    (NonLinearOperator.project adds data and measurement, which have different shapes for phase retrieval: unusable at HEAD.)"""
    g = torch.Generator().manual_seed(4321)
    x = torch.rand(2, 3, 64, 64, generator=g) * 2 - 1
    out = {"x": x}
    with _ref.quiet():
        ops = {"gaussian_blur": (get_operator("gaussian_blur", kernel_size=61, intensity=3.0, device="cpu"), {}),
               "super_resolution": (get_operator("super_resolution", in_shape=(1, 3, 64, 64), scale_factor=4, device="cpu"), {})}
        np.random.seed(8)
        ops["motion_blur"] = (get_operator("motion_blur", kernel_size=61, intensity=0.5, device="cpu"), {})
        out["motion_kernel"] = ops["motion_blur"][0].kernel.kernelMatrix
        np.random.seed(8)
        mask = mask_generator("box", mask_len_range=(16, 17), image_size=64)(x[:1])[:, 0].unsqueeze(0)
        out["mask"] = mask
        ops["inpainting"] = (get_operator("inpainting", device="cpu"), {"mask": mask})
        noiser = get_noise("gaussian", sigma=0.05)
    for name, (op, kw) in ops.items():
        with torch.no_grad():
            m = op.forward(torch.rand(1, 3, 64, 64, generator=g) * 2 - 1, **kw)
            m = m + 0.05 * torch.randn(m.shape, generator=g)
            out[f"{name}_measurement"] = m
            out[f"{name}_ortho"] = op.ortho_project(x, **kw)
            out[f"{name}_project"] = op.project(x, m, **kw)
            if not kw:   # Projection.conditioning forwards no kwargs (no mask)
                with _ref.quiet():
                    cond = get_conditioning_method("projection", op, noiser)
                out[f"{name}_projection_cond"] = cond.conditioning(x_t=x.clone(), noisy_measurement=m)
    save("project.npz", **out)


def gen_semantic_on():
    """ps_semantic WITH the semantic term inside the reference's base loop (condition_methods.py:145-195): the
    embedder / guidance embeddings are injected as attributes of the reference's own object (its constructor would load
    facenet, which is external) — nothing else of the reference is touched.  anneal_factor ≠ 1 exercises s_t(t)."""
    torch.manual_seed(23)
    np.random.seed(23)
    model = TinyEps(seed=23)
    emb = SemEmbedder(seed=9)
    guid = torch.randn(1, 2, 16, generator=torch.Generator().manual_seed(1))
    with _ref.quiet():
        op = get_operator("gaussian_blur", kernel_size=61, intensity=3.0, device="cpu")
        noiser = get_noise("gaussian", sigma=0.05)
        cond = get_conditioning_method("ps_semantic", op, noiser, scale=0.3, sem_guid_scale=0.5, anneal_factor=2.0)
    cond.resnet, cond.guid_image_emb, cond.n_guid_images = emb, guid, guid.shape[1]
    s = sampler("ddpm", "4")
    x_true = torch.rand(1, 3, 32, 32) * 2 - 1
    y = noiser(op.forward(x_true)).detach()
    x_start = torch.randn(3, 3, 32, 32)
    steps = []

    def cond_fn(**k):
        res = cond.conditioning(**k)
        steps.append({"x_prev": k["x_prev"].detach().clone(), "grad": res[0].detach().clone(), "dist": res[1].detach().clone(),
                      "sem": res[2].detach().clone()})
        return res
    with Recorder() as rec, _ref.quiet():
        final, md, sd = s.p_sample_loop(model=model, x_start=x_start.clone(), measurement=y, measurement_cond_fn=cond_fn,
                                        record=False, save_root=None)
    out = {"x_start": x_start, "y": y, "final": final.detach(), "final_dist": md.detach(), "final_sem": sd.detach(),
           "guid": guid, "n_steps": np.array(s.num_timesteps)}
    for i, r in enumerate(rec.randn):
        out[f"randn_{i}"] = r
    for i, st in enumerate(steps):
        for k2, v in st.items():
            out[f"step{i}_{k2}"] = v
    save("trace_ddpm_ps_semantic_on_gblur.npz", **out)


def gen_full_chain():
    """BASELINE config 1 end to end: Gaussian deblur, FFHQ-256 UNet (seeded random weights), ONE image, the full 50-step
    respaced chain through the reference's own p_sample_loop on the CPU.  Noise draws come from a seeded tape
    (regenerated by the test); stored: the final reconstruction, the per-step distances and a few sub-sampled states."""
    torch.set_num_threads(os.cpu_count() or 4)
    model = seeded_unet("model_config.yaml", seed=1234, device="cpu")
    with _ref.quiet():
        op = get_operator("gaussian_blur", kernel_size=61, intensity=3.0, device="cpu")
        noiser = get_noise("gaussian", sigma=0.05)
        cond = get_conditioning_method("ps_semantic", op, noiser, scale=0.3, sem_guid_scale=0.0)   # = ps at HEAD (App. B)
    g = torch.Generator().manual_seed(501)
    x_true = torch.rand(1, 3, 256, 256, generator=g) * 2 - 1
    y = (op.forward(x_true) + 0.05 * torch.randn(1, 3, 256, 256, generator=g)).detach()
    x_start = seeded_randn(5100, 0, (1, 3, 256, 256))
    s = sampler("ddpm", "50")
    dists, states = [], {}

    def cond_fn(**k):
        res = cond.conditioning(**k)
        dists.append(float(res[1]))
        n = len(dists)
        if n in (1, 10, 25, 40):
            states[n] = k["x_prev"].detach()[..., ::4, ::4].clone()
        return res
    with SeededDraws(5200) as draws, _ref.quiet():
        final, md, _ = s.p_sample_loop(model=model, x_start=x_start.clone(), measurement=y, measurement_cond_fn=cond_fn,
                                       record=False, save_root=None)
    out = {"y": y, "final": final.detach(), "dists": np.array(dists, np.float32), "n_draws": np.array(draws.calls),
           "x_start_sum": tensor_checksum(x_start), "weights_sum": tensor_checksum(torch.cat([p.detach().reshape(-1)[:4096]
                                                                                            for p in model.parameters()])),
           "eps_probe": model(x_start, torch.tensor([999.0])).detach()[..., ::8, ::8]}
    for n, v in states.items():
        out[f"x_prev_at_call{n}_sub"] = v
    save("full_chain_c1.npz", **out)


if __name__ == "__main__":
    only = sys.argv[1:]
    if only:                     # e.g. `python oracle/make_golden.py gen_phase256 gen_full_chain`
        for fn_name in only:
            globals()[fn_name]()
        sys.exit(0)
    gen_psnr()
    gen_var_types()
    gen_mean_types()
    gen_resample_update()
    gen_schedule()
    gen_resizer()
    gen_operators()
    gen_multinomial()
    gen_masks()
    # the combinations that run at HEAD …
    run_trace("ddpm_ps_semantic_gblur", "ddpm", "4", "ps_semantic", dict(scale=0.3, sem_guid_scale=0.0),
              "gaussian_blur", dict(kernel_size=61, intensity=3.0), n=2, size=32, seed=11)
    # DiffStateGrad: projection at idx 2 and 0 of a 4-step chain (the loop prints; it is silenced by _ref.quiet)
    run_trace("ddpm_ps_semantic_gblur_dsg", "ddpm", "4", "ps_semantic", dict(scale=0.3, sem_guid_scale=0.0),
              "gaussian_blur", dict(kernel_size=61, intensity=3.0), n=2, size=32, seed=17,
              loop_kwargs=dict(project=True, period=2))
    run_trace("ttc_ddim_mcg_sr", "ttc_ddim", "12", "mcg", dict(scale=0.5),
              "super_resolution", dict(in_shape=(1, 3, 32, 32), scale_factor=4), n=4, size=32, seed=12)
    run_trace("search_ddpm_gblur", "search_ddpm", "4", "ps", dict(scale=0.3),
              "gaussian_blur", dict(kernel_size=61, intensity=3.0), n=4, size=32, seed=13)
    # … and upstream-arity loops assembled from the reference's own classes (ps / ps_anneal zero the image
    # inside HEAD's base loop, App. B)
    np.random.seed(8)
    m = mask_generator("random", mask_prob_range=(0.3, 0.7), image_size=32)(torch.zeros(1, 3, 32, 32))[:, 0].unsqueeze(0)
    run_trace("ddpm_ps_inpaint", "ddpm", "4", "ps", dict(scale=0.5), "inpainting", {}, n=2, size=32, seed=14,
              use_loop=False, mask=m)
    run_trace("ddpm_ps_anneal_phase", "ddpm", "3", "ps_anneal", dict(scale=1.0), "phase_retrieval",
              dict(oversample=2.0), n=2, size=64, seed=15, use_loop=False, anneal=True)
    run_trace("ddim_ps_motion", "ddim", "3", "ps", dict(scale=0.3), "motion_blur", dict(kernel_size=61, intensity=0.5),
              n=2, size=64, seed=16, use_loop=False)
    # Poisson noise model: the other branch of grad_and_value (condition_methods.py:50-55)
    run_trace("ddpm_ps_poisson_gblur", "ddpm", "4", "ps", dict(scale=0.3), "gaussian_blur",
              dict(kernel_size=61, intensity=3.0), n=3, size=32, seed=18, use_loop=False,
              noise=dict(name="poisson", rate=1.0))
    # dynamic thresholding: x̂₀ = clip(pre·quantile(|pre|, 0.95)) over the whole batch, differentiated through the quantile
    run_trace("ddpm_ps_dynthresh_sr", "ddpm", "4", "ps", dict(scale=0.3), "super_resolution",
              dict(in_shape=(1, 3, 32, 32), scale_factor=4), n=3, size=32, seed=19, use_loop=False,
              diffusion=dict(dynamic_threshold=True))
