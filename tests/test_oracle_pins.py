"""Pins the oracle (oracle/dps_oracle.py) to fixtures produced by running the REFERENCE itself
(oracle/make_golden.py → tests/golden/*.npz).  CPU only.  Tolerances are written per check."""
import numpy as np
import pytest
import torch

from helpers import TinyEps, golden
from oracle import dps_oracle as O


# ------------------------------------------------------------------------------------------------
def test_schedule_tables_match_reference():
    g = golden("schedule.npz")
    for tag, resp in (("full", None), ("r50", "50"), ("r12", "12")):
        t = O.Tables(1000, resp)
        assert list(g[f"{tag}_timestep_map"]) == t.timestep_map
        for ours, theirs in ((t.betas, "betas"), (t.acp, "alphas_cumprod"), (t.acp_prev, "alphas_cumprod_prev"),
                             (t.sqrt_recip, "sqrt_recip_alphas_cumprod"), (t.sqrt_recipm1, "sqrt_recipm1_alphas_cumprod"),
                             (t.coef1, "posterior_mean_coef1"), (t.coef2, "posterior_mean_coef2"),
                             (t.post_logvar_clipped, "posterior_log_variance_clipped")):
            assert np.array_equal(ours, g[f"{tag}_{theirs}"]), (tag, theirs)   # fp64, bit-exact


def test_resizer_tables_match_reference():
    g = golden("resizer.npz")
    for n, s in ((256, 4), (256, 8), (64, 4), (32, 4)):
        w, fov = O.resizer_contributions(n, n // s, 1 / s)
        assert np.array_equal(fov.T, g[f"fov_{n}_{s}"])
        assert np.array_equal(w.T, g[f"w_{n}_{s}"])


OPS = {
    "gaussian_blur": lambda g: (lambda x: O.blur_forward(x, g["gaussian_kernel"].astype(np.float32)),
                                lambda u: O.blur_adjoint(u, g["gaussian_kernel"].astype(np.float32))),
    "motion_blur": lambda g: (lambda x: O.blur_forward(x, g["motion_kernel"].astype(np.float32)),
                              lambda u: O.blur_adjoint(u, g["motion_kernel"].astype(np.float32))),
    "super_resolution": lambda g: (lambda x: O.resize_forward(x, 0.25), lambda u: O.resize_adjoint(u, 0.25, 64, 64)),
    "inpainting": lambda g: (lambda x: O.inpaint_forward(x, g["mask"]), lambda u: O.inpaint_forward(u, g["mask"])),
}


@pytest.mark.parametrize("name", sorted(OPS))
def test_linear_operators_forward_and_gradient(name):
    g = golden("operators.npz")
    fwd, adj = OPS[name](g)
    x = g["x"]
    ax = fwd(x)
    assert np.abs(ax - g[f"{name}_Ax"]).max() <= 2e-6          # fp32 conv / gather summation order
    r = g[f"{name}_y"] - ax
    grad, norm = O.guidance_cotangent(r, adj, pre=None, mode="norm", scale=-1.0, clip=False)  # +Aᵀ(−r)/‖r‖ …
    grad = -grad                                               # d‖y−Ax‖/dx = −Aᵀr/‖r‖
    assert np.abs(norm - g[f"{name}_norm"]).max() / g[f"{name}_norm"].max() <= 1e-6
    assert np.abs(-grad - g[f"{name}_grad"]).max() <= 1e-6 or np.abs(grad - g[f"{name}_grad"]).max() <= 1e-6


def test_phase_retrieval_forward_and_vjp():
    g = golden("operators.npz")
    x = g["x"]
    amp = O.phase_forward(x, 64)
    assert np.abs(amp - g["phase_retrieval_Ax"]).max() <= 2e-5     # fp64 FFT here vs complex64 in torch
    r = g["phase_retrieval_y"] - amp
    norm, _ = O.particle_norms(r)
    grad = -O.phase_vjp(x, r / norm[:, None, None, None], 64)
    assert np.abs(grad - g["phase_retrieval_grad"]).max() <= 2e-5


def test_multinomial_restatement_is_bit_exact():
    g = golden("multinomial.npz")
    for i in range(4):
        w = g[f"w_{i}"]
        _, cdf, deg = O.weights_cdf(np.log(w.astype(np.float64)).astype(np.float32), linear=True)
        # the restatement must be exact given the SAME fp32 weights: rebuild the cdf from w itself
        cum = np.cumsum(w.astype(np.float32), dtype=np.float32)  # numpy cumsum is sequential in fp32
        cdf_w = (cum / cum[-1]).astype(np.float32)
        ids = O.search(cdf_w, g[f"u_{i}"])
        assert np.array_equal(ids, g[f"ids_{i}"])
        assert not deg


def test_multinomial_against_live_torch():
    """torch is on the GPU box too: 200 seeded trials of torch.multinomial (CPU) vs the restatement."""
    for trial in range(200):
        n = (2, 4, 8, 64, 256)[trial % 5]
        gen = torch.Generator().manual_seed(trial)
        d = torch.rand(n, generator=gen) * (50 if trial % 3 else 0.01) + 100
        logw = (-d / 100).numpy().astype(np.float32)
        w = torch.exp(torch.from_numpy(logw))
        torch.manual_seed(1000 + trial)
        ids = torch.multinomial(w, n, replacement=True).numpy()
        torch.manual_seed(1000 + trial)
        u = torch.rand(n, dtype=torch.float64).numpy()
        cum = np.zeros(n, np.float32)
        s = np.float32(0)
        for j in range(n):
            s = np.float32(s + w.numpy()[j])
            cum[j] = s
        assert np.array_equal(O.search((cum / s).astype(np.float32), u), ids)


# ------------------------------------------------------------------------------------------------
# step traces recorded from the reference's own loops
# ------------------------------------------------------------------------------------------------
def model_and_vjp(model, x, t_model):
    xt = torch.from_numpy(x).requires_grad_(True)
    out = model(xt, torch.tensor([t_model], dtype=torch.float32))

    def vjp(g_eps):
        g6 = torch.zeros_like(out)
        g6[:, :3] = torch.from_numpy(g_eps)
        return torch.autograd.grad(out, xt, g6, retain_graph=True)[0].numpy()
    return out.detach().numpy(), vjp


def test_trace_ddpm_ps_semantic_gaussian_blur():
    """HEAD-valid combination: base loop + ps_semantic(sem=0) → x' = sample − ∇(0.3‖r‖)."""
    g = golden("trace_ddpm_ps_semantic_gblur.npz")
    T = O.Tables(1000, "4")
    model = TinyEps(seed=11)
    from dps_ttc_b200.tables import gaussian_kernel
    kern = gaussian_kernel(61, 3.0).astype(np.float32)
    img = g["x_start"]
    y = g["y"]
    n_steps = int(g["n_steps"])
    for i, idx in enumerate(reversed(range(n_steps))):
        k = T.at(idx)
        assert np.abs(img - g[f"step{i}_x_prev"]).max() <= 1e-5 * max(1.0, np.abs(img).max())   # chained parity
        img = g[f"step{i}_x_prev"]                       # re-anchor: per-step parity
        out6, vjp = model_and_vjp(model, img, k["model_t"])
        z = g[f"randn_{2 * i}"]                           # recorded draws per step: z, then the q_sample noise
        sample, x0 = O.ddpm_sample(img, out6[:, :3], out6[:, 3:], z, k, idx)
        assert np.abs(x0 - g[f"step{i}_x0"]).max() == 0.0                 # bit-exact
        # exp() differs by an ulp between libms; TinyEps' v is not confined to [−1,1], so scale the tolerance
        assert np.abs(sample - g[f"step{i}_sample"]).max() <= 2e-6 * max(1.0, np.abs(sample).max())
        _, pre = O.x0_from_eps(img, out6[:, :3], k)
        r = y - O.blur_forward(x0, kern)
        gpre, norm = O.guidance_cotangent(r, lambda u: O.blur_adjoint(u, kern), pre, "norm", 0.3)
        grad = k["c1"] * gpre - k["c2"] * vjp(gpre)
        assert np.abs(norm - g[f"step{i}_dist"]).max() / norm.max() <= 1e-6
        assert np.abs(grad - g[f"step{i}_grad"]).max() <= 1e-5 * max(1.0, np.abs(grad).max())
        img = (sample - grad).astype(np.float32)
    assert np.abs(img - g["final"]).max() <= 1e-4


@pytest.mark.parametrize("mean_type", ["epsilon", "start_x", "previous_x"])
def test_mean_processors_restatement(mean_type):
    """posterior_mean_variance.py:45-129: every mean processor as x̂₀ = c1·x − c2·out with its own scalars; the
    guided x_{t-1} of `ps` through the chain rule c1·g − c2·VJP(g) (the reference's autograd)."""
    g = golden("mean_types.npz")
    T = O.Tables(1000)
    model = TinyEps(seed=53)
    from dps_ttc_b200.tables import gaussian_kernel
    kern = gaussian_kernel(61, 3.0).astype(np.float32)
    x, y = g["x"], g["y"]
    for idx in (999, 500, 1, 0):
        k = O.mean_consts(T, idx, mean_type)
        out6, vjp = model_and_vjp(model, x, k["model_t"])
        tag = f"{mean_type}_{idx}"
        sample, x0 = O.ddpm_sample(x, out6[:, :3], out6[:, 3:], g[f"{tag}_z"], k, idx)
        assert np.abs(x0 - g[f"{tag}_x0"]).max() == 0.0, tag                               # bit-exact
        assert np.abs(sample - g[f"{tag}_sample"]).max() <= 2e-6 * max(1.0, np.abs(sample).max()), tag
        if f"{tag}_next" not in g.files:
            continue                       # previous_x: the reference cannot run its guided step (see make_golden.py)
        _, pre = O.x0_from_eps(x, out6[:, :3], k)
        r = y - O.blur_forward(x0, kern)
        gpre, norm = O.guidance_cotangent(r, lambda u: O.blur_adjoint(u, kern), pre, "norm", 0.3)
        nxt = O.guided_update(sample, gpre, vjp(gpre), k)
        assert np.abs(norm - g[f"{tag}_dist"]).max() / norm.max() <= 1e-6, tag
        assert np.abs(nxt - g[f"{tag}_next"]).max() <= 1e-5 * max(1.0, np.abs(nxt).max()), tag


def test_trace_diffstategrad_projection():
    """Reference loop with project=True, period=2 on a 4-step chain: idx 2 and 0 subtract the projected gradient of
    particle 0 from every particle, idx 3 and 1 the plain per-particle gradient."""
    g = golden("trace_ddpm_ps_semantic_gblur_dsg.npz")
    n_steps = int(g["n_steps"])
    for i, idx in enumerate(reversed(range(n_steps))):
        sample, grad = g[f"step{i}_sample"], g[f"step{i}_grad"]
        nxt = g[f"step{i + 1}_x_prev"] if i + 1 < n_steps else g["final"]
        if idx % 2 == 0:
            mine, r = O.diffstategrad_update(sample, grad)
            assert 1 <= r <= 32
            assert np.abs(mine[1] - (sample[1] - (sample[0] - nxt[0]))).max() <= 1e-5     # particle 0's projection on particle 1
        else:
            mine = sample - grad
        assert np.abs(mine - nxt).max() <= 1e-5 * max(1.0, np.abs(nxt).max()), idx


def test_trace_ttc_ddim_resampling_indices():
    g = golden("trace_ttc_ddim_mcg_sr.npz")
    i = 0
    while f"mn_w_{i}" in g.files:
        w = g[f"mn_w_{i}"]
        cum = np.cumsum(w, dtype=np.float32)
        ids = O.search((cum / cum[-1]).astype(np.float32), g[f"mn_u_{i}"])
        assert np.array_equal(ids, g[f"mn_ids_{i}"])
        i += 1
    assert i >= 1


def test_trace_search_ddpm_greedy():
    g = golden("trace_search_ddpm_gblur.npz")
    final = g["final"]
    assert np.abs(final - final[:1]).max() == 0.0           # all particles collapse onto the best one (:633)


def test_best_of_n_rule():
    d = np.array([[3.0, 1.0, 2.0, 0.5], [1.0, 1.0, 0.2, 0.9]])
    assert np.array_equal(O.best_of_n(d), np.array([[0, 1, 1, 3], [0, 0, 2, 2]]))


def test_psnr_matches_reference_function():
    """oracle.psnr against compute_psnr_manual of the reference (compute_metrics.py:93-98), fixture by make_golden.gen_psnr."""
    g = golden("psnr_manual.npz")
    assert np.abs(O.psnr(g["real"], g["fake"]) - g["psnr"]).max() <= 5e-6


def test_trace_ddpm_ps_anneal_phase_retrieval():
    """Pins the oracle's `norm_sq` guidance mode and its phase-retrieval forward / VJP to the reference:
    ps_anneal (condition_methods.py:198-212: x_t −= β_t/σ² · ∇‖y − A(x̂₀)‖²) over the nonlinear |FFT| operator
    (measurements.py:179-189), 3 respaced steps of DDPM.p_sample + PosterorSamplingAnnealing.conditioning as recorded by
    oracle/make_golden.py (upstream-arity loop: beta_scale = betas[idx], anneal = 1)."""
    from helpers import oracle_guided_step
    g = golden("trace_ddpm_ps_anneal_phase.npz")
    T = O.Tables(1000, "3")
    model = TinyEps(seed=15)
    y = g["y"]
    pad = (y.shape[-1] - g["x_start"].shape[-1]) // 2
    n_steps = int(g["n_steps"])
    sigma = 0.05                                           # max(noiser.sigma, 0.05), :203
    img = g["x_start"]
    for i, idx in enumerate(reversed(range(n_steps))):
        assert np.abs(img - g[f"step{i}_x_prev"]).max() <= 1e-4 * max(1.0, np.abs(img).max())   # chained parity
        img = g[f"step{i}_x_prev"]                       # re-anchor: per-step parity
        scale = float(T.betas[idx]) / sigma ** 2
        nxt, norm, dbg = oracle_guided_step(O, model, T, img, idx, y, lambda a: O.phase_forward(a, pad), None,
                                            g[f"randn_{i}"], "norm_sq", scale,
                                            nonlinear_vjp=lambda x0, u: O.phase_vjp(x0, u, pad))
        assert np.abs(dbg["x0"] - g[f"step{i}_x0"]).max() == 0.0                               # bit-exact
        assert np.abs(norm - g[f"step{i}_dist"]).max() / norm.max() <= 2e-6
        assert np.abs(nxt - g[f"step{i}_x_t_out"]).max() <= 1e-4 * max(1.0, np.abs(nxt).max()), idx
        img = nxt
    assert np.abs(img - g["final"]).max() <= 1e-4 * max(1.0, np.abs(g["final"]).max())
    assert np.abs(norm - g["final_dist"]).max() / norm.max() <= 2e-6


def test_phase256_oracle_vs_reference():
    """The oracle's phase operator and VJP at 256² → 384² (the size the CUDA kernels are built for) against the reference
    (fixture inputs are regenerated from seeds; checksums guard the RNG)."""
    from helpers import tensor_checksum
    g = golden("phase256.npz")
    gen = torch.Generator().manual_seed(2560)
    x = torch.rand(1, 3, 256, 256, generator=gen) * 2 - 1
    y = torch.rand(1, 3, 384, 384, generator=gen) * 1.5
    assert np.allclose(tensor_checksum(x), g["x_sum"], rtol=1e-12) and np.allclose(tensor_checksum(y), g["y_sum"], rtol=1e-12)
    x, y = x.numpy(), y.numpy()
    amp = O.phase_forward(x, 64)
    assert np.abs(amp[..., ::3, ::3] - g["Ax_sub"]).max() <= 2e-5
    r = y - amp
    norm, _ = O.particle_norms(r)
    assert np.abs(norm - g["norm"]).max() / g["norm"].max() <= 1e-5      # the reference's norm is fp32 over a complex64 FFT
    grad = -O.phase_vjp(x, r / norm[:, None, None, None], 64)
    assert np.abs(grad - g["grad"]).max() <= 2e-5 * max(1.0, np.abs(g["grad"]).max())


@pytest.mark.parametrize("name", ["gaussian_blur", "motion_blur", "super_resolution", "inpainting"])
def test_projection_restatements_vs_reference(name):
    """§8f row 2: ortho_project / project / `projection` conditioning of the reference's classes (fixture: gen_project)."""
    g = golden("project.npz")
    x, m = g["x"], g[f"{name}_measurement"]
    if name == "super_resolution":
        fwd = lambda a: O.resize_forward(a, 0.25)  # noqa: E731
        ortho = O.ortho_project(x, fwd, lambda u: O.nearest_upsample(u, 4))
        proj = O.sr_project(x, m, 4)
    else:
        if name == "inpainting":
            fwd = lambda a: O.inpaint_forward(a, g["mask"])  # noqa: E731
        else:
            from dps_ttc_b200.tables import gaussian_kernel
            kern = (gaussian_kernel(61, 3.0) if name == "gaussian_blur" else g["motion_kernel"]).astype(np.float32)
            fwd = lambda a: O.blur_forward(a, kern)  # noqa: E731
        ortho = O.ortho_project(x, fwd)
        proj = O.project(x, m, fwd)
    assert np.abs(ortho - g[f"{name}_ortho"]).max() <= 2e-6
    assert np.abs(proj - g[f"{name}_project"]).max() <= 2e-6
    if f"{name}_projection_cond" in g.files:
        assert np.abs(proj - g[f"{name}_projection_cond"]).max() <= 2e-6     # Projection.conditioning = operator.project


def test_trace_ps_semantic_on_oracle():
    """ps_semantic with the semantic term on: the oracle's cotangent with the embedder's gradient as `extra`
    (g_pre = mask ⊙ (−ζ/‖r‖·Aᵀr + s_t·∂ℓ_sem/∂x̂₀)) against the reference's recorded gradient, step by step."""
    import math
    from helpers import SemEmbedder
    from dps_ttc_b200.tables import gaussian_kernel
    g = golden("trace_ddpm_ps_semantic_on_gblur.npz")
    T = O.Tables(1000, "4")
    model, emb = TinyEps(seed=23), SemEmbedder(seed=9)
    guid = torch.from_numpy(g["guid"])
    kern = gaussian_kernel(61, 3.0).astype(np.float32)
    y = g["y"]
    for i, idx in enumerate(reversed(range(4))):
        k = T.at(idx)
        img = g[f"step{i}_x_prev"]
        out6, vjp = model_and_vjp(model, img, k["model_t"])
        x0, pre = O.x0_from_eps(img, out6[:, :3], k)
        t = idx / 4
        s_t = 0.5 * (1 + (2.0 - 1) / (1 + math.exp(-10 * (0.3 - t))))            # condition_methods.py:155
        x0t = torch.from_numpy(x0).requires_grad_(True)
        d = torch.norm((emb(x0t).unsqueeze(1) - guid).reshape(x0.shape[0], -1), dim=-1) / guid.shape[1]
        (extra,) = torch.autograd.grad((s_t * d).sum(), x0t)
        assert np.abs(d.detach().numpy() - g[f"step{i}_sem"]).max() <= 1e-5
        r = y - O.blur_forward(x0, kern)
        gpre, norm = O.guidance_cotangent(r, lambda u: O.blur_adjoint(u, kern), pre, "norm", 0.3, extra=extra.numpy())
        grad = k["c1"] * gpre - k["c2"] * vjp(gpre)
        assert np.abs(norm - g[f"step{i}_dist"]).max() / norm.max() <= 1e-6
        assert np.abs(grad - g[f"step{i}_grad"]).max() <= 1e-5 * max(1.0, np.abs(g[f"step{i}_grad"]).max())


def test_philox_restatement_known_answers():
    """The oracle's Philox4x32-10 (device noise of the throughput mode) against the published Random123 known-answer vectors
    (kat_vectors: philox4x32 10 rounds)."""
    kat = [((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
           ((0xffffffff,) * 4, (0xffffffff,) * 2, (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
           ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0), (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1))]
    for ctr, key, want in kat:
        got = O.philox4x32_10(np.array([ctr], dtype=np.uint32), np.array(key, dtype=np.uint32))[0]
        assert tuple(int(v) for v in got) == want
    z = O.philox_normal(7, 999, 3, 3 * 256 * 256)
    assert abs(float(z.mean())) < 0.01 and abs(float(z.std()) - 1) < 0.01 and np.isfinite(z).all()
    assert not np.array_equal(z, O.philox_normal(7, 998, 3, 3 * 256 * 256))       # keyed by step …
    assert not np.array_equal(z, O.philox_normal(7, 999, 4, 3 * 256 * 256))       # … and by particle
