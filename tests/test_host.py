"""CPU-side tests: the C-ABI library loads and exports every symbol the header declares, host tables are
bit-identical to the reference's (golden fixtures), registries / error behaviour mirror the reference,
and the product refuses to run without CUDA (no fallback)."""
import os
import re

import numpy as np
import pytest
import torch

from helpers import REPO, golden


def test_library_exports_every_declared_symbol():
    from dps_ttc_b200 import _lib
    header = open(os.path.join(REPO, "include", "dpsttc.h")).read()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    declared = set(re.findall(r"\b(dps_[a-z0-9_]+)\s*\(", header))
    assert declared, "no declarations parsed"
    handle = _lib.lib()
    for name in sorted(declared):
        assert hasattr(handle, name), f"{name} declared in dpsttc.h but not exported"
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    assert handle.dps_compiled_sm() == 100 and handle.dps_version() >= 100


def test_library_contains_only_sm100a_code():
    import subprocess
    from dps_ttc_b200 import _lib
    out = subprocess.run(["cuobjdump", "-lelf", _lib.LIB_PATH], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_\d+a?", out))
    assert archs == {"sm_100a"}, archs


def test_schedule_matches_reference_tables():
    from dps_ttc_b200.schedule import Schedule
    g = golden("schedule.npz")
    for tag, resp in (("full", ""), ("r50", "50"), ("r12", "12")):
        s = Schedule.from_config(1000, "linear", True, resp)
        assert s.timestep_map == list(g[f"{tag}_timestep_map"])
        for attr in ("betas", "alphas_cumprod", "alphas_cumprod_prev", "sqrt_recip_alphas_cumprod",
                     "sqrt_recipm1_alphas_cumprod", "posterior_mean_coef1", "posterior_mean_coef2",
                     "posterior_log_variance_clipped", "sqrt_alphas_cumprod", "sqrt_one_minus_alphas_cumprod"):
            assert np.array_equal(getattr(s, attr), g[f"{tag}_{attr}"]), (tag, attr)
    k = Schedule.from_config(1000, "linear", True, "").consts(999)
    assert abs(k.c1 - 157.41046) < 1e-4 and k.model_t == 999.0 and k.noise_on == 1     # SURVEY App. C probe 3
    assert Schedule.from_config(1000, "linear", True, "").consts(0).noise_on == 0


def test_resizer_tables_bit_identical():
    from dps_ttc_b200.tables import resizer_band
    g = golden("resizer.npz")
    for n, s in ((256, 4), (256, 8), (64, 4), (32, 4)):
        fov, w = resizer_band(n, n // s, 1 / s)
        assert np.array_equal(fov, g[f"fov_{n}_{s}"]) and np.array_equal(w, g[f"w_{n}_{s}"])


def test_inpainting_masks_bit_exact():
    from dps_ttc_b200.tables import MaskGenerator
    g = golden("masks.npz")
    img = np.zeros((1, 3, 256, 256), np.float32)
    for seed in (0, 8):
        np.random.seed(seed)
        m = MaskGenerator("random", mask_prob_range=(0.3, 0.7), image_size=256)(img)
        assert np.array_equal(m[0, 0].astype(np.uint8), g[f"random_{seed}"])
        assert np.array_equal(m[0, 0], m[0, 2])
        np.random.seed(seed)
        m = MaskGenerator("box", mask_len_range=(128, 129), image_size=256)(img)
        assert np.array_equal(m[0, 0].astype(np.uint8), g[f"box_{seed}"])


def test_gaussian_kernel_matches_reference():
    from dps_ttc_b200.tables import gaussian_kernel
    g = golden("operators.npz")
    assert np.array_equal(gaussian_kernel(61, 3.0), g["gaussian_kernel"])


def test_motion_kernel_is_a_seeded_sparse_path():
    from dps_ttc_b200.tables import motion_kernel
    np.random.seed(8)
    a = motion_kernel(61, 0.5)
    np.random.seed(8)
    b = motion_kernel(61, 0.5)
    assert np.array_equal(a, b) and abs(a.sum() - 1) < 1e-12 and 10 < (a != 0).sum() < 1200


def test_registries_mirror_the_reference():
    import dps_ttc_b200.conditioning  # noqa: F401
    import dps_ttc_b200.operators  # noqa: F401
    import dps_ttc_b200.sampler  # noqa: F401
    from dps_ttc_b200 import registry as R
    assert R.OPERATORS.names() == ["gaussian_blur", "inpainting", "motion_blur", "noise", "nonlinear_blur",
                                   "phase_retrieval", "super_resolution"]
    assert R.CONDITIONING.names() == ["mcg", "projection", "ps", "ps+", "ps_anneal", "ps_semantic", "vanilla"]
    assert R.SAMPLERS.names() == ["ddim", "ddpm", "search_ddpm", "ttc_ddim"]
    assert R.NOISES.names() == ["clean", "gaussian", "poisson"]
    with pytest.raises(NameError):
        R.get_operator("no_such_operator", device="cpu")
    with pytest.raises(NameError):
        R.register_operator("gaussian_blur")(object)
    noiser = R.get_noise("gaussian", sigma=0.05)
    assert noiser.__name__ == "gaussian" and noiser.sigma == 0.05


def test_sampler_construction_and_errors():
    from dps_ttc_b200.sampler import create_sampler
    cfg = dict(steps=1000, noise_schedule="linear", model_mean_type="epsilon", model_var_type="learned_range",
               dynamic_threshold=False, clip_denoised=True, rescale_timesteps=True)
    s = create_sampler(sampler="ddpm", timestep_respacing="50", **cfg)
    assert s.num_timesteps == 50 and s.timestep_map[:3] == [0, 20, 41] and s.timestep_map[-1] == 999
    sx = create_sampler(sampler="ddpm", **{**cfg, "model_mean_type": "start_x"})
    k = sx._consts(500)
    assert (k.c1, k.c2, k.mean_mode) == (0.0, -1.0, 0)                 # x̂₀ = model output
    px = create_sampler(sampler="ddpm", **{**cfg, "model_mean_type": "previous_x"})
    k, base = px._consts(500), px.schedule.consts(500)
    assert k.mean_mode == 1 and k.c2 == -np.float32(1.0 / px.posterior_mean_coef1[500]) and k.p1 == base.p1
    with pytest.raises(NameError):
        create_sampler(sampler="ddpm", **{**cfg, "model_mean_type": "nope"})
    with pytest.raises(NotImplementedError):
        create_sampler(sampler="ddim", **{**cfg, "model_mean_type": "previous_x"})
    assert create_sampler(sampler="ddpm", **{**cfg, "dynamic_threshold": True}).dynamic_threshold is True
    with pytest.raises(NotImplementedError):
        create_sampler(sampler="ddpm", **{**cfg, "dynamic_threshold": True, "model_mean_type": "start_x"})
    with pytest.raises(NameError):
        create_sampler(sampler="nope", **cfg)
    with pytest.raises(ValueError):
        create_sampler(sampler="ddpm", timestep_respacing="2000", **cfg)


def test_no_cpu_fallback():
    """The product path must fail loudly without CUDA — it never routes through the oracle or torch-CPU."""
    from dps_ttc_b200._lib import DpsError
    from dps_ttc_b200.registry import get_conditioning_method, get_noise, get_operator
    from dps_ttc_b200.sampler import create_sampler
    op = get_operator("gaussian_blur", kernel_size=61, intensity=3.0, device="cpu")
    with pytest.raises(DpsError):
        op.forward(torch.zeros(1, 3, 64, 64))
    cond = get_conditioning_method("ps", op, get_noise("gaussian", sigma=0.05), scale=0.3)
    s = create_sampler(sampler="ddpm", steps=1000, noise_schedule="linear", model_mean_type="epsilon",
                       model_var_type="learned_range", dynamic_threshold=False, clip_denoised=True,
                       rescale_timesteps=True)
    with pytest.raises(DpsError):
        s.p_sample_loop(model=lambda x, t: x, x_start=torch.zeros(1, 3, 64, 64), measurement=torch.zeros(1, 3, 64, 64),
                        measurement_cond_fn=cond.conditioning, record=False, save_root=None)
    src = open(os.path.join(REPO, "dps_ttc_b200", "sampler.py")).read() + open(os.path.join(REPO, "dps_ttc_b200", "kernels.py")).read()
    assert "oracle" not in src.replace("oracle/", "")


def test_inpainting_requires_mask():
    from dps_ttc_b200.registry import get_operator
    op = get_operator("inpainting", device="cpu")
    with pytest.raises(ValueError):
        op.forward(torch.zeros(1, 3, 8, 8))


def test_semantic_and_anneal_schedules():
    from dps_ttc_b200.schedule import anneal_factor, semantic_scale
    assert semantic_scale(0.7, 0.5, 1.0) == 0.5
    assert abs(anneal_factor(0.5) - 0.5) < 1e-12 and anneal_factor(1.0, amp=2.0) > 1.98


def test_pathwise_log_and_best_of_n_curves(tmp_path):
    """PathwiseLog writes what best_of_n_simple.py loads; its curves equal that script's loop (:32-41) restated literally."""
    from dps_ttc_b200.driver import PathwiseLog
    rng = np.random.default_rng(0)
    n_data, n_max = 3, 6
    dist, ps = rng.random((n_data, n_max)), rng.random((n_data, n_max)) * 30
    log = PathwiseLog(n_data, n_max)
    for i in range(n_data):
        log.record(i, 0, distances=torch.from_numpy(dist[i, :4]), psnr=ps[i, :4])
        log.record(i, 4, distances=dist[i, 4:], psnr=torch.from_numpy(ps[i, 4:]))
    log.save(str(tmp_path))
    d2 = np.load(os.path.join(tmp_path, "pathwise_distances.npy"))
    p2 = np.load(os.path.join(tmp_path, "pathwise_psnr.npy"))
    assert np.array_equal(d2, dist) and np.array_equal(p2, ps)
    psnr_best = np.zeros((n_data, n_max))
    for n in range(n_max):                                   # best_of_n_simple.py:32-41
        best = np.argmin(d2[:, :n + 1], axis=1)
        for img in range(n_data):
            psnr_best[img, n] += p2[img, best[img]]
    assert np.allclose(log.best_of_n()["psnr"], psnr_best.mean(axis=0))


def test_driver_refuses_cpu_tensors():
    from dps_ttc_b200._lib import DpsError
    from dps_ttc_b200.driver import psnr
    with pytest.raises(DpsError):
        psnr(torch.zeros(1, 3, 8, 8), torch.zeros(2, 3, 8, 8))


def test_dynamic_threshold_torch_restatement_matches_reference():
    """_process_xstart (used by the generic autograd path) against the reference's dynamic_thresholding + clamp
    (posterior_mean_variance.py:40-45, util/img_utils.py:237-249), recorded in the dynamic-threshold trace."""
    from dps_ttc_b200.sampler import create_sampler
    cfg = dict(steps=1000, noise_schedule="linear", model_mean_type="epsilon", model_var_type="learned_range",
               dynamic_threshold=True, clip_denoised=True, rescale_timesteps=True)
    s = create_sampler(sampler="ddpm", timestep_respacing="4", **cfg)
    pre = torch.randn(3, 3, 16, 16, generator=torch.Generator().manual_seed(0)) * 2
    want = torch.clip(pre * torch.quantile(pre.abs(), 0.95), -1.0, 1.0).clamp(-1, 1)
    assert torch.equal(s._process_xstart(pre), want)
