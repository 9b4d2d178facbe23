"""Drop-in check: the reference's OWN, unmodified sampling loop (GaussianDiffusion.p_sample_loop at HEAD)
driving the B200 operator + conditioning classes through its plugin surface, on the GPU, must reproduce the trace
the reference produced with its own classes (tests/golden/trace_ddpm_ps_semantic_gblur.npz).

Needs a copy of the reference tree (baseline/_ref, staged by __graft_entry__.build(); never /root/reference at
run time on the GPU box) — skipped when it is not there."""
import numpy as np
import pytest
import torch

from helpers import CpuBridge, TinyEps, golden

pytestmark = pytest.mark.gpu

DIFF = dict(steps=1000, noise_schedule="linear", model_mean_type="epsilon", model_var_type="learned_range",
            dynamic_threshold=False, clip_denoised=True, rescale_timesteps=True)


@pytest.fixture()
def reference():
    from dps_ttc_b200 import _ref
    if _ref.reference_root() is None:
        pytest.skip("no staged reference tree (baseline/_ref)")
    _ref.ensure_reference()
    return _ref


def test_reference_loop_drives_b200_plugins(reference, monkeypatch):
    g = golden("trace_ddpm_ps_semantic_gblur.npz")
    with reference.quiet():
        import guided_diffusion.gaussian_diffusion as ref_gd
        ref_sampler_cls = ref_gd.__SAMPLER__["ddpm"]              # the reference's class, before any re-binding
        s = ref_gd.create_sampler(sampler="ddpm", timestep_respacing="4", **DIFF)
    assert type(s) is ref_sampler_cls and type(s).__module__ == "guided_diffusion.gaussian_diffusion"
    from dps_ttc_b200.registry import get_conditioning_method, get_noise, get_operator
    dev = torch.device("cuda:0")
    op = get_operator("gaussian_blur", kernel_size=61, intensity=3.0, device=dev)
    cond = get_conditioning_method("ps_semantic", op, get_noise("gaussian", sigma=0.05), scale=0.3, sem_guid_scale=0.0)
    # replay the recorded draws (the reference draws z then the q_sample noise every step, on its device)
    draws = [torch.from_numpy(g[f"randn_{i}"]).to(dev) for i in range(8)]
    it = iter(draws)
    monkeypatch.setattr(torch, "randn_like", lambda t, *a, **k: next(it).reshape(t.shape))
    from dps_ttc_b200 import _lib
    before = _lib.launch_count()
    with reference.quiet():
        img, dist, sem = s.p_sample_loop(model=CpuBridge(TinyEps(seed=11)), x_start=torch.from_numpy(g["x_start"]).to(dev),
                                         measurement=torch.from_numpy(g["y"]).to(dev),
                                         measurement_cond_fn=cond.conditioning, record=False, save_root=None)
    assert _lib.launch_count() - before >= 8          # forward + adjoint kernels ran inside the reference's autograd
    ref = g["final"]
    assert np.abs(img.detach().cpu().numpy() - ref).max() <= 1e-4 * max(1.0, np.abs(ref).max())
    assert np.abs(dist.detach().cpu().numpy() - g["step3_dist"]).max() <= 1e-5 * g["step3_dist"].max()


def test_install_into_reference_rebinds_names(reference):
    with reference.quiet():
        ref_meas, ref_cond, ref_gd = reference.install_into_reference()
        op = ref_meas.get_operator("super_resolution", in_shape=(1, 3, 256, 256), scale_factor=4, device="cuda:0")
        s = ref_gd.create_sampler(sampler="ddpm", timestep_respacing="", **DIFF)
    import dps_ttc_b200.operators as ops
    import dps_ttc_b200.sampler as smp
    assert isinstance(op, ops.SuperResolutionOperator) and isinstance(s, smp.DDPM)
    x = torch.rand(2, 3, 256, 256, device="cuda:0")
    assert tuple(op.forward(x).shape) == (2, 3, 64, 64)
