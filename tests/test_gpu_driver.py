"""SURVEY §8f row 3 on the GPU: measurement synthesis, per-path distance + PSNR, best-of-N pick — against the oracle."""
import functools

import numpy as np
import pytest
import torch

from helpers import TinyEps

pytestmark = pytest.mark.gpu


def _setup(dev):
    from dps_ttc_b200 import tables
    from dps_ttc_b200.registry import get_conditioning_method, get_noise, get_operator
    from dps_ttc_b200.sampler import create_sampler
    op = get_operator("gaussian_blur", kernel_size=61, intensity=3.0, device=dev)
    noiser = get_noise("gaussian", sigma=0.05)
    cond = get_conditioning_method("ps", op, noiser, scale=0.3)
    s = create_sampler(sampler="ddpm", steps=1000, noise_schedule="linear", model_mean_type="epsilon",
                       model_var_type="learned_range", dynamic_threshold=False, clip_denoised=True,
                       rescale_timesteps=True, timestep_respacing="3")
    kern = tables.gaussian_kernel(61, 3.0).astype(np.float32)
    return op, noiser, cond, s, kern


def test_psnr_and_distance_match_oracle():
    from dps_ttc_b200 import driver
    from oracle import dps_oracle as O
    dev = torch.device("cuda:0")
    op, noiser, _, _, kern = _setup(dev)
    g = torch.Generator().manual_seed(5)
    ref = torch.rand(1, 3, 64, 64, generator=g) * 2 - 1
    smp = ref + 0.2 * torch.randn(5, 3, 64, 64, generator=g)
    y = torch.from_numpy(O.blur_forward(ref.numpy(), kern)) + 0.05 * torch.randn(1, 3, 64, 64, generator=g)
    p = driver.psnr(ref.to(dev), smp.to(dev)).cpu().numpy()
    assert np.abs(p - O.psnr(ref.numpy(), smp.numpy())).max() <= 1e-4
    d = driver.measurement_distance(op, smp.to(dev), y.to(dev)).cpu().numpy()
    want = O.measurement_distance(y.numpy(), lambda a: O.blur_forward(a, kern), smp.numpy())
    assert np.abs(d - want).max() <= 1e-4 * np.abs(want).max()
    # odd particle sizes go through the same kernel (chw % 4 == 0 is the only requirement)
    a, b = torch.randn(3, 1, 6, 10, generator=g), torch.randn(3, 1, 6, 10, generator=g)
    from dps_ttc_b200 import kernels
    l2, l1 = kernels.particle_sqdiff(a.to(dev), b.to(dev), P=7)
    assert torch.allclose(l2.cpu(), (a - b).reshape(3, -1).norm(dim=1), rtol=1e-6)
    assert torch.allclose(l1.cpu(), (a - b).reshape(3, -1).abs().sum(dim=1), rtol=1e-6)


def test_run_paths_and_selection():
    """Two path groups of 3 through the public loop: distances / PSNR equal a per-path recomputation, the selected
    particle is the arg-min-distance path, and the pathwise tables feed best_of_n like the reference's .npy files."""
    from dps_ttc_b200 import driver
    from oracle import dps_oracle as O
    dev = torch.device("cuda:0")
    op, noiser, cond, s, kern = _setup(dev)
    torch.manual_seed(0)
    ref = (torch.rand(1, 3, 64, 64) * 2 - 1).to(dev)
    model = TinyEps(seed=3).to(dev)
    sample_fn = functools.partial(s.p_sample_loop, model=model, measurement_cond_fn=cond.conditioning)
    best, idx, res = driver.sample_and_select(sample_fn, op, noiser, ref, n_paths=6, batch_size=3,
                                              generator=torch.Generator(dev).manual_seed(1))
    smp = res["samples"].cpu().numpy()
    assert smp.shape == (6, 3, 64, 64) and np.isfinite(smp).all()
    want_d = O.measurement_distance(res["y_n"].cpu().numpy(), lambda a: O.blur_forward(a, kern), smp)
    assert np.abs(res["distances"].cpu().numpy() - want_d).max() <= 1e-4 * np.abs(want_d).max()
    assert np.abs(res["psnr"].cpu().numpy() - O.psnr(ref.cpu().numpy(), smp)).max() <= 1e-4
    assert int(idx) == int(np.argmin(res["distances"].cpu().numpy()))
    assert torch.equal(best[0], res["samples"][int(idx)])
    log = driver.PathwiseLog(1, 6)
    log.record(0, 0, distances=res["distances"], psnr=res["psnr"])
    assert np.array_equal(driver.best_paths(log.tables["distances"])[0], O.best_of_n(log.tables["distances"])[0])
