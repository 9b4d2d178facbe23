"""GPU edge cases and cross-checks: the reference's own operator fixtures (64×64, other code paths than the
256×256 fast paths), ragged particle counts, per-particle measurements, the semantic-guidance hook, error paths."""
import functools

import numpy as np
import pytest
import torch

from helpers import TinyEps, golden
from oracle import dps_oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
DIFF = dict(steps=1000, noise_schedule="linear", model_mean_type="epsilon", model_var_type="learned_range",
            dynamic_threshold=False, clip_denoised=True, rescale_timesteps=True)


def _ops(g):
    from dps_ttc_b200.registry import get_operator
    ops = {"gaussian_blur": (get_operator("gaussian_blur", kernel_size=61, intensity=3.0, device=DEV), {}),
           "super_resolution": (get_operator("super_resolution", in_shape=(1, 3, 64, 64), scale_factor=4, device=DEV), {}),
           "inpainting": (get_operator("inpainting", device=DEV), {"mask": torch.from_numpy(g["mask"]).to(DEV)})}
    mb = get_operator("motion_blur", kernel_size=61, intensity=0.5, device=DEV)
    mb.set_kernel(np.asarray(g["motion_kernel"], np.float32).T)       # set_kernel stores the transpose
    ops["motion_blur"] = (mb, {})
    return ops


@pytest.mark.parametrize("name", ["gaussian_blur", "super_resolution", "inpainting", "motion_blur"])
def test_operators_match_reference_fixtures(name):
    """A(x) and ∇ₓ‖y − A(x)‖ recorded from the REFERENCE's operators (autograd) at 64×64."""
    g = golden("operators.npz")
    op, kw = _ops(g)[name]
    x = torch.from_numpy(g["x"]).to(DEV).requires_grad_(True)
    y = torch.from_numpy(g[f"{name}_y"]).to(DEV)
    ax = op.forward(x, **kw)
    assert np.abs(ax.detach().cpu().numpy() - g[f"{name}_Ax"]).max() <= 2e-6
    norm = op.residual_norm(x, y, **kw)
    (grad,) = torch.autograd.grad(norm.sum(), x)
    assert np.abs(norm.detach().cpu().numpy() - g[f"{name}_norm"]).max() <= 1e-5 * g[f"{name}_norm"].max()
    assert np.abs(grad.cpu().numpy() - g[f"{name}_grad"]).max() <= 1e-6
    # autograd through forward() itself (adjoint kernel as backward) gives the same gradient
    x2 = torch.from_numpy(g["x"]).to(DEV).requires_grad_(True)
    d = y - op.forward(x2, **kw)
    (grad2,) = torch.autograd.grad(torch.linalg.norm(d.reshape(2, -1), dim=-1).sum(), x2)
    assert np.abs(grad2.cpu().numpy() - g[f"{name}_grad"]).max() <= 1e-6


def test_phase_retrieval_rejects_unsupported_size():
    from dps_ttc_b200._lib import DpsError
    from dps_ttc_b200.registry import get_operator
    op = get_operator("phase_retrieval", oversample=2.0, device=DEV)
    with pytest.raises(DpsError):
        op.forward(torch.zeros(1, 3, 96, 96, device=DEV))      # kernels exist for 256 → 384, 128 → 256 and 64 → 192, and say so


def test_phase_retrieval_matches_reference_fixture_at_64():
    """The reference's own phase-retrieval fixture (64² image → 192² measurement; oracle/make_golden.py gen_operators): A(x)
    and ∇ₓ‖y − A(x)‖ recorded through its autograd, now reaching the CUDA kernels directly (radices 8·8·3)."""
    from dps_ttc_b200.registry import get_operator
    g = golden("operators.npz")
    op = get_operator("phase_retrieval", oversample=2.0, device=DEV)
    x = torch.from_numpy(g["x"]).to(DEV).requires_grad_(True)
    y = torch.from_numpy(g["phase_retrieval_y"]).to(DEV)
    ax = op.forward(x)
    assert tuple(ax.shape[-2:]) == (192, 192)
    assert np.abs(ax.detach().cpu().numpy() - g["phase_retrieval_Ax"]).max() <= 2e-5     # complex64 FFT in torch vs the Stockham kernel
    norm = op.residual_norm(x, y)
    (grad,) = torch.autograd.grad(norm.sum(), x)
    assert np.abs(norm.detach().cpu().numpy() - g["phase_retrieval_norm"]).max() <= 1e-5 * g["phase_retrieval_norm"].max()
    assert np.abs(grad.cpu().numpy() - g["phase_retrieval_grad"]).max() <= 2e-5


@pytest.mark.parametrize("n", [1, 5, 257])
def test_ragged_particle_counts(n):
    from dps_ttc_b200 import kernels
    from dps_ttc_b200.schedule import Schedule, named_beta_schedule
    rng = np.random.default_rng(n)
    k = Schedule(named_beta_schedule("linear", 1000)).consts(321)
    ko = O.Tables(1000).at(321)
    size = 64 if n > 16 else 256
    x = rng.standard_normal((n, 3, size, size)).astype(np.float32)
    o6 = rng.standard_normal((n, 6, size, size)).astype(np.float32)
    z = rng.standard_normal(x.shape).astype(np.float32)
    t6 = torch.from_numpy(o6).to(DEV)
    xn, s, x0 = kernels.posterior_update("ddpm", torch.from_numpy(x).to(DEV), t6[:, :3], t6[:, 3:], torch.from_numpy(z).to(DEV),
                                         k, want_sample=True, want_x0=True)
    s_ref, x0_ref = O.ddpm_sample(x, o6[:, :3], o6[:, 3:], z, ko, 321)
    assert np.abs(x0.cpu().numpy() - x0_ref).max() == 0.0
    assert np.abs(s.cpu().numpy() - s_ref).max() <= 1e-5
    ids = torch.from_numpy(rng.integers(0, n, n)).to(DEV)
    src = torch.from_numpy(x).to(DEV)
    assert torch.equal(kernels.gather_particles(src, ids), src[ids])
    d = torch.from_numpy(rng.random(n).astype(np.float32) * 50).to(DEV)
    w, cdf, lse, deg = kernels.weights_cdf(kernels.particle_logweights(d, tau=0.01), linear_mode=True)
    u = torch.from_numpy(rng.random(n)).to(DEV)
    a = kernels.ancestors(cdf, u, n, degenerate=deg).cpu().numpy()
    if n == 1:
        assert a.tolist() == [0] and int(deg.item()) == 1
    else:
        assert np.array_equal(a, O.search(cdf.cpu().numpy(), u.cpu().numpy()))


def test_per_particle_measurements():
    """y may be (N, …) instead of the broadcast (1, …): one measurement per particle."""
    from dps_ttc_b200 import kernels, tables
    from dps_ttc_b200.kernels import OperatorPlan
    rng = np.random.default_rng(3)
    kern = tables.gaussian_kernel(61, 3.0).astype(np.float32)
    plan = OperatorPlan.blur(kern, 3, 256, 256, DEV)
    x = rng.standard_normal((3, 3, 256, 256)).astype(np.float32)
    y = rng.standard_normal((3, 3, 256, 256)).astype(np.float32)
    r, partials, _ = plan.forward(torch.from_numpy(x).to(DEV), y=torch.from_numpy(y).to(DEV), want_partials=True)
    ref = y - O.blur_forward(x, kern)
    assert np.abs(r.cpu().numpy() - ref).max() <= 2e-6
    l2 = kernels.particle_norms(partials).cpu().numpy()
    assert np.abs(l2 - O.particle_norms(ref)[0]).max() <= 1e-5 * l2.max()


class Embedder(torch.nn.Module):
    """Stand-in for the face-embedding network of ps_semantic (external facenet in the reference)."""

    def __init__(self):
        super().__init__()
        g = torch.Generator().manual_seed(9)
        self.c = torch.nn.Conv2d(3, 4, 5, stride=4)
        self.l = torch.nn.Linear(4, 16)
        with torch.no_grad():
            for p in self.parameters():
                p.copy_(torch.randn(p.shape, generator=g) * 0.3)

    def forward(self, x):
        return self.l(torch.tanh(self.c(x)).mean(dim=(2, 3)))


def test_semantic_guidance_fused_equals_autograd_path():
    """ps_semantic with an embedder: the fused step (semantic gradient injected as `extra` into the cotangent
    kernel) must equal the generic autograd path through the same classes (reference-style conditioning)."""
    from dps_ttc_b200.registry import get_conditioning_method, get_noise, get_operator
    from dps_ttc_b200.sampler import NoiseTape, create_sampler
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    emb = Embedder().to(DEV)
    guid = torch.randn(1, 2, 16, generator=torch.Generator().manual_seed(1)).to(DEV)
    rng = np.random.default_rng(4)
    x = torch.from_numpy(rng.standard_normal((2, 3, 256, 256)).astype(np.float32)).to(DEV)
    y = torch.from_numpy(rng.standard_normal((1, 3, 64, 64)).astype(np.float32)).to(DEV)
    z = {999: torch.from_numpy(rng.standard_normal((2, 3, 256, 256)).astype(np.float32))}
    outs = []
    for fused in (True, False):
        op = get_operator("super_resolution", in_shape=(1, 3, 256, 256), scale_factor=4, device=DEV)
        cond = get_conditioning_method("ps_semantic", op, get_noise("gaussian", sigma=0.05), scale=0.01,
                                       sem_guid_scale=0.5, anneal_factor=2.0, embedder=emb, guid_emb=guid)
        s = create_sampler(sampler="ddpm", **DIFF)
        s.noise, s.parity_rng = NoiseTape(z=z), False
        img, md, sd = s.p_sample_loop(model=TinyEps(seed=5).to(DEV), x_start=x, measurement=y,
                                      measurement_cond_fn=cond.conditioning, record=False, save_root=None, num_steps=1,
                                      fused=fused)
        outs.append((img, md, sd))
    (a, amd, asd), (b, bmd, bsd) = outs
    assert float((a - b).abs().max()) <= 1e-4 * max(1.0, float(b.abs().max()))
    assert float((amd - bmd).abs().max()) <= 1e-4 * float(bmd.abs().max())
    assert float((asd - bsd).abs().max()) <= 1e-5


def test_ttc_semantic_reweighting_changes_weights_not_code_path():
    """Config 5: the semantic term enters the log-weights (−τ(ℓ_meas + w·ℓ_sem)); kernel vs oracle."""
    from dps_ttc_b200 import kernels
    rng = np.random.default_rng(7)
    meas = rng.random(256).astype(np.float32) * 30 + 50
    sem = rng.random(256).astype(np.float32)
    logw = kernels.particle_logweights(torch.from_numpy(meas).to(DEV), torch.from_numpy(sem).to(DEV), tau=0.1,
                                       meas_scale=0.5, meas_pow=1, sem_scale=2.0, sem_pow=2)
    ref = O.logweights(meas, sem, 0.1, 0.5, 1, 2.0, 2)
    assert np.abs(logw.cpu().numpy() - ref).max() <= 2e-6
    w, cdf, lse, deg = kernels.weights_cdf(logw, linear_mode=False)
    wn, cdf_ref, _ = O.weights_cdf(logw.cpu().numpy(), linear=False)
    assert np.abs(w.cpu().numpy() - wn).max() <= 1e-7 and abs(float(w.sum()) - 1) <= 1e-5
    u0 = torch.tensor([0.37], dtype=torch.float64, device=DEV)
    a = kernels.ancestors(cdf, u0, 256, systematic=True, degenerate=deg).cpu().numpy()
    assert np.array_equal(a, O.ancestors_systematic(cdf.cpu().numpy(), 0.37, 256)) and np.all(np.diff(a) >= 0)
