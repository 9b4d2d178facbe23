"""Round-2 kernels: the fused super-resolution guidance kernel (thread-block cluster, halo rows pushed over DSMEM), the posterior update
with the deferred guidance coefficient, and the in-kernel Philox noise — against the oracle and against the two-kernel
path they replace.  All through the C ABI."""
import numpy as np
import pytest
import torch

from helpers import CpuBridge, TinyEps, oracle_guided_step, psnr
from oracle import dps_oracle as O

pytestmark = pytest.mark.gpu
DEV = torch.device("cuda:0")
DIFF = dict(steps=1000, noise_schedule="linear", model_mean_type="epsilon", model_var_type="learned_range",
            dynamic_threshold=False, clip_denoised=True, rescale_timesteps=True)


@pytest.fixture(autouse=True)
def _exact_fp32():
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.deterministic = True
    yield


def _consts(idx):
    from dps_ttc_b200.schedule import Schedule, named_beta_schedule
    return Schedule(named_beta_schedule("linear", 1000)).consts(idx)


@pytest.mark.parametrize("factor,n,idx,clip", [(4, 1, 999, True), (4, 5, 500, True), (8, 3, 999, True), (4, 2, 10, False)])
def test_fused_sr_guidance_vs_two_kernels_and_oracle(factor, n, idx, clip):
    from dps_ttc_b200 import kernels, tables
    from dps_ttc_b200.kernels import OperatorPlan
    (fh, wh), (fw, ww), _ = tables.resizer_tables((1, 3, 256, 256), 1.0 / factor)
    plan = OperatorPlan.resize(fh, wh, fw, ww, 3, 256, 256, DEV)
    assert plan.guidance_partials == 3 * 8
    k = _consts(idx)
    gen = torch.Generator(DEV).manual_seed(100 * factor + n)
    x = torch.randn(n, 3, 256, 256, device=DEV, generator=gen) / k.c1
    o6 = torch.randn(n, 6, 256, 256, device=DEV, generator=gen) * 0.3 / max(k.c2, 1e-3)
    eps = o6[:, :3]
    m = 256 // factor
    y = torch.randn(1, 3, m, m, device=DEV, generator=gen)
    # two-kernel path
    r2, p2, _ = plan.forward(x, eps, k, clip, y, want_partials=True)
    g2 = torch.zeros(n, 6, 256, 256, device=DEV)
    plan.adjoint(r2, None, x, eps, k, clip, None, out=g2[:, :3])
    # fused path (into a channel-slice view like the sampler's cotangent buffer)
    g1 = torch.full((n, 6, 256, 256), float("nan"), device=DEV)
    g1[:, 3:] = 0
    p1, r1, _ = plan.guidance(x, eps, k, clip, y, out=g1[:, :3], want_r=True)
    assert torch.isfinite(g1).all()
    scale_r, scale_g = float(r2.abs().max()), float(g2.abs().max())
    assert float((r1 - r2).abs().max()) <= 2e-6 * max(1.0, scale_r)
    assert float((g1[:, :3] - g2[:, :3]).abs().max()) <= 5e-6 * max(1.0, scale_g)
    n1, n2 = kernels.particle_norms(p1, want_l1=True), kernels.particle_norms(p2, want_l1=True)
    assert float((n1[0] - n2[0]).abs().max()) <= 1e-5 * float(n2[0].max())
    assert float((n1[1] - n2[1]).abs().max()) <= 1e-5 * float(n2[1].max())
    # oracle
    xn, en = x.cpu().numpy(), eps.cpu().numpy()
    x0, pre = O.x0_from_eps(xn, en, dict(c1=np.float32(k.c1), c2=np.float32(k.c2)), clip)
    r_ref = y.cpu().numpy() - O.resize_forward(x0, 1.0 / factor)
    g_ref = O.resize_adjoint(r_ref, 1.0 / factor, 256, 256)
    if clip:
        g_ref = g_ref * ((pre >= -1) & (pre <= 1))
    assert np.abs(r1.cpu().numpy() - r_ref).max() <= 5e-6 * max(1.0, np.abs(r_ref).max())
    assert np.abs(g1[:, :3].cpu().numpy() - g_ref).max() <= 1e-5 * max(1.0, np.abs(g_ref).max())
    # without r_out the residual never leaves the chip; same cotangent and partial sums, bit for bit
    g3 = torch.zeros(n, 3, 256, 256, device=DEV)
    p3, r3, _ = plan.guidance(x, eps, k, clip, y, out=g3)
    assert r3 is None and torch.equal(g3, g1[:, :3]) and torch.equal(p3, p1)


@pytest.mark.parametrize("factor", [4, 8])
def test_fused_sr_push_exchange_repeatable_under_load(factor):
    """The cluster kernel's neighbour exchange (st.async pushes completing on the destination CTA's mbarriers) under load:
    more clusters than the machine holds at once, launches interleaved on two streams so clusters of different launches
    share SMs, every repeat bit-identical and equal to the per-particle result of a one-particle launch (a plane's cluster
    must not depend on who runs beside it)."""
    from dps_ttc_b200 import tables
    from dps_ttc_b200.kernels import OperatorPlan
    (fh, wh), (fw, ww), _ = tables.resizer_tables((1, 3, 256, 256), 1.0 / factor)
    plan = OperatorPlan.resize(fh, wh, fw, ww, 3, 256, 256, DEV)
    k = _consts(700)
    n = 80                                        # 240 clusters = 1 920 CTAs; at most 4 per SM are resident
    gen = torch.Generator(DEV).manual_seed(77 + factor)
    x = torch.randn(n, 3, 256, 256, device=DEV, generator=gen) / k.c1
    eps = torch.randn(n, 3, 256, 256, device=DEV, generator=gen) * 0.3 / k.c2
    y = torch.randn(1, 3, 256 // factor, 256 // factor, device=DEV, generator=gen)
    g0 = torch.zeros(n, 3, 256, 256, device=DEV)
    p0, _, _ = plan.guidance(x, eps, k, True, y, out=g0)
    torch.cuda.synchronize()
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    outs = []
    for rep in range(6):
        for st in (s1, s2):
            with torch.cuda.stream(st):
                g = torch.full((n, 3, 256, 256), float("nan"), device=DEV)
                p, _, _ = plan.guidance(x, eps, k, True, y, out=g)
                outs.append((g, p))
    torch.cuda.synchronize()
    for g, p in outs:
        assert torch.equal(g, g0) and torch.equal(p, p0)
    for i in (0, 41, n - 1):                      # one particle alone = the same particle inside the batch
        gi = torch.zeros(1, 3, 256, 256, device=DEV)
        pi, _, _ = plan.guidance(x[i:i + 1], eps[i:i + 1], k, True, y, out=gi)
        assert torch.equal(gi[0], g0[i]) and torch.equal(pi[0], p0[i])


@pytest.mark.parametrize("sampler,mode", [("ddpm", 1), ("ddpm", 2), ("ddim", 1)])
def test_update_with_deferred_coefficient_equals_scaled_path(sampler, mode):
    from dps_ttc_b200 import kernels
    k = _consts(700)
    gen = torch.Generator(DEV).manual_seed(5)
    n = 5
    x = torch.randn(n, 3, 256, 256, device=DEV, generator=gen)
    o6 = torch.randn(n, 6, 256, 256, device=DEV, generator=gen) * 0.3
    z = torch.randn(n, 3, 256, 256, device=DEV, generator=gen)
    g_un = torch.randn(n, 3, 256, 256, device=DEV, generator=gen)
    vjp_un = torch.randn(n, 3, 256, 256, device=DEV, generator=gen) * 0.1
    partials = torch.rand(n, 24, 2, device=DEV, generator=gen) * 50
    scale = 0.37
    l2, coef = kernels.guidance_coef(partials, mode, scale)
    v = o6[:, 3:] if sampler == "ddpm" else None
    want, _, _ = kernels.posterior_update(sampler, x, o6[:, :3], v, z, k, g=(g_un * coef.view(-1, 1, 1, 1)).contiguous(),
                                          vjp=(vjp_un * coef.view(-1, 1, 1, 1)).contiguous())
    dist = torch.empty(n, device=DEV)
    got, _, _ = kernels.posterior_update(sampler, x, o6[:, :3], v, z, k, g=g_un, vjp=vjp_un, deferred=(partials, mode, scale, dist))
    assert torch.equal(dist, l2)                                             # same reduction order as dps_guidance_coef
    assert float((got - want).abs().max()) <= 2e-6 * max(1.0, float(want.abs().max()))
    # a particle with r = 0 gets a zero coefficient (torch's norm backward at 0), not a NaN
    partials[2] = 0
    got, _, _ = kernels.posterior_update(sampler, x, o6[:, :3], v, z, k, g=g_un, vjp=vjp_un, deferred=(partials, 1, scale, dist))
    plain, _, _ = kernels.posterior_update(sampler, x, o6[:, :3], v, z, k)
    assert torch.equal(got[2], plain[2]) and float(dist[2]) == 0.0 and torch.isfinite(got).all()


def test_philox_noise_in_the_update_kernel():
    from dps_ttc_b200 import kernels
    k = _consts(400)
    gen = torch.Generator(DEV).manual_seed(9)
    n = 4
    x = torch.randn(n, 3, 256, 256, device=DEV, generator=gen)
    o6 = torch.randn(n, 6, 256, 256, device=DEV, generator=gen) * 0.3
    seed, step, offset = (1 << 40) + 12345, 400, 6
    got, _, _ = kernels.posterior_update("ddpm", x, o6[:, :3], o6[:, 3:], None, k, philox=(seed, step, offset))
    # the same update with the oracle's restatement of the generator as an explicit z tensor
    z = np.stack([O.philox_normal(seed, step, offset + i, 3 * 256 * 256).reshape(3, 256, 256) for i in range(n)])
    want, _, _ = kernels.posterior_update("ddpm", x, o6[:, :3], o6[:, 3:], torch.from_numpy(z).to(DEV), k)
    assert float((got - want).abs().max()) <= 1e-4 * max(1.0, float(want.abs().max()))     # SFU lg2 / rsq / sin / cos vs libm
    # sharded == unsharded: particles 2..3 of the launch above = a launch of two particles at offset + 2
    part, _, _ = kernels.posterior_update("ddpm", x[2:], o6[2:, :3], o6[2:, 3:], None, k, philox=(seed, step, offset + 2))
    assert torch.equal(part, got[2:])
    # idx 0 adds no noise (gaussian_diffusion.py:473)
    k0 = _consts(0)
    a, _, _ = kernels.posterior_update("ddpm", x, o6[:, :3], o6[:, 3:], None, k0, philox=(seed, 0, 0))
    b, _, _ = kernels.posterior_update("ddpm", x, o6[:, :3], o6[:, 3:], None, k0)
    assert torch.equal(a, b)


@pytest.mark.parametrize("op_name,op_cfg,method,params,mode,scale_of", [
    ("super_resolution", dict(in_shape=(1, 3, 256, 256), scale_factor=4), "ps", dict(scale=0.01), "norm", lambda t, i: 0.01),
    ("gaussian_blur", dict(kernel_size=61, intensity=3.0), "ps", dict(scale=0.3), "norm", lambda t, i: 0.3),
    ("phase_retrieval", dict(oversample=2.0), "ps_anneal", dict(scale=1.0), "norm_sq", lambda t, i: t.at(i)["beta"] / 0.05 ** 2),
])
def test_deferred_step_vs_oracle_and_vs_round1_sequence(op_name, op_cfg, method, params, mode, scale_of):
    """The guided step with the deferred coefficient (default) against the oracle (per step, 1e-4) and against the
    residual → coefficient → cotangent sequence of round 1 (deferred_coef=False)."""
    from dps_ttc_b200.registry import get_conditioning_method, get_noise, get_operator
    from dps_ttc_b200.sampler import NoiseTape, create_sampler
    from dps_ttc_b200.tables import gaussian_kernel
    kern = gaussian_kernel(61, 3.0).astype(np.float32)
    fwd = {"super_resolution": lambda a: O.resize_forward(a, 0.25), "gaussian_blur": lambda a: O.blur_forward(a, kern),
           "phase_retrieval": lambda a: O.phase_forward(a, 64)}[op_name]
    adj = {"super_resolution": lambda u: O.resize_adjoint(u, 0.25, 256, 256), "gaussian_blur": lambda u: O.blur_adjoint(u, kern),
           "phase_retrieval": None}[op_name]
    nl = (lambda x0, u: O.phase_vjp(x0, u, 64)) if op_name == "phase_retrieval" else None
    rng = np.random.default_rng(3)
    x = rng.standard_normal((2, 3, 256, 256)).astype(np.float32)
    y = fwd((rng.random((1, 3, 256, 256)) * 2 - 1).astype(np.float32))
    y = (y + 0.05 * rng.standard_normal(y.shape)).astype(np.float32)
    z = rng.standard_normal(x.shape).astype(np.float32)
    idx = 999
    model_cpu = TinyEps(seed=3)
    outs = {}
    for deferred in (True, False):
        op = get_operator(op_name, device=DEV, **op_cfg)
        cond = get_conditioning_method(method, op, get_noise("gaussian", sigma=0.05), **params)
        s = create_sampler(sampler="ddpm", **DIFF)
        s.deferred_coef, s.parity_rng = deferred, False
        s.noise = NoiseTape(z={idx: torch.from_numpy(z)})
        img, dist, _ = s.p_sample_loop(model=CpuBridge(model_cpu), x_start=torch.from_numpy(x).to(DEV),
                                       measurement=torch.from_numpy(y).to(DEV), measurement_cond_fn=cond.conditioning,
                                       record=False, save_root=None, start_idx=idx, num_steps=1)
        outs[deferred] = (img.cpu().numpy(), dist.cpu().numpy())
    tab = O.Tables(1000)
    ref, norm, _ = oracle_guided_step(O, model_cpu, tab, x, idx, y, fwd, adj, z, mode, scale_of(tab, idx), nonlinear_vjp=nl)
    for deferred, (img, dist) in outs.items():
        assert np.abs(img - ref).max() <= 1e-4 * max(1.0, np.abs(ref).max()), f"deferred={deferred}"
        assert np.abs(dist - norm).max() <= 1e-5 * norm.max(), f"deferred={deferred}"
    assert psnr(outs[True][0], outs[False][0]) >= 100.0


@pytest.mark.parametrize("size", [256, 128, 64])
@pytest.mark.parametrize("n,idx,clip", [(1, 999, True), (4, 500, True), (2, 10, False)])
def test_fused_phase_guidance_vs_two_kernels_and_oracle(n, idx, clip, size):
    """Phase retrieval guidance as rows → fused columns (both transforms, residual and cotangent on chip) → rows, against
    the forward + adjoint kernels it replaces and against the oracle's |FFT| / Jᵀ restatement."""
    from dps_ttc_b200 import kernels
    from dps_ttc_b200.kernels import OperatorPlan
    plan = OperatorPlan.phase(64, 3, size, size, DEV)
    assert plan.guidance_partials == plan.partials_per_particle > 0
    k = _consts(idx)
    gen = torch.Generator(DEV).manual_seed(31 + n)
    x = torch.randn(n, 3, size, size, device=DEV, generator=gen) / k.c1
    o6 = torch.randn(n, 6, size, size, device=DEV, generator=gen) * 0.3 / max(k.c2, 1e-3)
    eps = o6[:, :3]
    y = torch.rand(1, 3, size + 128, size + 128, device=DEV, generator=gen) * 1.5
    r2, p2, aux2 = plan.forward(x, eps, k, clip, y, want_partials=True)
    g2 = torch.zeros(n, 6, size, size, device=DEV)
    plan.adjoint(r2, None, x, eps, k, clip, None, out=g2[:, :3], aux=aux2)
    g1 = torch.full((n, 6, size, size), float("nan"), device=DEV)
    g1[:, 3:] = 0
    p1, r1, _ = plan.guidance(x, eps, k, clip, y, out=g1[:, :3], want_r=True)
    assert torch.isfinite(g1).all()
    assert float((r1 - r2).abs().max()) <= 1e-6 * max(1.0, float(r2.abs().max()))
    assert float((g1[:, :3] - g2[:, :3]).abs().max()) <= 2e-5 * max(1.0, float(g2.abs().max()))
    n1, n2 = kernels.particle_norms(p1, want_l1=True), kernels.particle_norms(p2, want_l1=True)
    assert float((n1[0] - n2[0]).abs().max()) <= 1e-5 * float(n2[0].max())
    assert float((n1[1] - n2[1]).abs().max()) <= 1e-5 * float(n2[1].max())
    if n <= 2:
        xn, en = x.cpu().numpy(), eps.cpu().numpy()
        x0, pre = O.x0_from_eps(xn, en, dict(c1=np.float32(k.c1), c2=np.float32(k.c2)), clip)
        r_ref = y.cpu().numpy() - O.phase_forward(x0, 64)
        g_ref = O.phase_vjp(x0, r_ref, 64)
        if clip:
            g_ref = g_ref * ((pre >= -1) & (pre <= 1))
        assert np.abs(r1.cpu().numpy() - r_ref).max() <= 2e-5 * max(1.0, np.abs(r_ref).max())
        assert np.abs(g1[:, :3].cpu().numpy() - g_ref).max() <= 5e-5 * max(1.0, np.abs(g_ref).max())
    g3 = torch.zeros(n, 3, size, size, device=DEV)
    p3, r3, _ = plan.guidance(x, eps, k, clip, y, out=g3)
    assert r3 is None and torch.equal(g3, g1[:, :3]) and torch.equal(p3, p1)


@pytest.mark.parametrize("n,idx,clip", [(1, 999, True), (5, 500, True), (3, 10, False)])
def test_fused_inpainting_guidance_vs_two_kernels(n, idx, clip):
    from dps_ttc_b200 import kernels
    from dps_ttc_b200.kernels import OperatorPlan
    rng = np.random.default_rng(5)
    mask = (rng.random((256, 256)) > 0.4).astype(np.float32)
    plan = OperatorPlan.inpainting(mask, 3, 256, 256, DEV)
    assert plan.guidance_partials == plan.partials_per_particle > 0
    k = _consts(idx)
    gen = torch.Generator(DEV).manual_seed(77 + n)
    x = torch.randn(n, 3, 256, 256, device=DEV, generator=gen) / k.c1
    o6 = torch.randn(n, 6, 256, 256, device=DEV, generator=gen) * 0.3 / max(k.c2, 1e-3)
    eps = o6[:, :3]
    y = torch.randn(1, 3, 256, 256, device=DEV, generator=gen)
    r2, p2, _ = plan.forward(x, eps, k, clip, y, want_partials=True)
    g2 = torch.zeros(n, 6, 256, 256, device=DEV)
    plan.adjoint(r2, None, x, eps, k, clip, None, out=g2[:, :3])
    g1 = torch.full((n, 6, 256, 256), float("nan"), device=DEV)
    p1, r1, _ = plan.guidance(x, eps, k, clip, y, out=g1[:, :3], want_r=True)
    assert torch.equal(r1, r2)                                              # same arithmetic
    assert float((p1 - p2).abs().max()) <= 1e-6 * float(p2.abs().max())     # same reduction tree, FMA contraction may differ
    assert torch.equal(g1[:, :3], g2[:, :3])
    g3 = torch.zeros(n, 3, 256, 256, device=DEV)
    p3, r3, _ = plan.guidance(x, eps, k, clip, y, out=g3)
    assert r3 is None and torch.equal(g3, g2[:, :3]) and torch.equal(p3, p1)


def _sep_kernels():
    from dps_ttc_b200.tables import gaussian_kernel
    rng = np.random.default_rng(21)
    v = rng.random(11).astype(np.float64) + 0.1          # asymmetric rank-1 kernel, radius 5: exercises the flipped taps
    h = rng.random(11).astype(np.float64) + 0.1          # and the folded border terms of the adjoint
    asym = np.outer(v / v.sum(), h / h.sum()).astype(np.float32)
    return {"gauss61": gaussian_kernel(61, 3.0).astype(np.float32), "asym11": asym,
            "gauss9": gaussian_kernel(9, 1.0).astype(np.float32)}


@pytest.mark.parametrize("kname,n,idx,clip", [("gauss61", 1, 999, True), ("gauss61", 5, 500, True), ("asym11", 3, 999, True),
                                               ("gauss9", 2, 10, False)])
def test_fused_separable_blur_guidance_vs_two_kernels_and_oracle(kname, n, idx, clip):
    from dps_ttc_b200 import kernels
    from dps_ttc_b200.kernels import OperatorPlan
    kern = _sep_kernels()[kname]
    plan = OperatorPlan.blur(kern, 3, 256, 256, DEV)
    assert plan.kind == "blur_separable" and plan.guidance_partials == 3 * 8
    k = _consts(idx)
    gen = torch.Generator(DEV).manual_seed(300 + n)
    x = torch.randn(n, 3, 256, 256, device=DEV, generator=gen) / k.c1
    o6 = torch.randn(n, 6, 256, 256, device=DEV, generator=gen) * 0.3 / max(k.c2, 1e-3)
    eps = o6[:, :3]
    y = torch.randn(1, 3, 256, 256, device=DEV, generator=gen)
    r2, p2, _ = plan.forward(x, eps, k, clip, y, want_partials=True)
    g2 = torch.zeros(n, 6, 256, 256, device=DEV)
    plan.adjoint(r2, None, x, eps, k, clip, None, out=g2[:, :3])
    g1 = torch.full((n, 6, 256, 256), float("nan"), device=DEV)
    g1[:, 3:] = 0
    p1, r1, _ = plan.guidance(x, eps, k, clip, y, out=g1[:, :3], want_r=True)
    assert torch.isfinite(g1).all()
    assert float((r1 - r2).abs().max()) <= 2e-6 * max(1.0, float(r2.abs().max()))
    assert float((g1[:, :3] - g2[:, :3]).abs().max()) <= 5e-6 * max(1.0, float(g2.abs().max()))
    n1, n2 = kernels.particle_norms(p1, want_l1=True), kernels.particle_norms(p2, want_l1=True)
    assert float((n1[0] - n2[0]).abs().max()) <= 1e-5 * float(n2[0].max())
    assert float((n1[1] - n2[1]).abs().max()) <= 1e-5 * float(n2[1].max())
    if n <= 3:
        xn, en = x.cpu().numpy(), eps.cpu().numpy()
        x0, pre = O.x0_from_eps(xn, en, dict(c1=np.float32(k.c1), c2=np.float32(k.c2)), clip)
        r_ref = y.cpu().numpy() - O.blur_forward(x0, kern)
        g_ref = O.blur_adjoint(r_ref, kern)
        if clip:
            g_ref = g_ref * ((pre >= -1) & (pre <= 1))
        assert np.abs(r1.cpu().numpy() - r_ref).max() <= 5e-6 * max(1.0, np.abs(r_ref).max())
        assert np.abs(g1[:, :3].cpu().numpy() - g_ref).max() <= 1e-5 * max(1.0, np.abs(g_ref).max())
    g3 = torch.zeros(n, 3, 256, 256, device=DEV)
    p3, r3, _ = plan.guidance(x, eps, k, clip, y, out=g3)
    assert r3 is None and torch.equal(g3, g1[:, :3]) and torch.equal(p3, p1)
