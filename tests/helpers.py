"""Shared test helpers: a tiny deterministic ε-model and small utilities.  Used by oracle/make_golden.py
(which runs the reference) and by the tests, so both sides see the same model."""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if REPO not in sys.path:
    sys.path.insert(0, REPO)
GOLDEN = os.path.join(REPO, "tests", "golden")


class TinyEps(torch.nn.Module):
    """(N,3,H,W), t → (N,6,H,W): two 3×3 convs and a time-dependent gain.  Stands in for the UNet where
    only the data path around it is under test (the real UNet has 93 M parameters)."""

    def __init__(self, seed=0, width=8):
        super().__init__()
        g = torch.Generator().manual_seed(seed)
        self.c1 = torch.nn.Conv2d(3, width, 3, padding=1)
        self.c2 = torch.nn.Conv2d(width, 6, 3, padding=1)
        with torch.no_grad():
            for p in self.parameters():
                p.copy_(torch.randn(p.shape, generator=g) * 0.2)

    def forward(self, x, t):
        gain = 1.0 + 0.001 * t.float().reshape(-1)[0]
        h = torch.tanh(self.c1(x))
        return self.c2(h) * gain


class SemEmbedder(torch.nn.Module):
    """Seeded stand-in for the face-embedding network of ps_semantic (the reference's is facenet_pytorch's
    InceptionResnetV1 with downloaded weights: external).  (N,3,H,W) → (N,16).  Shared by
    oracle/make_golden.py (injected into the REFERENCE's PosteriorSamplingSemanticGuid) and the tests."""

    def __init__(self, seed=9):
        super().__init__()
        g = torch.Generator().manual_seed(seed)
        self.c = torch.nn.Conv2d(3, 4, 5, stride=4)
        self.l = torch.nn.Linear(4, 16)
        with torch.no_grad():
            for p in self.parameters():
                p.copy_(torch.randn(p.shape, generator=g) * 0.3)

    def forward(self, x):
        return self.l(torch.tanh(self.c(x)).mean(dim=(2, 3)))


def seeded_randn(base, i, shape):
    """The i-th normal draw of a reproducible tape (regenerated from seeds on both sides instead of being stored)."""
    return torch.randn(tuple(shape), generator=torch.Generator().manual_seed(int(base) + int(i)))


def seeded_unet(config="model_config.yaml", seed=1234, device="cpu"):
    """The reference's UNet with reproducible random weights: torch's global generator is seeded before the module's
    default init, then the zero-initialised output convs are re-drawn under `seed` (dps_ttc_b200._ref.create_unet)."""
    from dps_ttc_b200 import _ref
    state = torch.get_rng_state()
    torch.manual_seed(seed)
    try:
        return _ref.create_unet(config, reinit_zero_seed=seed, device=device)
    finally:
        torch.set_rng_state(state)


def tensor_checksum(t):
    """(float64 sum, float64 sum of squares) — detects a drifted RNG / weight init between fixture and test."""
    t = torch.as_tensor(t).double()
    return np.array([float(t.sum()), float((t * t).sum())])


def golden(name):
    path = os.path.join(GOLDEN, name)
    return np.load(path, allow_pickle=False)


def psnr(a, b, peak=2.0):
    """PSNR for images in [−1,1] (peak-to-peak 2), compute_metrics.py:93-98 convention."""
    mse = float(np.mean((np.asarray(a, np.float64) - np.asarray(b, np.float64)) ** 2))
    return float("inf") if mse == 0 else 10.0 * np.log10(peak ** 2 / mse)


# ------------------------------------------------------------------------------------------------
# oracle trajectories (numpy oracle + torch-CPU model for ε and its VJP)
# ------------------------------------------------------------------------------------------------
def model_and_vjp(model, x, t_model):
    xt = torch.from_numpy(np.ascontiguousarray(x)).requires_grad_(True)
    out = model(xt, torch.tensor([t_model], dtype=torch.float32))

    def vjp(g_eps):
        g6 = torch.zeros_like(out)
        g6[:, :3] = torch.from_numpy(np.ascontiguousarray(g_eps))
        return torch.autograd.grad(out, xt, g6, retain_graph=True)[0].numpy()
    return out.detach().numpy(), vjp


def oracle_guided_step(O, model, tables, img, idx, y, fwd, adj, z, mode="norm", scale=0.3, sampler="ddpm",
                       extra=None, nonlinear_vjp=None):
    """One guided reverse step restated with the oracle.  fwd/adj: numpy operator and adjoint;
    nonlinear_vjp(x0, g): Jᵀg at x0 for nonlinear operators (replaces adj)."""
    k = tables.at(idx)
    out6, vjp = model_and_vjp(model, img, k["model_t"])
    eps, v = out6[:, :3], out6[:, 3:]
    if sampler == "ddpm":
        sample, x0 = O.ddpm_sample(img, eps, v, z, k, idx)
    else:
        sample, x0 = O.ddim_sample(img, eps, z, k, idx)
    _, pre = O.x0_from_eps(img, eps, k)
    r = y - fwd(x0)
    adjoint = adj if nonlinear_vjp is None else (lambda u: nonlinear_vjp(x0, u))
    gpre, norm = O.guidance_cotangent(r, adjoint, pre, mode, scale, extra)
    x_next = O.guided_update(sample, gpre, vjp(gpre), k)
    return x_next, norm, dict(x0=x0, sample=sample, r=r, gpre=gpre)


class _CpuBridgeFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, t, model):
        xc = x.detach().cpu().requires_grad_(True)
        with torch.enable_grad():
            out = model(xc, t.detach().cpu())
        ctx.xc, ctx.out = xc, out
        return out.detach().to(x.device)

    @staticmethod
    def backward(ctx, g):
        (gx,) = torch.autograd.grad(ctx.out, ctx.xc, g.cpu())
        return gx.to(g.device), None, None


class CpuBridge(torch.nn.Module):
    """Evaluates a CPU model (and its VJP) for CUDA inputs, so that ε and the VJP are bit-identical to what
    the reference / the oracle saw on the CPU.  Test infrastructure: it isolates the parity of the CUDA kernels
    from cuDNN-vs-CPU rounding of the stand-in model (which, amplified by c1 ≈ 157 at t ≈ T, can flip a
    clamp-mask bit)."""

    def __init__(self, model):
        super().__init__()
        self.model = model

    def forward(self, x, t):
        return _CpuBridgeFn.apply(x, t, self.model)
