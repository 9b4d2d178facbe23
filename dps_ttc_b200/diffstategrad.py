"""DiffStateGrad projection of the guidance gradient (guided_diffusion/diffstategrad_utils.py,
hooked at gaussian_diffusion.py:240-255).

Every `period` steps the reference projects the gradient onto the leading singular subspaces of the CURRENT SAMPLE of
particle 0 — per channel, P(G) = U_r U_rᵀ G V_r V_rᵀ with an adaptive rank r — and subtracts that single projected
gradient from every particle (the projected tensor has batch 1 and broadcasts, :255).  Both quirks are kept: results
must equal the reference's on the same inputs.

The SVD (cuSOLVER) and the four 256×256×r products (cuBLAS) are library calls through torch — three small matrices
once every `period` steps; the graft's own kernels materialise the gradient (dps_guidance_grad) and apply the
projected one (dps_apply_gradient).
"""
from __future__ import annotations

import numpy as np
import torch

from . import kernels


def adaptive_rank(singular_values: np.ndarray, cutoff: float) -> int:
    """compute_rank_for_explained_variance (diffstategrad_utils.py:5-22) AS CALLED (:41): the list holds ONE (C, r)
    array, so the cumulative sum runs over the FLATTENED channel-major singular values and the `/ 3` divides a single
    searchsorted position."""
    sq = np.asarray(singular_values) ** 2
    cumulative = np.cumsum(sq) / np.sum(sq)
    return int((int(np.searchsorted(cumulative, cutoff)) + 1) / 3)


def project_gradient(sample: torch.Tensor, grad: torch.Tensor, var_cutoff: float = 0.99) -> torch.Tensor:
    """(N,C,H,W) sample, (N,C,H,W) gradient → (1,C,H,W) projected gradient of particle 0 (:37, :66-74)."""
    U, s, Vh = torch.linalg.svd(sample[0], full_matrices=False)
    r = adaptive_rank(s.detach().cpu().numpy(), var_cutoff)
    A, B = U[:, :, :r], Vh[:, :r, :]
    low = torch.matmul(A.permute(0, 2, 1), grad[0]) @ B.permute(0, 2, 1)
    return (torch.matmul(A, low) @ B).float().unsqueeze(0).contiguous()


def projected_update(sample: torch.Tensor, grad: torch.Tensor, var_cutoff: float = 0.99) -> torch.Tensor:
    """x_{t-1} = sample − P(grad[0]) for every particle."""
    return kernels.apply_gradient(sample, project_gradient(sample, grad, var_cutoff))
