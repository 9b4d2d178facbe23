"""Best-of-N selection (best_of_n_simple.py:32-41): per image, the path with the smallest final measurement
distance ‖y − A(x)‖ among the first n+1 paths.  The reference does this offline in numpy from pathwise_*.npy
files; here the same rule is available on the host (tables for plots) and on the device (pick the particle)."""
from __future__ import annotations

import numpy as np
import torch


def best_paths(distances) -> np.ndarray:
    """(n_data, n_max) distances → (n_data, n_max) int: column n holds argmin over paths [0, n]."""
    d = np.asarray(distances)
    return np.stack([np.argmin(d[:, :n + 1], axis=1) for n in range(d.shape[1])], axis=1)


def best_of_n_curves(distances, **metrics):
    """Mean over images of each metric at the selected path, for every n — what best_of_n_simple.py saves."""
    d = np.asarray(distances)
    sel = best_paths(d)
    rows = np.arange(d.shape[0])[:, None]
    out = {"distances": d[rows, sel].mean(axis=0)}
    for name, m in metrics.items():
        out[name] = np.asarray(m)[rows, sel].mean(axis=0)
    return out


def select_best(particles: torch.Tensor, distances: torch.Tensor):
    """Device-side pick of the best particle of one image: (best particle (1,C,H,W), index tensor, its distance).
    First minimum wins, like np.argmin / torch.argmin."""
    from . import kernels
    best, cost = kernels.argmin(distances.contiguous())
    return kernels.gather_particles(particles, best), best, cost
