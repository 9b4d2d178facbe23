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


def golden(name):
    path = os.path.join(GOLDEN, name)
    return np.load(path, allow_pickle=False)


def psnr(a, b, peak=2.0):
    """PSNR for images in [−1,1] (peak-to-peak 2), compute_metrics.py:93-98 convention."""
    mse = float(np.mean((np.asarray(a, np.float64) - np.asarray(b, np.float64)) ** 2))
    return float("inf") if mse == 0 else 10.0 * np.log10(peak ** 2 / mse)
