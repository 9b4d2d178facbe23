"""The steps on either side of the sampling loop, as the reference's drivers do them
(sample_condition_batched_ttc.py:143-196, sample_condition_hyper.py:150-244) — on the device:

    y_n = noiser(operator.forward(ref_img))                      measurement synthesis      (:164-165)
    for every path group: x_start ~ N(0, I) → sample_fn(...)     the particle loop          (:179-181)
    per path: ‖y_n − A(sample)‖₂, PSNR(ref, sample)              what the drivers log / save (:183-196)
    pathwise_{psnr,distances}.npy  (n_data, n_paths)             the inputs of best_of_n*.py (best_of_n_simple.py:22-24)
    best-of-N: argmin of the final distance over the first n+1 paths            (best_of_n_simple.py:32-41)

The reference writes every path to a PNG, reloads it and computes the metrics on the CPU; here the distance and the
PSNR of all paths of a group come from two kernel launches each (operator residual / squared difference + the norm
finish) and only the (n_paths,) vectors travel to the host.  PNG / LPIPS / CSV output stays out of scope.
"""
from __future__ import annotations

import os

import numpy as np
import torch

from . import kernels
from ._lib import DpsError
from .best_of_n import best_of_n_curves, best_paths, select_best
from .operators import B200Operator


def synthesize_measurement(operator, noiser, ref_img: torch.Tensor, **op_kwargs):
    """(y, y_n) = (A(ref), noiser(A(ref))) — sample_condition_batched_ttc.py:153-165 (mask=… for inpainting)."""
    with torch.no_grad():
        y = operator.forward(ref_img, **op_kwargs)
        y_n = noiser(y)
    return y, y_n


def psnr(ref_img: torch.Tensor, samples: torch.Tensor) -> torch.Tensor:
    """Per-particle PSNR with the reference's convention, compute_psnr_manual (compute_metrics.py:93-98):
    20·log10(1/√mse) on the tensors as they are (range [−1, 1], peak taken as 1).  (n,) on the device."""
    if not samples.is_cuda:
        raise DpsError("psnr(): CUDA tensors only — dps_ttc_b200 has no CPU path")
    l2, _ = kernels.particle_sqdiff(samples.float(), ref_img.to(samples.device, torch.float32))
    mse = l2 * l2 / samples[0].numel()
    return 20.0 * torch.log10(1.0 / torch.sqrt(mse))


def measurement_distance(operator, samples: torch.Tensor, y_n: torch.Tensor, **op_kwargs) -> torch.Tensor:
    """Per-particle ‖y_n − A(sample)‖₂ (the distance best-of-N selects on, gaussian_diffusion.py:303): fused residual +
    norm for the B200 operators, autograd-free torch otherwise (an external operator such as bkse)."""
    with torch.no_grad():
        if isinstance(operator, B200Operator):
            _, partials, _ = operator.residual(samples.float().contiguous(), y=y_n, **op_kwargs)
            return kernels.particle_norms(partials)
        diff = y_n - operator.forward(samples, **op_kwargs)
        return torch.linalg.norm(diff.reshape(diff.shape[0], -1), dim=-1)


def run_paths(sample_fn, operator, ref_img: torch.Tensor, y_n: torch.Tensor, n_paths: int, batch_size: int,
              generator: torch.Generator | None = None, keep_samples: bool = True, **op_kwargs):
    """The path-group loop of the drivers (sample_condition_batched_ttc.py:179-196).  `sample_fn` is the reference's
    partial(sampler.p_sample_loop, model=…, measurement_cond_fn=…, …); whatever tuple it returns, its first element is
    the particle batch.  Returns {"samples": (n_paths, C, H, W) or None, "distances": (n_paths,), "psnr": (n_paths,)}
    — distances and PSNR on the device, in path order (group-major)."""
    if n_paths % batch_size:
        raise ValueError(f"n_paths={n_paths} must be a multiple of batch_size={batch_size}")  # the reference drops the tail
    _, C, H, W = ref_img.shape
    dev = ref_img.device
    samples, dists, psnrs = [], [], []
    for _ in range(n_paths // batch_size):
        x_start = torch.randn((batch_size, C, H, W), device=dev, generator=generator)
        out = sample_fn(x_start=x_start, measurement=y_n, record=False, save_root=None)
        sample = (out[0] if isinstance(out, tuple) else out).detach()
        dists.append(measurement_distance(operator, sample, y_n, **op_kwargs))
        psnrs.append(psnr(ref_img, sample))
        if keep_samples:
            samples.append(sample)
    return {"samples": torch.cat(samples) if keep_samples else None, "distances": torch.cat(dists),
            "psnr": torch.cat(psnrs)}


class PathwiseLog:
    """pathwise_{distances,psnr,…}.npy, shape (n_data, n_paths): what the drivers accumulate per image and path and what
    best_of_n.py / best_of_n_simple.py load (best_of_n_simple.py:22-24)."""

    def __init__(self, n_data: int, n_paths: int, metrics=("distances", "psnr")):
        self.tables = {m: np.zeros((n_data, n_paths)) for m in metrics}

    def record(self, img_idx: int, path_start: int = 0, **values):
        for name, v in values.items():
            v = v.detach().double().cpu().numpy() if torch.is_tensor(v) else np.asarray(v, dtype=np.float64)
            self.tables[name][img_idx, path_start:path_start + v.shape[0]] = v

    def save(self, out_dir: str):
        os.makedirs(out_dir, exist_ok=True)
        for name, t in self.tables.items():
            np.save(os.path.join(out_dir, f"pathwise_{name}.npy"), t)

    def best_of_n(self):
        """The curves best_of_n_simple.py saves: mean over images of each metric at the arg-min-distance path."""
        others = {k: v for k, v in self.tables.items() if k != "distances"}
        return best_of_n_curves(self.tables["distances"], **others)


def sample_and_select(sample_fn, operator, noiser, ref_img: torch.Tensor, n_paths: int, batch_size: int,
                      generator: torch.Generator | None = None, **op_kwargs):
    """One image end to end: synthesise the measurement, run all path groups, pick the best particle on the device.
    Returns (best particle (1,C,H,W), index tensor, results dict of run_paths + "y_n")."""
    _, y_n = synthesize_measurement(operator, noiser, ref_img, **op_kwargs)
    res = run_paths(sample_fn, operator, ref_img, y_n, n_paths, batch_size, generator=generator, **op_kwargs)
    best, idx, _ = select_best(res["samples"], res["distances"])
    res["y_n"] = y_n
    return best, idx, res


__all__ = ["synthesize_measurement", "psnr", "measurement_distance", "run_paths", "PathwiseLog", "sample_and_select",
           "best_paths", "best_of_n_curves", "select_best"]
