"""What the graft replaces, measured on the SAME B200: the reference's own loop in torch eager on the GPU versus the
fused samplers, with a TRIVIAL ε-model (two pointwise ops) so that the time is the guidance / operator / update path
around the UNet and nothing else.  Per operator and particle count: µs per step, reference vs graft.

    python tools/eager_baseline.py [--n 8] [--steps 20]

reference arm : the unmodified classes from baseline/_ref (create_sampler('ddpm') + get_operator + ps_semantic(sem=0), the
                HEAD-valid spelling of ps inside the base loop) on cuda — ~430-650 ATen launches, 8 table
                uploads and ≥3 host syncs per step;
graft arm     : dps_ttc_b200 registries, same YAML values, the 4-kernel step.
Both loops run the same `steps`-step respaced chain; wall time between two synchronisations divided by the steps.
"""
from __future__ import annotations

import argparse
import functools
import json
import os
import sys
import time

import numpy as np
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)

DIFF = dict(steps=1000, noise_schedule="linear", model_mean_type="epsilon", model_var_type="learned_range",
            dynamic_threshold=False, clip_denoised=True, rescale_timesteps=True)
CASES = [
    ("gaussian_blur", dict(kernel_size=61, intensity=3.0), 0.3),
    ("motion_blur", dict(kernel_size=61, intensity=0.5), 0.3),
    ("super_resolution", dict(in_shape=(1, 3, 256, 256), scale_factor=4), 0.01),
    ("inpainting", {}, 0.5),
    ("phase_retrieval", dict(oversample=2.0), 1.0),
]


class TrivialEps(torch.nn.Module):
    """(N,3,H,W) → (N,6,H,W) with two pointwise ops: the cheapest differentiable stand-in for the UNet."""

    def forward(self, x, t):
        return torch.cat([0.3 * x, 0.1 * x], dim=1)


def timed(fn, reps=2):
    best = None
    for _ in range(reps):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        fn()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    return best


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=8)
    ap.add_argument("--steps", type=int, default=20)
    a = ap.parse_args()
    dev = torch.device("cuda:0")
    from dps_ttc_b200 import _ref
    from dps_ttc_b200 import registry as R
    from dps_ttc_b200.sampler import create_sampler as b200_sampler
    from dps_ttc_b200.tables import MaskGenerator
    if _ref.reference_root() is None:
        raise SystemExit("reference tree not staged (baseline/_ref): run __graft_entry__.build() where /root/reference exists")
    _ref.ensure_reference()
    with _ref.quiet():
        from guided_diffusion.condition_methods import get_conditioning_method as ref_cond
        from guided_diffusion.gaussian_diffusion import create_sampler as ref_sampler
        from guided_diffusion.measurements import get_noise as ref_noise, get_operator as ref_op
    model = TrivialEps().to(dev)
    g = torch.Generator().manual_seed(1)
    x_true = (torch.rand(1, 3, 256, 256, generator=g) * 2 - 1).to(dev)
    x_start = torch.randn(a.n, 3, 256, 256, generator=g).to(dev)
    np.random.seed(8)
    mask = torch.from_numpy(MaskGenerator("random", mask_prob_range=(0.3, 0.7), image_size=256)(np.zeros((1, 3, 256, 256)))[:, :1]).to(dev)
    for name, cfg, scale in CASES:
        kw = {"mask": mask} if name == "inpainting" else {}
        # ---- reference, torch eager on the GPU
        np.random.seed(8)
        with _ref.quiet():
            op_r = ref_op(name, device=dev, **cfg)
            cond_r = ref_cond("ps_semantic", op_r, ref_noise("gaussian", sigma=0.05), scale=scale, sem_guid_scale=0.0)
            s_r = ref_sampler(sampler="ddpm", timestep_respacing=str(a.steps), **DIFF)
            y = op_r.forward(x_true, **kw).detach()
        fn_r = functools.partial(cond_r.conditioning, **kw) if kw else cond_r.conditioning

        def run_ref():
            with _ref.quiet():
                s_r.p_sample_loop(model=model, x_start=x_start.clone(), measurement=y, measurement_cond_fn=fn_r,
                                  record=False, save_root=None)
        # ---- graft
        np.random.seed(8)
        op_b = R.get_operator(name, device=dev, **cfg)
        cond_b = R.get_conditioning_method("ps", op_b, R.get_noise("gaussian", sigma=0.05), scale=scale)
        s_b = b200_sampler(sampler="ddpm", timestep_respacing=str(a.steps), **DIFF)
        s_b.parity_rng = False
        fn_b = functools.partial(cond_b.conditioning, **kw) if kw else cond_b.conditioning

        def run_b200():
            s_b.p_sample_loop(model=model, x_start=x_start.clone(), measurement=y, measurement_cond_fn=fn_b, record=False,
                              save_root=None)
        run_ref(); run_b200()                                   # warm-up (cuDNN plans, first-launch attributes)
        t_ref, t_b = timed(run_ref), timed(run_b200)
        print(json.dumps({"operator": name, "n_particles": a.n, "steps": a.steps,
                          "reference_eager_us_per_step": round(1e6 * t_ref / a.steps, 1),
                          "graft_us_per_step": round(1e6 * t_b / a.steps, 1),
                          "speedup": round(t_ref / t_b, 1)}), flush=True)


if __name__ == "__main__":
    main()
