#!/usr/bin/env python
"""bench.py — DPS particle-steps/sec on B200 (BASELINE.json metric) + guidance-kernel roofline.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--workload auto|c1|c2|c3|c4|c5]

A "step" is one measurement-guided reverse-diffusion step over this rank's batch of particles:
UNet forward (the reference's own module, random-init, fp32) + residual / coefficient / cotangent kernels +
UNet VJP + fused posterior update (+ on resampling indices: log-weights, all-gather, CDF, ancestors, particle exchange).

Workloads (BASELINE.json configs):
  c1  Gaussian deblur, ps, ddpm, 8 particles/GPU                                                   (configs[0] on the GPU)
  c2  best-of-N N=8/GPU, 4x super-resolution, ps, ddpm 1000-step chain — no data-path collective  (configs[1]; N=1 headline)
  c3  motion deblur, ttc_ddim + multinomial resampling every 10th index, 8 particles/GPU, particles SHARDED:
      NCCL all-gather of (log-weight, distance) + fused P2P exchange kernel at resampling steps    (configs[2]; weak scaling)
  c4  phase retrieval, ps_anneal + annealing schedule, ttc_ddim resampling, N=32 GLOBAL split over the GPUs   (configs[3]; strong)
  c5  ImageNet-256 UNet, inpainting, ps_semantic + semantic term in the reweighting, 32 particles/GPU, micro-batched
      UNet forward+VJP                                                                              (configs[4]; weak scaling)
`--workload auto` (default): c2 on one GPU, c3 under torchrun — so the 1→8 scaling runs exercise the collectives.

--impl reference : the reference's own CPU implementation of the same step (from baseline/_ref, staged by
__graft_entry__.build(); else the oracle port) on the host cores — rank 0 only.
"""
from __future__ import annotations

import argparse
import functools
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

T_BYTES = 3 * 256 * 256 * 4
DIFF = dict(steps=1000, noise_schedule="linear", model_mean_type="epsilon", model_var_type="learned_range",
            dynamic_threshold=False, clip_denoised=True, rescale_timesteps=True)
FP32_PEAK_TFLOPS = 148 * 128 * 2 * 1.965e9 / 1e12      # nominal CUDA-core fp32 (FFMA) peak at clocks.max.sm

WORKLOADS = {
    "c1": dict(op="gaussian_blur", op_cfg=dict(kernel_size=61, intensity=3.0), method="ps", params=dict(scale=0.3),
               m_bytes=T_BYTES, sampler="ddpm", n=8, scaling="weak", model="ffhq",
               desc="DPS Gaussian deblur k=61 sigma=3, ps zeta=0.3, ddpm"),
    "c2": dict(op="super_resolution", op_cfg=dict(in_shape=(1, 3, 256, 256), scale_factor=4), method="ps",
               params=dict(scale=0.01), m_bytes=T_BYTES // 16, sampler="ddpm", n=8, scaling="weak", model="ffhq",
               desc="best-of-N N=8/GPU, 4x super-resolution (Resizer bicubic), ps zeta=0.01, ddpm 1000-step chain"),
    "c3": dict(op="motion_blur", op_cfg=dict(kernel_size=61, intensity=0.5), method="ps", params=dict(scale=0.3),
               m_bytes=T_BYTES, sampler="ttc_ddim", n=8, scaling="weak", model="ffhq",
               desc="batched TTC: motion deblur k=61 (synthetic sparse kernel, np seed 8), ps zeta=0.3, ttc_ddim with multinomial "
                    "resampling (w = exp(-d/100)) every 10th index, particles sharded over the GPUs"),
    "c4": dict(op="phase_retrieval", op_cfg=dict(oversample=2.0), method="ps_anneal", params=dict(scale=0.0001),
               m_bytes=T_BYTES * 9 // 4, sampler="ttc_ddim", n_global=32, scaling="strong", model="ffhq",
               loop=dict(anneal_amp=1.0, anneal_scale=10.0, anneal_loc=0.5),
               desc="phase retrieval (oversample 2 -> 384x384 |FFT|), ps_anneal with the annealing schedule, ttc_ddim "
                    "resampling every 10th index, N=32 particles GLOBAL split over the GPUs"),
    "c5": dict(op="inpainting", op_cfg={}, method="ps_semantic", params=dict(scale=0.5, sem_guid_scale=0.5), m_bytes=T_BYTES,
               sampler="ttc_ddim", n=32, scaling="weak", model="imagenet", loop=dict(semantic_weight=1.0), chunk="auto",
               desc="ImageNet-256 UNet, inpainting random mask p in (0.3,0.7) np seed 8, ps_semantic zeta=0.5 + semantic guidance "
                    "(seeded stand-in embedder: facenet is external), ttc_ddim with the semantic distance in the resampling "
                    "weights, 32 particles/GPU, UNet forward+VJP micro-batched"),
}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="auto", choices=["auto"] + sorted(WORKLOADS))
    ap.add_argument("--particles", type=int, default=None, help="particles per GPU (default: the workload's)")
    ap.add_argument("--model", default="auto", choices=["auto", "ffhq", "imagenet", "tiny"])
    ap.add_argument("--transport", default=None, choices=["p2p", "p2p_barrier", "allgather", "all_to_all"],
                    help="particle exchange transport of the sharded workloads (default: p2p, NCCL all-gather as fallback)")
    ap.add_argument("--chunk", default=None, help="UNet micro-batch (particles per forward+VJP pass): int or 'auto'")
    ap.add_argument("--cpu-particles", type=int, default=None,
                    help="particles per step of the CPU baseline (default: the workload's per-GPU count, capped at 8)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true",
                    help="skip the extra measurements (eager reference on the GPU, strict-fp32 UNet, sharded verification)")
    ap.add_argument("--channels-last", action="store_true", help="run the UNet module in NHWC memory format")
    ap.add_argument("--cudnn-benchmark", action="store_true", help="let cuDNN autotune its convolution algorithms")
    ap.add_argument("--sync-readback", action="store_true",
                    help="e2e: block the host on every step's distance read-back (the reference's .item() behaviour)")
    ap.add_argument("--eager-unet", action="store_true",
                    help="launch the UNet forward/VJP kernel by kernel instead of replaying them from CUDA graphs")
    return ap.parse_args()


def resolve_workload(args, world):
    name = args.workload
    if name == "auto":
        name = "c2" if world == 1 else "c3"
    wl = dict(WORKLOADS[name])
    if args.particles is not None:
        n = args.particles
    elif "n_global" in wl:
        if wl["n_global"] % world:
            raise SystemExit(f"{name}: {wl['n_global']} global particles do not split over {world} GPUs")
        n = wl["n_global"] // world
    else:
        n = wl["n"]
    wl["n"] = n
    if args.chunk is not None:
        wl["chunk"] = args.chunk if args.chunk == "auto" else int(args.chunk)
    return name, wl


# ------------------------------------------------------------------------------------------------
class StandInEmbedder(torch.nn.Module):
    """(N,3,H,W) → (N,64) seeded conv net standing in for the reference's external facenet embedder
    (facenet_pytorch InceptionResnetV1 + pretrained weights: not in the reference tree)."""

    def __init__(self, seed=0, dim=64):
        super().__init__()
        g = torch.Generator().manual_seed(seed)
        self.c1 = torch.nn.Conv2d(3, 16, 5, stride=4, padding=2)
        self.c2 = torch.nn.Conv2d(16, 32, 3, stride=2, padding=1)
        self.fc = torch.nn.Linear(32, dim)
        with torch.no_grad():
            for p in self.parameters():
                p.copy_(torch.randn(p.shape, generator=g) * 0.1)

    def forward(self, x):
        h = torch.tanh(self.c1(x))
        h = torch.tanh(self.c2(h)).mean(dim=(2, 3))
        return self.fc(h)


def load_model(kind, device):
    """The reference's UNet (random init + seeded re-init of the zeroed output convs) or, when the
    reference tree is not available, a small stand-in — reported in config.unet either way."""
    from dps_ttc_b200 import _ref
    if kind in ("ffhq", "imagenet") and _ref.reference_root() is not None:
        cfg = "imagenet_model_config.yaml" if kind == "imagenet" else "model_config.yaml"
        torch.manual_seed(0)    # the module's default init draws from the global generator: every rank must build the SAME network
        model = _ref.create_unet(cfg, reinit_zero_seed=0, device=device)
        name = "ImageNet-256 ADM UNet (552.8M)" if kind == "imagenet" else "FFHQ-256 ADM UNet (93.6M)"
        return model, name + ", reference module, random-init + seeded re-init of zeroed convs"
    if kind in ("ffhq", "imagenet") and os.environ.get("DPS_BENCH_ALLOW_STANDIN") != "1":
        raise RuntimeError("reference tree (baseline/_ref) not found: run __graft_entry__.build() in the build container")
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from helpers import TinyEps
    return TinyEps(seed=0, width=64).to(device).eval(), "STAND-IN TinyEps conv net (reference UNet unavailable)"


def make_inputs(n, seed):
    """Synthetic inputs on the HOST: x_true = 2U−1, x_start = randn (per-rank seed)."""
    g = torch.Generator().manual_seed(1234)
    x_true = torch.rand(1, 3, 256, 256, generator=g) * 2 - 1
    gs = torch.Generator().manual_seed(42 + seed)
    x_start = torch.randn(n, 3, 256, 256, generator=gs)
    return x_true, x_start


def build_b200(wl, device):
    from dps_ttc_b200.registry import get_conditioning_method, get_noise, get_operator
    from dps_ttc_b200.sampler import create_sampler
    np.random.seed(8)
    op = get_operator(wl["op"], device=device, **wl["op_cfg"])
    params = dict(wl["params"])
    if wl["method"] == "ps_semantic" and params.get("sem_guid_scale", 0):
        emb = StandInEmbedder(seed=3).to(device).eval()
        g = torch.Generator().manual_seed(99)
        with torch.no_grad():
            guid = emb((torch.rand(2, 3, 256, 256, generator=g) * 2 - 1).to(device))
        params.update(embedder=emb, guid_emb=guid.unsqueeze(0))
    cond = get_conditioning_method(wl["method"], op, get_noise("gaussian", sigma=0.05), **params)
    sampler = create_sampler(sampler=wl["sampler"], **DIFF)
    sampler.unet_chunk = wl.get("chunk")
    kw = {}
    if wl["op"] == "inpainting":
        from dps_ttc_b200.tables import MaskGenerator
        np.random.seed(8)
        mask = MaskGenerator("random", mask_prob_range=(0.3, 0.7), image_size=256)(np.zeros((1, 3, 256, 256)))[:, :1]
        kw["mask"] = torch.from_numpy(mask).to(device)
    return op, cond, sampler, kw


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md clocks line)."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self.proc = index, [], None

    def run(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.index), "-lms", "100"], stdout=subprocess.PIPE, text=True)
            for line in self.proc.stdout:
                self.rows.append([c.strip() for c in line.split(",")])
        except Exception:  # noqa: BLE001
            pass

    def stop(self):
        if self.proc is not None:
            self.proc.terminate()
        mhz = [float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows for i in range(4) if len(r) >= 6 and r[2 + i].lower() == "active"})
        return {"sm_mhz": float(np.median(mhz)) if mhz else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(mhz)}


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


# ------------------------------------------------------------------------------------------------
# The reference's own loop (unmodified classes from baseline/_ref), on the CPU or — as the like-for-like eager
# baseline — on the GPU.  Its p_sample_loop walks the WHOLE chain behind a tqdm progress bar; to time the same window of
# the same 1000-step chain as the B200 arm, the `tqdm` name inside the reference's module is bound to a pass-through
# that yields only the window's indices (and timestamps the first timed one).  No line of the reference is changed.
# ------------------------------------------------------------------------------------------------
class _Window:
    def __init__(self, start, warmup, steps, sync=None):
        self.start, self.warmup, self.steps, self.sync = start, warmup, steps, sync
        self.t0 = self.t1 = None

    def __call__(self, iterable, *a, **k):
        idxs = [i for i in iterable if self.start - self.warmup - self.steps < i <= self.start]
        win = self

        class It:
            def __iter__(self_inner):
                for j, i in enumerate(idxs):
                    if j == win.warmup:
                        if win.sync:
                            win.sync()
                        win.t0 = time.perf_counter()
                    yield i
                if win.sync:
                    win.sync()
                win.t1 = time.perf_counter()

            def set_postfix(self_inner, *a, **k):
                pass

            def set_description(self_inner, *a, **k):
                pass
        return It()


def reference_loop(wl, n_particles, steps, warmup, device):
    """`steps` timed guided steps of `n_particles` particles through the reference's p_sample_loop after `warmup` steps, window
    idx 999 … of the 1000-step chain.  Returns (particle_steps_per_sec, description) or None if the reference is not staged."""
    from dps_ttc_b200 import _ref
    if _ref.reference_root() is None:
        return None
    _ref.ensure_reference()
    with _ref.quiet():
        import guided_diffusion.gaussian_diffusion as ref_gd
        from guided_diffusion.condition_methods import get_conditioning_method
        from guided_diffusion.measurements import get_noise, get_operator
    on_gpu = torch.device(device).type == "cuda"
    if on_gpu:
        from dps_ttc_b200.graphed import make_reference_capturable  # noqa: F401  (not used: the eager loop stays eager)
    model = _ref.create_unet("imagenet_model_config.yaml" if wl["model"] == "imagenet" else "model_config.yaml",
                             reinit_zero_seed=0, device=device)
    x_true, x_start = make_inputs(n_particles, 0)
    x_true, x_start = x_true.to(device), x_start.to(device)
    np.random.seed(8)
    with _ref.quiet():
        op = get_operator(wl["op"], device=device, **wl["op_cfg"])
        noiser = get_noise("gaussian", sigma=0.05)
        # ps_semantic(sem_guid_scale=0) is the HEAD-valid spelling of ps inside the base loop
        cond = get_conditioning_method("ps_semantic", op, noiser, scale=wl["params"].get("scale", 0.3), sem_guid_scale=0.0)
    kw = {}
    if wl["op"] == "inpainting":
        from util.img_utils import mask_generator
        np.random.seed(8)
        kw["mask"] = mask_generator("random", mask_prob_range=(0.3, 0.7), image_size=256)(x_true)[:, 0].unsqueeze(0)
    y = noiser(op.forward(x_true, **kw)).detach()
    fn = functools.partial(cond.conditioning, **kw) if kw else cond.conditioning
    with _ref.quiet():
        s = ref_gd.create_sampler(sampler="ddpm", **DIFF)
    win = _Window(999, warmup, steps, sync=(lambda: torch.cuda.synchronize(device)) if on_gpu else None)
    saved = ref_gd.tqdm
    ref_gd.tqdm = win
    try:
        with _ref.quiet():
            s.p_sample_loop(model=model, x_start=x_start.clone(), measurement=y, measurement_cond_fn=fn, record=False,
                            save_root=None)
    finally:
        ref_gd.tqdm = saved
    dt = win.t1 - win.t0
    what = (f"reference p_sample_loop (ddpm + ps_semantic sem=0 = ps; {wl['op']}) on {'cuda eager' if on_gpu else 'CPU'}, "
            f"{n_particles} particle(s)/step x {steps} timed steps after {warmup} warm-up, idx {999 - warmup}..{999 - warmup - steps + 1} "
            f"of the 1000-step chain, {'ImageNet' if wl['model'] == 'imagenet' else 'FFHQ'} UNet fp32")
    del model
    return n_particles * steps / dt, what


def cpu_particle_steps(wl, n_particles, steps, warmup):
    """(particle_steps_per_sec, kind, cores, description) of the CPU arm."""
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    res = reference_loop(wl, n_particles, steps, warmup, "cpu")
    if res is not None:
        return res[0], "reference", cores, res[1]
    # oracle port (numpy) with the stand-in model
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from helpers import TinyEps, oracle_guided_step
    from oracle import dps_oracle as O
    model = TinyEps(seed=0, width=64)
    tab = O.Tables(1000)
    fwd = lambda a: O.resize_forward(a, 0.25)  # noqa: E731
    adj = lambda u: O.resize_adjoint(u, 0.25, 256, 256)  # noqa: E731
    rng = np.random.default_rng(0)
    x_true, x_start = make_inputs(n_particles, 0)
    y = fwd(x_true.numpy())
    img = x_start.numpy()
    t0 = None
    for i, idx in enumerate(range(999, 999 - warmup - steps, -1)):
        if i == warmup:
            t0 = time.perf_counter()
        img, _, _ = oracle_guided_step(O, model, tab, img, idx, y, fwd, adj, rng.standard_normal(img.shape).astype(np.float32),
                                       "norm", 0.01)
    dt = time.perf_counter() - t0
    return n_particles * steps / dt, "port", 1, f"oracle port (numpy) + stand-in model, {n_particles} particle(s) x {steps} steps, SR x4"


def run_reference(args, rank, world):
    if rank != 0:
        return
    name, wl = resolve_workload(args, world)
    n_cpu = args.cpu_particles or min(8, wl["n"])
    val, kind, cores, what = cpu_particle_steps(wl, n_cpu, args.steps, args.warmup)
    line = {"impl": "reference", "metric": "DPS particle-steps/sec", "value": val, "unit": "particle-steps/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1000.0 * n_cpu / val, "higher_is_better": True, "scaling": wl["scaling"],
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"{name}: {wl['desc']}", "device": "cpu", "particles_per_step": n_cpu,
                       "note": "the reference has no particle sharding or resampling collective: its CPU arm is the guided step "
                               "(UNet forward + conditioning + VJP + update) of the same operator on the same chain window"},
            "cpu_baseline": {"value": val, "unit": "particle-steps/s", "cores": cores, "kind": kind, "sample": what},
            "e2e": {"value": val, "unit": "particle-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    emit(line)


# ------------------------------------------------------------------------------------------------
def graph_durations(plan, sampler_kind, k, n, y, device, want_gather, philox, reps=4):
    """µs per launch of the operator forward / adjoint, the posterior update (and the resampling gather) at N = n particles:
    `reps` passes over S rotating argument sets captured in ONE CUDA graph per kernel and replayed, CUDA events around
    the replays on the launching stream."""
    from dps_ttc_b200 import kernels
    S = max(3, -(-256 * 2**20 // (2 * n * T_BYTES)))           # one tensor stream alone cycles through ≥ 2x L2
    g = torch.Generator(device).manual_seed(5)
    rnd = lambda *s: torch.randn(*s, device=device, generator=g)  # noqa: E731
    X = [rnd(n, 3, 256, 256) / k.c1 for _ in range(S)]
    O6 = [rnd(n, 6, 256, 256) * 0.3 / k.c2 for _ in range(S)]
    Z = [rnd(n, 3, 256, 256) for _ in range(S)]
    G6 = [rnd(n, 6, 256, 256) * 1e-2 for _ in range(S)]
    VJ = [rnd(n, 3, 256, 256) * 1e-2 for _ in range(S)]
    OUT = [torch.empty(n, 3, 256, 256, device=device) for _ in range(S)]
    R = [torch.empty((n,) + tuple(plan.out_shape), device=device) for _ in range(S)]
    AUX = [plan.new_aux(n) for _ in range(S)]
    coef = torch.full((n,), -0.01, device=device)
    ids = torch.randint(0, n, (n,), generator=torch.Generator().manual_seed(1)).to(device)
    upd = f"posterior_update_{sampler_kind}"
    fns = {f"{plan.kind}_forward": lambda i: plan.forward(X[i], O6[i][:, :3], k, True, y, want_partials=True, aux=AUX[i], out=R[i]),
           f"{plan.kind}_adjoint": lambda i: plan.adjoint(R[i], coef, X[i], O6[i][:, :3], k, True, None, out=G6[i][:, :3], aux=AUX[i])}
    PART = torch.rand(n, max(plan.guidance_partials, plan.partials_per_particle), 2, device=device) * 10
    dfr = (PART, 1, 0.01, torch.empty(n, device=device))
    if sampler_kind == "ddpm":
        fns[upd] = lambda i: kernels.posterior_update("ddpm", X[i], O6[i][:, :3], O6[i][:, 3:], None if philox else Z[i], k,
                                                      g=G6[i][:, :3], vjp=VJ[i], out=OUT[i], deferred=dfr,
                                                      philox=(1, 5, 0) if philox else None)
    else:
        fns[upd] = lambda i: kernels.posterior_update("ddim", X[i], O6[i][:, :3], None, None, k, g=G6[i][:, :3], vjp=VJ[i],
                                                      out=OUT[i], deferred=dfr)
    if plan.guidance_partials > 0:
        fns[f"{plan.kind}_guidance"] = lambda i: plan.guidance(X[i], O6[i][:, :3], k, True, y, out=G6[i][:, :3])
    if want_gather:
        fns["gather_particles"] = lambda i: kernels.gather_particles(X[i], ids, out=OUT[i])
    out = {}
    for name, fn in fns.items():
        for i in range(3):
            fn(i)                                              # warm-up (and first-launch attribute setup)
        torch.cuda.synchronize()
        graph, side = torch.cuda.CUDAGraph(), torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side), torch.cuda.graph(graph, stream=side, capture_error_mode="thread_local"):
            for _ in range(reps):
                for i in range(S):
                    fn(i)
        torch.cuda.current_stream().wait_stream(side)
        graph.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            graph.replay()
        e1.record()
        torch.cuda.synchronize()
        out[name] = e0.elapsed_time(e1) * 1e3 / (3 * reps * S)
    return out


def run_b200(args, rank, world, local_rank):
    import torch.distributed as dist
    from dps_ttc_b200 import _lib, kernels
    from dps_ttc_b200.dist import ParticleShards, shared_uniforms
    from dps_ttc_b200.sampler import NoiseTape, PhiloxNoise, TorchNoise
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py --impl b200 needs a GPU: dps_ttc_b200 has no CPU fallback")
    device = torch.device(f"cuda:{local_rank}")
    torch.cuda.set_device(device)
    name, wl = resolve_workload(args, world)
    n, K, W = wl["n"], args.steps, args.warmup
    m_bytes, desc = wl["m_bytes"], wl["desc"]
    searching = wl["sampler"] in ("ttc_ddim", "search_ddpm")
    model_kind = wl["model"] if args.model == "auto" else args.model
    model, model_name = load_model(model_kind, device)
    if args.channels_last:
        model = model.to(memory_format=torch.channels_last)
        model_name += " [channels_last]"
    if args.cudnn_benchmark:
        torch.backends.cudnn.benchmark = True
        model_name += " [cudnn.benchmark]"
    op, cond, sampler, kw = build_b200(wl, device)
    cond_fn = functools.partial(cond.conditioning, **kw) if kw else cond.conditioning
    x_true, x_start_h = make_inputs(n, rank)
    x_start_h = x_start_h.pin_memory()
    with torch.no_grad():
        y_dev = op.forward(x_true.to(device), **kw)
        y_dev = y_dev + 0.05 * torch.randn(y_dev.shape, device=device, generator=torch.Generator(device).manual_seed(1235))
    y_h = y_dev.cpu().pin_memory()
    torch.manual_seed(1000 + rank)
    graph_model = not args.eager_unet
    loop_kw = dict(wl.get("loop", {}))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def new_shards(timing=False):
        return ParticleShards(n, transport=args.transport, timing=timing) if searching else None

    def loop(smp, x0, y, start, steps, shards=None, **extra):
        kws = dict(model=model, x_start=x0, measurement=y, measurement_cond_fn=cond_fn, record=False, save_root=None,
                   start_idx=start, num_steps=steps, **loop_kw)
        kws["graph_model"] = graph_model
        kws.update(extra)
        if shards is not None:
            kws["shards"] = shards
        res = smp.p_sample_loop(**kws)
        return res[0], res[1]

    # ---------------- device-resident throughput (`value`) ----------------
    # device-resident leg: the update kernel generates its own noise (Philox, keyed by seed/step/particle) — the throughput
    # mode; the e2e leg below feeds recorded host noise through a NoiseTape instead
    sampler.noise, sampler.parity_rng = PhiloxNoise(seed=1000 + rank), False
    shards = new_shards(timing=True)
    x_dev = x_start_h.to(device)
    if shards is not None:
        shards.publish_target(x_dev)                            # symmetric-memory allocation + rendezvous happen here, untimed
    if shards is not None:
        # one resampling step outside the W warm-up steps (which contain no resampling index): loads the resampling /
        # exchange kernels (CUDA loads a kernel lazily at its first launch) and runs the first NCCL all-gather of this size
        loop(sampler, x_dev, y_dev, 990, 1, shards)
    img, _ = loop(sampler, x_dev, y_dev, 999, W, shards)       # warm-up steps (untimed)
    if shards is not None:
        shards.spans.clear()
        shards.bytes_exchanged, shards.exchanges = 0, 0
    kernels.TIMER = kernels.KernelTimer()
    clocks = ClockSampler(local_rank)
    clocks.start()
    barrier()
    _lib.reset_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    img, dist_dev = loop(sampler, img, y_dev, 999 - W, K, shards)   # exactly K timed steps
    e1.record()
    barrier()
    launches = _lib.launch_count()
    eff_chunk = sampler._chunk_size(model, x_dev)               # particles per UNet forward+VJP pass actually used
    ms = e0.elapsed_time(e1)
    clock_info = clocks.stop()
    spans = kernels.TIMER.summary()
    kernels.TIMER = None
    exch = None
    if shards is not None:
        idxs_timed = range(999 - W, 999 - W - K, -1)
        resample_idx = [i for i in idxs_timed if i % sampler.resample_every_steps == 0]
        exch = {"transport": shards.transport, "resampling_indices_in_window": resample_idx,
                "exchanges": shards.exchanges, "exchange_us": None if shards.exchange_us() is None else round(shards.exchange_us(), 1),
                "bytes_exchanged_per_rank": int(shards.bytes_exchanged),
                "collectives": ("NCCL all-gather of (log-weight, distance) pairs, N x 8 B" if world > 1 else "none (one rank)") +
                               {"p2p": " + dps_exchange_particles_p2p (in-kernel rendezvous, NVLink loads from the owners' "
                                       "symmetric buffers; x_{t-1} written into the buffer by the update kernel)",
                                "p2p_barrier": " + staging copy + symmetric-memory barriers + dps_gather_particles_p2p",
                                "allgather": " + NCCL all-gather of all particles + local gather kernel",
                                "all_to_all": " + NCCL all_to_all_single of the needed particles (host-side split sizes)",
                                "local": " + local gather kernel"}[shards.transport]}

    # ---------------- end to end through the public API with HOST buffers (`e2e`) ----------------
    gz = torch.Generator().manual_seed(77 + rank)
    idxs = list(range(999 - W - K, 999 - W - 2 * K, -1))
    needs_z = sampler.kind == "ddpm"
    tape = NoiseTape(z={i: torch.randn(n, 3, 256, 256, generator=gz) for i in idxs} if needs_z else None)
    sampler.noise = tape
    d2h = {"bytes": 0, "step": 0}
    # The reference reads the distance every step for its progress bar (:295) with a blocking .item().  Here every
    # step's distance vector is copied into its own slot of a pinned host ring on the sampling stream (one D2H per
    # step, inside the timed region); the host does not block on it, so the next step's launches are not held back
    # (--sync-readback restores the blocking read).  All slots are checked after the final synchronisation.
    host_ring = torch.full((K, n), float("nan"), dtype=torch.float32).pin_memory()

    def read_back(idx, im, d, sd):
        host_ring[d2h["step"]].copy_(d.reshape(-1), non_blocking=not args.sync_readback)
        d2h["step"] += 1
        d2h["bytes"] += d.numel() * 4

    e2e_shards = new_shards()
    if e2e_shards is not None:
        e2e_shards.publish_target(x_dev)
    barrier()
    t0 = time.perf_counter()
    x_in = x_start_h.to(device, non_blocking=True)             # H2D of the particles and the measurement
    y_in = y_h.to(device, non_blocking=True)
    out, _ = loop(sampler, x_in, y_in, idxs[0], K, e2e_shards, callback=read_back)
    out_h = out.cpu()                                          # D2H of the result
    barrier()
    e2e_s = time.perf_counter() - t0
    if d2h["step"] != K or not bool(torch.isfinite(host_ring).all()):
        raise RuntimeError("e2e: a per-step distance read-back did not arrive on the host")
    h2d_step = (tape.h2d_bytes + x_start_h.numel() * 4 + y_h.numel() * 4) / K
    d2h_step = (d2h["bytes"] + out_h.numel() * 4) / K

    # ---------------- max over ranks ----------------
    t = torch.tensor([ms / 1000.0, e2e_s], device=device, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    sec, e2e_sec = float(t[0]), float(t[1])

    # ---------------- sharded == unsharded? (all ranks take part; rank 0 re-runs the loop unsharded) ----------------
    verify = None
    if searching and world > 1 and not args.no_extras and n * world <= 64 and model_kind != "imagenet":
        # (rank 0 re-runs all N particles alone with an eager, deterministic UNet: bounded to N ≤ 64 particles of the FFHQ
        # model so that the default run stays within minutes; c5 shares the sharded code path with c3 / c4)
        verify = verify_sharded(args, wl, model, cond_fn, y_dev, device, rank, world, n, loop)
    same_1gpu = None
    if searching and world > 1 and not args.no_extras:
        # the SAME workload on ONE GPU (rank 0 alone, the others wait): the like-for-like denominator of the scaling curve
        barrier()
        if rank == 0 and wl["scaling"] == "weak":
          try:
            from dps_ttc_b200.sampler import create_sampler
            s1 = create_sampler(sampler=wl["sampler"], **DIFF)
            s1.unet_chunk, s1.noise, s1.parity_rng = wl.get("chunk"), PhiloxNoise(seed=1000), False
            im1, _ = loop(s1, x_dev, y_dev, 999, W)
            torch.cuda.synchronize()
            a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a0.record()
            loop(s1, im1, y_dev, 999 - W, K)
            a1.record()
            torch.cuda.synchronize()
            same_1gpu = {"value": n * K / (a0.elapsed_time(a1) / 1e3), "unit": "particle-steps/s",
                         "ms_per_step": a0.elapsed_time(a1) / K,
                         "what": f"workload {name} unsharded on rank 0 alone ({n} particles, local gather kernel), same window"}
          except Exception as e:  # noqa: BLE001
            same_1gpu = {"unavailable": f"{type(e).__name__}: {e}"[:300]}
        barrier()
    if rank != 0:
        return
    n_global = world * n
    value = n_global * K / sec
    e2e_value = n_global * K / e2e_sec

    # ---------------- roofline of the graft kernels ----------------
    # Two live measurements, both with CUDA events on the launching stream:
    #  * `bracket_us`: one event pair around every graft launch INSIDE the timed region (gives the graft's share of
    #    the step).  An event pair costs ≈5 µs on this system (x0_from_eps: 7.5 µs bracketed, 2.5 µs back to back),
    #    which is as long as the kernels themselves at N = 8, so a bracket cannot serve as the kernel duration.
    #  * `mean_us`: the same launch (same plan, same shapes, this step's constants) replayed back to back from one
    #    CUDA graph over rotating argument sets that together exceed 2x the L2 — no event and no CPU between launches,
    #    inputs cold in L2 like in the real step, where a UNet forward+VJP runs between two graft launches.
    #    ncu's gpu__time_duration for the same kernels (profiles/) agrees with this number, not with the bracket.
    plan = op.plan_for(x_dev, **kw)
    kind = plan.kind
    upd = f"posterior_update_{sampler.kind}"
    alg_bytes = {f"{kind}_forward": n * (2 * T_BYTES + m_bytes), f"{kind}_adjoint": n * (3 * T_BYTES + m_bytes),
                 upd: n * (7 if sampler.kind == "ddpm" else 5) * T_BYTES, "gather_particles": n * 2 * T_BYTES, "guidance_coef": 0}
    peak, peak_src = peaks()
    philox = sampler.kind == "ddpm"
    alg_bytes[f"{kind}_guidance"] = n * (3 * T_BYTES + m_bytes)
    alg_bytes[f"{kind}_forward+adjoint"] = alg_bytes[f"{kind}_forward"] + alg_bytes[f"{kind}_adjoint"]
    if philox:
        alg_bytes[upd] = n * 6 * T_BYTES                        # no z tensor: the noise is generated in the kernel
    live = graph_durations(plan, sampler.kind, sampler.schedule.consts(500), n, y_dev, device, want_gather=searching, philox=philox)
    if f"{kind}_forward" in live and f"{kind}_adjoint" in live:
        live[f"{kind}_forward+adjoint"] = live[f"{kind}_forward"] + live[f"{kind}_adjoint"]
    ktab = {}
    for kname, (cnt, mean_ms) in spans.items():
        b = alg_bytes.get(kname, 0)
        us = live.get(kname)
        ktab[kname] = {"launches": cnt, "bracket_us": round(mean_ms * 1e3, 2), "mean_us": None if us is None else round(us, 2),
                       "alg_bytes": b, "gbs": round(b / (us * 1e-6) / 1e9, 1) if us else None,
                       "frac": round(b / (us * 1e-6) / 1e9 / peak, 4) if us else None}
    if searching and "gather_particles" in live:
        us = live["gather_particles"]
        ktab["gather_particles"] = {"launches": len(exch["resampling_indices_in_window"]) if exch else 0, "bracket_us": None,
                                    "mean_us": round(us, 2), "alg_bytes": alg_bytes["gather_particles"],
                                    "gbs": round(alg_bytes["gather_particles"] / (us * 1e-6) / 1e9, 1),
                                    "frac": round(alg_bytes["gather_particles"] / (us * 1e-6) / 1e9 / peak, 4),
                                    "note": "local form of the exchange kernel (same copy loop); the cross-GPU exchange is timed as exchange_us"}
    # dominant kernel = largest launches x mean_us over the timed region, no tie-break
    cands = [kn for kn in ktab if ktab[kn]["alg_bytes"] > 0 and ktab[kn]["mean_us"] and ktab[kn]["launches"]]
    t_of = lambda kn: ktab[kn]["mean_us"] * ktab[kn]["launches"]  # noqa: E731
    dom = max(cands, key=t_of)
    achieved = ktab[dom]["gbs"]
    agg_bytes = sum(ktab[kn]["alg_bytes"] * ktab[kn]["launches"] for kn in cands)
    agg_us = sum(t_of(kn) for kn in cands)
    roofline = {"kernel": dom, "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": round(achieved / peak, 4), "traffic": None, "peak_source": peak_src,
                "aggregate_frac": round(agg_bytes / (agg_us * 1e-6) / 1e9 / peak, 4),
                "aggregate": f"sum of algorithmic bytes / sum of mean_us over all graft launches of the timed region "
                             f"({agg_bytes / 1e6:.0f} MB in {agg_us:.0f} us)",
                "traffic_note": "not measured by this run; ncu dram__bytes per launch of the same kernels at the same size are in "
                                "profiles/ (r2*_ncu_*.csv) — reads equal the algorithmic input bytes, writes partly still in L2",
                "timing": "mean_us = CUDA-graph replay of the launch over L2-exceeding rotating arguments, CUDA events on the "
                          "launching stream; bracket_us = per-launch event pair inside the timed region (includes ~5 us of "
                          "event overhead)",
                "note": f"N={n} particles/launch: {alg_bytes[dom] / 1e6:.0f} MB per launch, {ktab[dom]['mean_us']} us — at this size a "
                        f"launch is paced by its ~3 us ramp + one DRAM latency chain; HBM-regime numbers (N>=128) are in profiles/",
                "dominant_rule": "largest launches x mean_us (no tie-break; every kernel is listed under `kernels`)",
                "kernels": ktab}
    if kind == "blur_sparse":
        # motion blur is not HBM-bound: report where it stands against the fp32 pipe too (taps x 2 flop per output pixel)
        flop = 2.0 * plan.taps * n * 3 * 256 * 256
        roofline["fp32"] = {k_: {"tflops": round(flop / (ktab[k_]["mean_us"] * 1e-6) / 1e12, 2),
                                 "frac_of_fp32_peak": round(flop / (ktab[k_]["mean_us"] * 1e-6) / 1e12 / FP32_PEAK_TFLOPS, 3)}
                            for k_ in (f"{kind}_forward", f"{kind}_adjoint") if k_ in ktab and ktab[k_]["mean_us"]}
        roofline["fp32"]["peak_tflops"] = round(FP32_PEAK_TFLOPS, 1)
        roofline["fp32"]["note"] = (f"{plan.taps} non-zero taps: {2 * plan.taps} flop per output pixel against 12-16 B — "
                                    "compute/shared-memory bound, not HBM; nominal peak = 148 SM x 128 lanes x 2 x 1.965 GHz")
    graft_ms = sum((v["bracket_us"] or 0) * v["launches"] for v in ktab.values()) / 1e3
    roofline["graft_share_of_step"] = round(graft_ms / ms, 5)

    extras = {}
    if world == 1 and not args.no_extras:
        extras = single_gpu_extras(args, wl, name, model, sampler, loop, x_dev, y_dev, device, n, K, W)
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        n_cpu = args.cpu_particles or min(8, n)
        del model
        torch.cuda.empty_cache()
        v, ckind, cores, what = cpu_particle_steps(wl, n_cpu, 2, 1)
        cpu = {"value": v, "unit": "particle-steps/s", "cores": cores, "kind": ckind, "sample": what}
    par = f"particle-sharded dp{world}, " + (exch["collectives"] if exch else "no data-path collective (best-of-N selects after the loop)")
    line = {"metric": "DPS particle-steps/sec", "value": value, "unit": "particle-steps/s", "n_gpus": world, "steps": K,
            "warmup": W, "ms_per_step": 1000.0 * sec / K, "higher_is_better": True, "scaling": wl["scaling"],
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"{name}: {desc}", "particles_per_gpu": n, "global_particles": n_global,
                       "image": "3x256x256", "sampler": wl["sampler"],
                       "chain": "1000 steps, timed window idx %d..%d" % (999 - W, 999 - W - K + 1),
                       "noise": "value: Philox4x32-10 inside the update kernel (no z tensor); e2e: z copied from pinned host memory every step",
                       "guidance_coefficient": "deferred to the update kernel (dps_update_ext): residual+cotangent = " +
                                               ("ONE fused cluster kernel" if plan.guidance_partials > 0 else "forward + adjoint launches") +
                                               ", no coefficient launch",
                       "unet": model_name, "unet_chunk": eff_chunk,
                       "unet_launch": "eager (kernel by kernel)" if args.eager_unet else
                                      "the module's own forward and input-VJP kernels replayed from two CUDA graphs "
                                      "(dps_ttc_b200/graphed.py; same kernels, same order)",
                       "unet_math": "fp32 weights/activations; cuDNN conv TF32 = torch default "
                       f"({torch.backends.cudnn.allow_tf32}), matmul TF32 {torch.backends.cuda.matmul.allow_tf32}; the strict-fp32 "
                       "number is under `strict_fp32`",
                       "parallelism": par,
                       "l2_policy": "inputs larger than L2: between two launches of any graft kernel the UNet forward+VJP "
                                    "streams several GB of activations (1.9 GB saved per particle), so no explicit flush is needed"},
            "clocks": clock_info, "gpu_launches": int(launches),
            "e2e": {"value": e2e_value, "unit": "particle-steps/s", "h2d_bytes_per_step": int(h2d_step),
                    "d2h_bytes_per_step": int(d2h_step), "ms_per_step": 1000.0 * e2e_sec / K,
                    "readback": "blocking per step" if args.sync_readback else
                                "per-step D2H into a pinned host ring on the sampling stream, host waits once at the end"},
            "roofline": roofline, "cpu_baseline": cpu}
    if exch is not None:
        line["exchange"] = exch
    if verify is not None:
        line["sharded_bit_identical"] = verify["bit_identical"]
        line["sharded_check"] = verify
    if same_1gpu is not None:
        line["single_gpu_same_workload"] = same_1gpu
    line.update(extras)
    emit(line)


def verify_sharded(args, wl, model, cond_fn, y_dev, device, rank, world, n, loop):
    """Sharded run vs the unsharded run of the SAME N particles on rank 0 (same noise tape, same resampling uniforms, UNet
    micro-batched in the ranks' slices so every cuDNN call has the shape it has in the sharded run): ancestors at every
    resampling index and the final particles + distances must agree bit for bit.  The UNet runs EAGER with
    cudnn.deterministic=True in both runs: cuDNN's default backward-data kernels accumulate with atomics (forward+VJP of
    the FFHQ UNet is not repeatable bit for bit without the flag — tools/determinism_probe.py), and a CUDA graph freezes
    whatever algorithm the heuristics picked at capture time."""
    import torch.distributed as dist
    from dps_ttc_b200.dist import ParticleShards, shared_uniforms
    from dps_ttc_b200.sampler import NoiseTape, create_sampler
    start, steps = 992, 13                                      # idx 992..980: resampling at 990 and 980
    det = torch.backends.cudnn.deterministic
    torch.backends.cudnn.deterministic = True
    N = n * world
    try:
        def tape_for(ranks):
            z = None
            if wl["sampler"] != "ttc_ddim":                    # DDIM (eta = 0) draws no z
                z = {}
                for i in range(start, start - steps, -1):
                    z[i] = torch.cat([torch.randn(n, 3, 256, 256, generator=torch.Generator().manual_seed(7000 * r + i))
                                      for r in ranks])
            uni = {i: shared_uniforms(0, i, N, "cpu") for i in range(start, start - steps, -1)}
            return NoiseTape(z=z, uniforms=uni)

        def fresh():
            s = create_sampler(sampler=wl["sampler"], **DIFF)
            s.unet_chunk, s.parity_rng = wl.get("chunk"), False
            return s

        s_sh = fresh()
        s_sh.noise = tape_for([rank])
        x0 = make_inputs(n, rank)[1].to(device)
        sh = ParticleShards(n, transport=args.transport)
        img, d = loop(s_sh, x0, y_dev, start, steps, sh, graph_model=False)
        anc = s_sh.last_stats["ancestors"]
        anc_keys = sorted(anc, reverse=True)
        anc_mine = torch.stack([anc[i] for i in anc_keys]) if anc_keys else torch.zeros((0, N), dtype=torch.int64, device=device)
        all_img = torch.empty((N,) + tuple(img.shape[1:]), device=device)
        all_d = torch.empty((N,), device=device)
        all_anc = torch.empty((world,) + tuple(anc_mine.shape), dtype=torch.int64, device=device)
        dist.all_gather_into_tensor(all_img, img.contiguous())
        dist.all_gather_into_tensor(all_d, d.contiguous())
        dist.all_gather_into_tensor(all_anc, anc_mine.contiguous())
        res = None
        if rank == 0:
          try:
            s_un = fresh()
            s_un.noise = tape_for(range(world))
            if s_un.unet_chunk is None or s_un.unet_chunk == "auto" or int(s_un.unet_chunk) > n:
                s_un.unet_chunk = n                              # UNet launches of the ranks' shape
            x_all = torch.cat([make_inputs(n, r)[1] for r in range(world)]).to(device)
            u_img, u_d = loop(s_un, x_all, y_dev, start, steps, graph_model=False)
            u_anc = s_un.last_stats["ancestors"]
            same_anc = sorted(u_anc, reverse=True) == anc_keys and all(
                bool((all_anc[:, j] == u_anc[i].unsqueeze(0)).all()) for j, i in enumerate(anc_keys))
            dmax = float((all_img - u_img).abs().max())
            ddist = float((all_d - u_d).abs().max())
            res = {"bit_identical": bool(same_anc and dmax == 0.0 and ddist == 0.0), "ancestors_identical_on_all_ranks": bool(same_anc),
                   "resampling_indices": anc_keys, "max_abs_particle_diff": dmax, "max_abs_distance_diff": ddist,
                   "window": f"idx {start}..{start - steps + 1}", "transport": sh.transport,
                   "what": f"{N} particles: {world} ranks x {n} sharded vs rank 0 unsharded (UNet micro-batched by {n}), "
                           "eager UNet with cudnn.deterministic=True for both"}
          except Exception as e:  # noqa: BLE001  (never leave the other ranks waiting at the barrier)
            res = {"bit_identical": None, "error": f"{type(e).__name__}: {e}"[:300]}
        dist.barrier()
        return res
    finally:
        torch.backends.cudnn.deterministic = det


def single_gpu_extras(args, wl, name, model, sampler, loop, x_dev, y_dev, device, n, K, W):
    """Driver-recorded context for the N=1 line: (1) the reference's OWN loop in torch eager on cuda:0 with the same UNet, same
    window, same particle count — what the graft replaces, like for like; (2) the B200 arm again with TF32 switched off in
    cuDNN and cuBLAS (every parity test runs that way)."""
    from dps_ttc_b200.sampler import PhiloxNoise, create_sampler
    out = {}
    sampler.__dict__.pop("_graphs", None)                      # release the main arm's CUDA-graph pools
    torch.cuda.empty_cache()
    n_ref = min(n, 8)                                          # the eager loop keeps every particle's activations at once
    try:
        res = reference_loop(wl, n_ref, K, W, device)
        if res is not None:
            out["gpu_eager_reference"] = {"value": res[0], "unit": "particle-steps/s", "ms_per_step": 1000.0 * n_ref / res[0],
                                          "what": res[1] + "; TF32 as the torch default, like the B200 arm"}
    except Exception as e:  # noqa: BLE001
        out["gpu_eager_reference"] = {"unavailable": f"{type(e).__name__}: {e}"[:300]}
    torch.cuda.empty_cache()
    tf = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.allow_tf32 = torch.backends.cuda.matmul.allow_tf32 = False
    try:
        s2 = create_sampler(sampler=wl["sampler"], **DIFF)     # fresh sampler: its CUDA graphs are captured without TF32
        s2.unet_chunk, s2.noise, s2.parity_rng = wl.get("chunk"), PhiloxNoise(seed=1000), False
        im, _ = loop(s2, x_dev, y_dev, 999, W)
        torch.cuda.synchronize()
        a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a0.record()
        loop(s2, im, y_dev, 999 - W, K)
        a1.record()
        torch.cuda.synchronize()
        out["strict_fp32"] = {"value": n * K / (a0.elapsed_time(a1) / 1e3), "unit": "particle-steps/s",
                              "ms_per_step": a0.elapsed_time(a1) / K,
                              "what": "same timed window with torch.backends.cudnn.allow_tf32 = cuda.matmul.allow_tf32 = False"}
        del s2
    except Exception as e:  # noqa: BLE001
        out["strict_fp32"] = {"unavailable": f"{type(e).__name__}: {e}"[:300]}
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = tf
    torch.cuda.empty_cache()
    return out


_JSON_FD = None


def emit(line):
    """The ONE JSON line, on the process's original stdout."""
    data = (json.dumps(line) + "\n").encode()
    if _JSON_FD is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_JSON_FD, data)


def main():
    global _JSON_FD
    args = parse()
    # libraries print to stdout behind Python's back (NCCL's version banner, tqdm of the reference loop): keep the
    # original stdout for the JSON line only and send everything else to stderr
    sys.stdout.flush()
    _JSON_FD = os.dup(1)
    os.dup2(2, 1)
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local_rank = int(os.environ.get("LOCAL_RANK", 0))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.cuda.set_device(local_rank)
        import datetime
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local_rank}"), timeout=datetime.timedelta(minutes=6))
    try:
        run_b200(args, rank, world, local_rank)
    finally:
        if world > 1:
            import torch.distributed as dist
            dist.barrier()
            dist.destroy_process_group()


if __name__ == "__main__":
    main()
