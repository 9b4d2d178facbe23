"""Is the UNet forward + VJP of one guided step launch-bound?  Times the reference module eagerly and through
torch.cuda.make_graphed_callables (forward graph + backward graph, parameters frozen so that only the input
gradient is computed — the same work `autograd.grad(out, x)` does).   python tools/graph_probe.py [--n 8]"""
from __future__ import annotations

import argparse
import json
import os
import sys

import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import bench  # noqa: E402


def timed(fn, iters):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=8)
    ap.add_argument("--iters", type=int, default=10)
    a = ap.parse_args()
    dev = torch.device("cuda:0")
    model, name = bench.load_model("auto", dev)
    res = {"model": name}
    for n in sorted({max(1, a.n // 2), a.n, 2 * a.n}):
        x = torch.randn(n, 3, 256, 256, device=dev)
        t = torch.full((1,), 500.0, device=dev)
        g6 = torch.randn(n, 6, 256, 256, device=dev) * 1e-2

        def eager():
            xi = x.detach().requires_grad_(True)
            out = model(xi, t)
            return torch.autograd.grad(out, xi, g6)[0]

        res[f"eager_ms_n{n}"] = round(timed(eager, a.iters), 3)
    print(json.dumps(res), flush=True)
    n = a.n
    x = torch.randn(n, 3, 256, 256, device=dev)
    g6 = torch.randn(n, 6, 256, 256, device=dev) * 1e-2
    ref = None
    # the reference builds the sinusoidal frequency table on the CPU and copies it to the device on EVERY forward
    # (guided_diffusion/nn.py:103-121) — not capturable; cache the identical table on the device instead
    import math
    from guided_diffusion import nn as ref_nn, unet as ref_unet
    cache = {}

    def timestep_embedding(timesteps, dim, max_period=10000):
        key = (dim, max_period, timesteps.device)
        if key not in cache:
            half = dim // 2
            cache[key] = torch.exp(-math.log(max_period) * torch.arange(start=0, end=half, dtype=torch.float32) / half
                                   ).to(device=timesteps.device)
        args = timesteps[:, None].float() * cache[key][None]
        emb = torch.cat([torch.cos(args), torch.sin(args)], dim=-1)
        if dim % 2:
            emb = torch.cat([emb, torch.zeros_like(emb[:, :1])], dim=-1)
        return emb

    ref_nn.timestep_embedding = ref_unet.timestep_embedding = timestep_embedding

    def eager():
        xi = x.detach().requires_grad_(True)
        out = model(xi, t)
        return torch.autograd.grad(out, xi, g6)[0]

    ref = eager()
    # manual capture: forward graph, then backward graph (input gradient only) from the same memory pool
    xs = x.detach().clone().requires_grad_(True)
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for _ in range(3):
            out = model(xs, t)
            torch.autograd.grad(out, xs, g6)
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    gf, gb = torch.cuda.CUDAGraph(), torch.cuda.CUDAGraph()
    pool = torch.cuda.graph_pool_handle()
    with torch.cuda.graph(gf, pool=pool):
        out = model(xs, t)
    with torch.cuda.graph(gb, pool=pool):
        vjp = torch.autograd.grad(out, xs, g6)[0]

    def replay():
        gf.replay()
        gb.replay()
        return vjp

    res[f"graphed_ms_n{n}"] = round(timed(replay, a.iters), 3)
    with torch.no_grad():
        xs.copy_(x)
    got = replay().clone()
    res["max_abs_diff_vjp"] = float((got - ref).abs().max())
    res["vjp_scale"] = float(ref.abs().max())
    print(json.dumps(res))


if __name__ == "__main__":
    main()
