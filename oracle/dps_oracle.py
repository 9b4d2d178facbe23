"""CPU restatement (numpy, fp32) of the dps-ttc hot path — TEST INFRASTRUCTURE, NOT PRODUCT.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may import this
module; the product (dps_ttc_b200) never does and has no CPU path at all.

Every function cites the reference code it follows (paths relative to /root/reference).  The oracle
is PINNED: tests/test_oracle_pins.py checks it against fixtures produced by running the reference
itself in the build container (oracle/make_golden.py → tests/golden/*.npz).  The reference ships no
tests, golden vectors or KATs of its own, so these fixtures are the pin.

Parity UNPINNED items (third-party code absent from the reference tree): motion-kernel
generation (`motionblur`), the bkse nonlinear blur, facenet embeddings.  They enter as inputs.
"""
from __future__ import annotations

import math

import numpy as np

f32 = np.float32


# ------------------------------------------------------------------------------------------------
# schedule  (guided_diffusion/gaussian_diffusion.py:59-117, :338-418, :718-763)
# ------------------------------------------------------------------------------------------------
def linear_betas(steps):
    scale = 1000 / steps
    return np.linspace(scale * 0.0001, scale * 0.02, steps, dtype=np.float64)


def spaced_steps(num, counts):
    """space_timesteps for a list/comma-string of section counts (gaussian_diffusion.py:366-392)."""
    if isinstance(counts, str):
        counts = [int(c) for c in counts.split(",")]
    size_per, extra = num // len(counts), num % len(counts)
    start, steps = 0, []
    for i, c in enumerate(counts):
        size = size_per + (1 if i < extra else 0)
        stride = 1 if c <= 1 else (size - 1) / (c - 1)
        cur = 0.0
        for _ in range(c):
            steps.append(start + round(cur))
            cur += stride
        start += size
    return sorted(set(steps))


class Tables:
    """fp64 tables of a (possibly respaced) chain; element → fp32 after indexing (extract_and_expand)."""

    def __init__(self, steps=1000, respacing=None):
        base = linear_betas(steps)
        keep = list(range(steps)) if not respacing else spaced_steps(steps, respacing)
        acp_base = np.cumprod(1.0 - base)
        betas, last = [], 1.0
        for i in keep:
            betas.append(1 - acp_base[i] / last)
            last = acp_base[i]
        b = np.array(betas)
        self.timestep_map = keep
        self.original_steps = steps
        self.betas = b
        alphas = 1.0 - b
        self.acp = np.cumprod(alphas)
        self.acp_prev = np.append(1.0, self.acp[:-1])
        self.sqrt_recip = np.sqrt(1.0 / self.acp)
        self.sqrt_recipm1 = np.sqrt(1.0 / self.acp - 1)
        self.post_var = b * (1.0 - self.acp_prev) / (1.0 - self.acp)
        self.post_logvar_clipped = np.log(np.append(self.post_var[1], self.post_var[1:]))
        self.coef1 = b * np.sqrt(self.acp_prev) / (1.0 - self.acp)
        self.coef2 = (1.0 - self.acp_prev) * np.sqrt(alphas) / (1.0 - self.acp)
        self.T = len(b)

    def at(self, idx):
        return dict(c1=f32(self.sqrt_recip[idx]), c2=f32(self.sqrt_recipm1[idx]), p1=f32(self.coef1[idx]),
                    p2=f32(self.coef2[idx]), max_log=f32(np.log(self.betas)[idx]),
                    min_log=f32(self.post_logvar_clipped[idx]), acp=f32(self.acp[idx]), acp_prev=f32(self.acp_prev[idx]),
                    beta=float(self.betas[idx]),
                    model_t=float(f32(self.timestep_map[idx]) * f32(1000.0 / self.original_steps)))


# ------------------------------------------------------------------------------------------------
# posterior update  (posterior_mean_variance.py:110-129, :230-242; gaussian_diffusion.py:468-509)
# ------------------------------------------------------------------------------------------------
def x0_from_eps(x, eps, k, clip=True):
    pre = k["c1"] * x - k["c2"] * eps                     # predict_xstart :120-123 (mul, mul, sub in fp32)
    return (np.clip(pre, f32(-1), f32(1)) if clip else pre), pre


def mean_consts(tables, idx, mean_type="epsilon"):
    """Per-step scalars of the reference's three mean processors (posterior_mean_variance.py:45-129), all written
    as x̂₀ = c1·x − c2·out:  epsilon (c1, c2) = (√(1/ᾱ), √(1/ᾱ−1));  start_x x̂₀ = out (:86-92);  previous_x
    x̂₀ = f32(1/coef1)·out − f32(coef2/coef1)·x and the posterior mean is the model output (:57-65)."""
    k = dict(tables.at(idx))
    k["mean_is_output"] = False
    if mean_type == "start_x":
        k["c1"], k["c2"] = f32(0.0), f32(-1.0)
    elif mean_type == "previous_x":
        k["c1"] = -f32(tables.coef2[idx] / tables.coef1[idx])
        k["c2"] = -f32(1.0 / tables.coef1[idx])
        k["mean_is_output"] = True
    elif mean_type != "epsilon":
        raise NameError(mean_type)
    return k


def ddpm_sample(x, eps, v, z, k, idx, clip=True):
    x0, _ = x0_from_eps(x, eps, k, clip)
    mean = eps if k.get("mean_is_output") else k["p1"] * x0 + k["p2"] * x   # q_posterior_mean :110-118 / :62-65
    if idx == 0:
        return mean, x0                                   # no noise when t == 0, gaussian_diffusion.py:473
    frac = (v + f32(1.0)) / f32(2.0)                      # learned_range :239
    logvar = frac * k["max_log"] + (f32(1) - frac) * k["min_log"]
    return mean + np.exp(f32(0.5) * logvar) * z, x0      # :474


def ddim_sample(x, eps, z, k, idx, clip=True, eta=0.0):
    x0, _ = x0_from_eps(x, eps, k, clip)
    eps2 = (k["c1"] * x - x0) / k["c2"]                   # predict_eps_from_x_start :506-509
    one = f32(1)
    sigma = f32(eta) * np.sqrt((one - k["acp_prev"]) / (one - k["acp"])) * np.sqrt(one - k["acp"] / k["acp_prev"])
    s = x0 * np.sqrt(k["acp_prev"]) + np.sqrt(one - k["acp_prev"] - sigma ** 2) * eps2   # :495-498
    if idx != 0:
        s = s + sigma * z
    return s.astype(f32), x0


# ------------------------------------------------------------------------------------------------
# operators  (guided_diffusion/measurements.py, util/img_utils.py, util/resizer.py, util/fastmri_utils.py)
# ------------------------------------------------------------------------------------------------
def reflect_pad(x, r):
    return np.pad(x, ((0, 0), (0, 0), (r, r), (r, r)), mode="reflect")   # nn.ReflectionPad2d(k//2)


def blur_forward(x, kernel):
    """Blurkernel.forward (util/img_utils.py:271-283): reflect-pad k//2, depthwise CROSS-correlation."""
    k = kernel.shape[0]
    r = k // 2
    xp = reflect_pad(x.astype(f32), r)
    H, W = x.shape[-2:]
    out = np.zeros(x.shape, dtype=f32)
    for a, b in zip(*np.nonzero(kernel)):
        out += f32(kernel[a, b]) * xp[..., a:a + H, b:b + W]
    return out


def blur_adjoint(u, kernel):
    """Exact adjoint of blur_forward: full correlation-transpose to the padded size,
    then fold the reflected borders back (reflect without edge repeat)."""
    k = kernel.shape[0]
    r = k // 2
    H, W = u.shape[-2:]
    t = np.zeros(u.shape[:-2] + (H + 2 * r, W + 2 * r), dtype=np.float64)
    for a, b in zip(*np.nonzero(kernel)):
        t[..., a:a + H, b:b + W] += np.float64(kernel[a, b]) * u
    # fold: padded index p ↔ image index reflect(p − r)
    rows = np.abs(np.arange(-r, H + r))
    rows = np.where(rows >= H, 2 * (H - 1) - rows, rows)
    cols = np.abs(np.arange(-r, W + r))
    cols = np.where(cols >= W, 2 * (W - 1) - cols, cols)
    tmp = np.zeros(u.shape[:-2] + (H, W + 2 * r), dtype=np.float64)
    for p in range(H + 2 * r):
        tmp[..., rows[p], :] += t[..., p, :]
    g = np.zeros(u.shape, dtype=np.float64)
    for q in range(W + 2 * r):
        g[..., :, cols[q]] += tmp[..., :, q]
    return g.astype(f32)


def cubic(x):
    a = np.abs(x)
    return ((1.5 * a ** 3 - 2.5 * a ** 2 + 1) * (a <= 1) + (-0.5 * a ** 3 + 2.5 * a ** 2 - 4 * a + 2) * ((1 < a) & (a <= 2)))


def resizer_contributions(in_len, out_len, scale):
    """Resizer.contributions (util/resizer.py:104-167) for the antialiased cubic kernel."""
    width = 4.0 / scale
    out_coord = np.arange(1, out_len + 1)
    match = (out_coord - (out_len - in_len * scale) / 2) / scale + 0.5 * (1 - 1 / scale)
    left = np.floor(match - width / 2)
    ew = math.ceil(width) + 2
    fov = np.squeeze(np.int16(np.expand_dims(left, 1) + np.arange(ew) - 1))
    w = scale * cubic(scale * (1.0 * np.expand_dims(match, 1) - fov - 1))
    s = np.sum(w, axis=1)
    s[s == 0] = 1.0
    w = 1.0 * w / np.expand_dims(s, 1)
    mirror = np.uint(np.concatenate((np.arange(in_len), np.arange(in_len - 1, -1, step=-1))))
    fov = mirror[np.mod(fov, mirror.shape[0])]
    nz = np.nonzero(np.any(w, axis=0))
    return np.squeeze(w[:, nz]).astype(f32), np.squeeze(fov[:, nz]).astype(np.int64)   # (out, taps) each


def resize_matrix(in_len, out_len, scale):
    w, fov = resizer_contributions(in_len, out_len, scale)
    A = np.zeros((out_len, in_len), dtype=np.float64)
    for j in range(out_len):
        np.add.at(A[j], fov[j], w[j].astype(np.float64))
    return A


def resize_forward(x, scale):
    """Resizer.forward (util/resizer.py:55-74): gather-multiply-sum along W (dim 3) then H (dim 2)."""
    H, W = x.shape[-2:]
    oh, ow = int(np.ceil(H * scale)), int(np.ceil(W * scale))
    ww, fw = resizer_contributions(W, ow, scale)
    wh, fh = resizer_contributions(H, oh, scale)
    t = np.zeros(x.shape[:-1] + (ow,), dtype=f32)
    for k in range(ww.shape[1]):
        t += x[..., fw[:, k]] * ww[:, k]
    out = np.zeros(x.shape[:-2] + (oh, ow), dtype=f32)
    for k in range(wh.shape[1]):
        out += t[..., fh[:, k], :] * wh[:, k][:, None]
    return out


def resize_adjoint(g, scale, H, W):
    Ah = resize_matrix(H, g.shape[-2], scale)
    Aw = resize_matrix(W, g.shape[-1], scale)
    return (Ah.T @ g.astype(np.float64) @ Aw).astype(f32)   # A_hᵀ·G·A_w (App. A.4)


def inpaint_forward(x, mask):
    return (x * mask).astype(f32)                           # measurements.py:158-162


def phase_forward(x, pad):
    """PhaseRetrievalOperator.forward (measurements.py:186-189) = fftshift(|fft2(pad(x), ortho)|)."""
    p = np.pad(x.astype(np.float64), ((0, 0), (0, 0), (pad, pad), (pad, pad)))
    F = np.fft.fft2(np.fft.ifftshift(p, axes=(-2, -1)), norm="ortho")
    return np.abs(np.fft.fftshift(F, axes=(-2, -1))).astype(f32)


def phase_vjp(x, g_out, pad):
    """Jᵀ g of phase_forward at x, zero where |Z| = 0."""
    H, W = x.shape[-2:]
    p = np.pad(x.astype(np.float64), ((0, 0), (0, 0), (pad, pad), (pad, pad)))
    Z = np.fft.fftshift(np.fft.fft2(np.fft.ifftshift(p, axes=(-2, -1)), norm="ortho"), axes=(-2, -1))
    mag = np.abs(Z)
    phase = np.where(mag > 0, Z / np.where(mag > 0, mag, 1), 0)
    back = np.fft.fftshift(np.fft.ifft2(np.fft.ifftshift(g_out * phase, axes=(-2, -1)), norm="ortho"), axes=(-2, -1))
    return np.real(back)[..., pad:pad + H, pad:pad + W].astype(f32)


# ---- projections (measurements.py:48-54, :90-91, :167-168; condition_methods.py:72-82 `projection`) ---------------
def ortho_project(x, forward, transpose=lambda u: u):
    """LinearOperator.ortho_project: (I − AᵀA)x with the reference's `transpose` — the IDENTITY for the blur and
    inpainting operators (measurements.py:115-116, :147-148, :164-165), nearest up-sampling for super-resolution."""
    return (x - transpose(forward(x))).astype(f32)


def project(x, measurement, forward, transpose=lambda u: u):
    """LinearOperator.project: ortho_project(measurement) − A(x)."""
    return (ortho_project(measurement, forward, transpose) - forward(x)).astype(f32)


def nearest_upsample(u, s):
    return np.repeat(np.repeat(u, s, axis=-2), s, axis=-1)   # F.interpolate(scale_factor=s), default mode 'nearest'


def sr_project(x, measurement, scale_factor):
    """SuperResolutionOperator.project (measurements.py:90-91): x − up(A x) + up(y)."""
    ax = resize_forward(x, 1.0 / scale_factor)
    return (x - nearest_upsample(ax, scale_factor) + nearest_upsample(measurement, scale_factor)).astype(f32)


# ------------------------------------------------------------------------------------------------
# guidance  (guided_diffusion/condition_methods.py:33-60, :101-106, :145-195, :206-212)
# ------------------------------------------------------------------------------------------------
def particle_norms(r):
    flat = r.reshape(r.shape[0], -1).astype(np.float64)
    return np.sqrt((flat ** 2).sum(-1)).astype(f32), np.abs(flat).sum(-1).astype(f32)


def guidance_cotangent(r, adjoint, pre, mode, scale, extra=None, clip=True):
    """g_pre = 1[−1 ≤ pre ≤ 1] ⊙ (coef·Aᵀr + extra); coef = −scale/‖r‖ (norm) or −2·scale (norm²)."""
    norm, _ = particle_norms(r)
    if mode == "norm":
        coef = np.where(norm > 0, -f32(scale) / np.where(norm > 0, norm, 1), 0).astype(f32)
    else:
        coef = np.full(norm.shape, -2.0 * scale, dtype=f32)
    g = coef[:, None, None, None] * adjoint(r)
    if extra is not None:
        g = g + extra
    if clip:
        g = g * ((pre >= -1) & (pre <= 1))
    return g.astype(f32), norm


def guided_update(sample, g_pre, vjp, k):
    """x' = sample − (c1·g_pre − c2·VJP_ε(g_pre))  (chain rule through c1·x − c2·ε(x))."""
    grad = k["c1"] * g_pre - (k["c2"] * vjp if vjp is not None else f32(0))
    return (sample - grad).astype(f32)


# ------------------------------------------------------------------------------------------------
# device noise of the throughput mode (no reference counterpart: the reference calls torch.randn_like, :472).  Restated
# from the published algorithm: Philox4x32-10 (Salmon, Moraes, Dror, Shaw: "Parallel random numbers: as easy as 1, 2, 3",
# SC'11; Random123 constants) and the Box–Muller transform, with the counter/key layout of include/dpsttc.h.
# ------------------------------------------------------------------------------------------------
def philox4x32_10(counter, key):
    """counter: (n,4) uint32, key: (2,) uint32 → (n,4) uint32."""
    M0, M1, W0, W1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57), 0x9E3779B9, 0xBB67AE85
    c = counter.astype(np.uint64)
    k0, k1 = int(key[0]), int(key[1])
    mask = np.uint64(0xFFFFFFFF)
    for _ in range(10):
        p0, p1 = M0 * c[:, 0], M1 * c[:, 2]
        hi0, lo0, hi1, lo1 = p0 >> np.uint64(32), p0 & mask, p1 >> np.uint64(32), p1 & mask
        c = np.stack([hi1 ^ c[:, 1] ^ np.uint64(k0), lo1, hi0 ^ c[:, 3] ^ np.uint64(k1), lo0], axis=1)
        k0, k1 = (k0 + W0) & 0xFFFFFFFF, (k1 + W1) & 0xFFFFFFFF
    return c.astype(np.uint32)


def philox_normal(seed, step, particle, n_elems):
    """The z of one particle (n_elems values, a multiple of 4): element group i4 ↦ Philox(counter = (i4 lo, i4 hi, particle,
    step), key = (seed lo, seed hi)) → u = ((r >> 8) + ½)·2⁻²⁴ → Box–Muller pairs (√(−2 ln u₀)·cos θ, ·sin θ, …) with
    θ = 2π(u₁ − ½)."""
    i4 = np.arange(n_elems // 4, dtype=np.uint64)
    ctr = np.stack([i4 & np.uint64(0xFFFFFFFF), i4 >> np.uint64(32), np.full_like(i4, particle & 0xFFFFFFFF),
                    np.full_like(i4, step & 0xFFFFFFFF)], axis=1).astype(np.uint32)
    r = philox4x32_10(ctr, np.array([seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF], dtype=np.uint32))
    u = ((r >> np.uint32(8)).astype(np.float64) + 0.5) * 2.0 ** -24
    m0, m1 = np.sqrt(-2.0 * np.log(u[:, 0])), np.sqrt(-2.0 * np.log(u[:, 2]))
    t0, t1 = 2 * np.pi * (u[:, 1] - 0.5), 2 * np.pi * (u[:, 3] - 0.5)
    z = np.stack([m0 * np.cos(t0), m0 * np.sin(t0), m1 * np.cos(t1), m1 * np.sin(t1)], axis=1)
    return z.reshape(-1).astype(f32)


# ------------------------------------------------------------------------------------------------
# reweighting / resampling  (gaussian_diffusion.py:537-552, :685-698; torch CPU multinomial kernel)
# ------------------------------------------------------------------------------------------------
def logweights(meas, sem=None, tau=0.01, meas_scale=1.0, meas_pow=1, sem_scale=0.0, sem_pow=1):
    cost = f32(meas_scale) * (meas.astype(f32) ** meas_pow)
    if sem is not None:
        cost = cost + f32(sem_scale) * (sem.astype(f32) ** sem_pow)
    return (-f32(tau) * cost).astype(f32)


def weights_cdf(logw, linear=True):
    """fp32 sequential cumulative sum, fp32 divide — the arithmetic of torch.multinomial on CPU
    (pinned against torch in tests/test_oracle_pins.py)."""
    logw = logw.astype(f32)
    shift = f32(0) if linear else logw.max()
    w = np.exp(logw - shift).astype(f32)
    cum = np.zeros(len(w), dtype=f32)
    s = f32(0)
    for j in range(len(w)):
        s = f32(s + w[j])
        cum[j] = s
    degenerate = (not (s > 0)) or (not np.isfinite(s)) or (w.max() == w.min())
    if s > 0 and np.isfinite(s):
        cdf = (cum / s).astype(f32)
        cdf[-1] = f32(1)
        wn = (w / s).astype(f32)
    else:
        cdf = (np.arange(1, len(w) + 1) / len(w)).astype(f32)
        wn = np.full(len(w), 1 / len(w), dtype=f32)
    return wn, cdf, degenerate


def search(cdf, u):
    out = np.zeros(len(u), dtype=np.int64)
    for i, x in enumerate(u):
        lo, hi = 0, len(cdf)
        while hi - lo > 0:
            mid = lo + (hi - lo) // 2
            if np.float64(cdf[mid]) < x:
                lo = mid + 1
            else:
                hi = mid
        out[i] = min(lo, len(cdf) - 1)
    return out


def ancestors_multinomial(cdf, uniforms, degenerate=False):
    if degenerate:
        return np.arange(len(uniforms), dtype=np.int64)
    return search(cdf, np.asarray(uniforms, dtype=np.float64))


def ancestors_systematic(cdf, u0, n, degenerate=False):
    """One uniform u0: positions (i + u0)/n, same CDF and search rule (SURVEY §8 'no reference counterpart')."""
    if degenerate:
        return np.arange(n, dtype=np.int64)
    return search(cdf, (np.arange(n, dtype=np.float64) + np.float64(u0)) / np.float64(n))


# ------------------------------------------------------------------------------------------------
# DiffStateGrad projection  (guided_diffusion/diffstategrad_utils.py; hook gaussian_diffusion.py:240-255)
# ------------------------------------------------------------------------------------------------
def diffstategrad_rank(singular_values, cutoff):
    """compute_rank_for_explained_variance as called (:41): ONE (C, r) array in the list → flattened cumsum, then /3."""
    sq = np.asarray(singular_values) ** 2
    cum = np.cumsum(sq) / np.sum(sq)
    return int((int(np.searchsorted(cum, cutoff)) + 1) / 3)


def diffstategrad_update(sample, grad, cutoff=0.99):
    """x' = sample − U_r U_rᵀ grad[0] V_r V_rᵀ per channel, SVD of sample[0]; the batch-1 result broadcasts (:255)."""
    U, sv, Vh = np.linalg.svd(sample[0].astype(np.float32), full_matrices=False)
    r = diffstategrad_rank(sv, cutoff)
    A, B = U[:, :, :r], Vh[:, :r, :]
    low = np.matmul(np.matmul(A.transpose(0, 2, 1), grad[0]), B.transpose(0, 2, 1))
    proj = np.matmul(np.matmul(A, low), B).astype(np.float32)
    return (sample - proj[None]).astype(np.float32), r


def greedy_best(costs):
    return int(np.argmin(costs))                            # torch.argmin: first minimum (:631)


def best_of_n(distances):
    """best_of_n_simple.py:32-41: per image, argmin over the first n+1 paths, for every n."""
    d = np.asarray(distances)
    return np.stack([np.argmin(d[:, :n + 1], axis=1) for n in range(d.shape[1])], axis=1)


def psnr(real, fake):
    """compute_psnr_manual (compute_metrics.py:93-98): 20·log10(1/√mean((real − fake)²)), one value per particle of
    `fake` (the drivers call it with one path at a time, sample_condition_batched_ttc.py:191)."""
    real = np.asarray(real, dtype=np.float32)
    fake = np.asarray(fake, dtype=np.float32)
    d = (fake - real).reshape(fake.shape[0], -1).astype(np.float64)
    mse = (d * d).mean(axis=1)
    return (20.0 * np.log10(1.0 / np.sqrt(mse))).astype(np.float32)


def measurement_distance(y, forward, samples):
    """‖y − A(sample)‖₂ per particle (the drivers' y_space / the loops' measurement distance, gaussian_diffusion.py:303)."""
    r = y - forward(samples)
    return particle_norms(r)[0]
