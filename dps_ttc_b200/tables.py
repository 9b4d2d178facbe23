"""Host-side builders of the immutable tables an operator plan uploads once: blur kernels, Resizer
bands, inpainting masks.  Numerics follow the reference so that the tables are bit-identical:
  util/img_utils.py:286-299   Blurkernel.weights_init (scipy gaussian_filter of a delta)
  util/resizer.py:104-167     Resizer.contributions (antialiased cubic, mirror-index reflection)
  util/img_utils.py:164-235   random_sq_bbox / mask_generator (numpy global RNG, same draw order)
The motion-blur kernel generator of the reference is the un-vendored `motionblur` package
(measurements.py:8,104; parity unpinned): motion_kernel() below is a stand-in with the
same interface (size, intensity → (k,k) matrix summing to 1, numpy global RNG).
"""
from __future__ import annotations

import math

import numpy as np


# ------------------------------------------------------------------------------------------------
# blur kernels
# ------------------------------------------------------------------------------------------------
def gaussian_kernel(kernel_size: int, std: float) -> np.ndarray:
    """fp64 (k,k) kernel exactly as Blurkernel.weights_init builds it (cast to fp32 by the conv weight)."""
    import scipy.ndimage
    delta = np.zeros((kernel_size, kernel_size))
    delta[kernel_size // 2, kernel_size // 2] = 1
    return scipy.ndimage.gaussian_filter(delta, sigma=std)


def motion_kernel(kernel_size: int, intensity: float, rng=None) -> np.ndarray:
    """Random thin-path blur kernel: a random walk whose step-to-step turning grows with `intensity`,
    rasterised with bilinear splatting and normalised to sum 1.  Draws from numpy's global RNG (or
    `rng`) like the reference's Kernel(size, intensity) does, so `np.random.seed(kernel_idx)` pins it."""
    rs = np.random if rng is None else rng
    k = int(kernel_size)
    diag = math.hypot(k, k)
    max_len = 0.75 * diag * (rs.uniform() + rs.uniform(0, intensity ** 2))
    steps = int(rs.randint(3, 7) if intensity > 0 else 1)  # number of segments
    angle = rs.uniform(0, 2 * math.pi)
    max_turn = rs.uniform(0, intensity * math.pi)
    pts = [np.zeros(2)]
    seg = max_len / steps
    for _ in range(steps):
        angle += rs.triangular(-max_turn, 0, max_turn) if max_turn > 0 else 0.0
        length = seg * rs.uniform(0.5, 1.5)
        pts.append(pts[-1] + length * np.array([math.cos(angle), math.sin(angle)]))
    pts = np.array(pts)
    pts -= (pts.max(0) + pts.min(0)) / 2  # centre the path
    half = (k - 1) / 2
    scale = min(1.0, (half - 1) / max(1e-9, np.abs(pts).max()))
    pts = pts * scale + half
    img = np.zeros((k, k), dtype=np.float64)
    for a, b in zip(pts[:-1], pts[1:]):
        n_sub = max(2, int(4 * np.linalg.norm(b - a)) + 1)
        for t in np.linspace(0, 1, n_sub):
            x, y = a + t * (b - a)
            x0, y0 = int(math.floor(x)), int(math.floor(y))
            fx, fy = x - x0, y - y0
            for dy, wy in ((0, 1 - fy), (1, fy)):
                for dx, wx in ((0, 1 - fx), (1, fx)):
                    yy, xx = y0 + dy, x0 + dx
                    if 0 <= yy < k and 0 <= xx < k:
                        img[yy, xx] += wy * wx
    # 8-bit quantisation like a rasterised path: most of the canvas is exactly zero
    img = np.round(img / img.max() * 255.0)
    return (img / img.sum()).astype(np.float64)


class MotionKernel:
    """Duck type of motionblur.motionblur.Kernel as the reference uses it (.kernelMatrix)."""

    def __init__(self, size=(61, 61), intensity=0.5):
        self.SIZE = tuple(size)
        self.INTENSITY = intensity
        self.kernelMatrix = motion_kernel(size[0], intensity)


# ------------------------------------------------------------------------------------------------
# Resizer bands
# ------------------------------------------------------------------------------------------------
def _cubic(x):
    a = np.abs(x)
    a2, a3 = a * a, a * a * a
    near = (1.5 * a3 - 2.5 * a2 + 1) * (a <= 1)
    far = (-0.5 * a3 + 2.5 * a2 - 4 * a + 2) * ((1 < a) & (a <= 2))
    return near + far


def _linear(x):
    return (x + 1) * ((-1 <= x) & (x < 0)) + (1 - x) * ((0 <= x) & (x <= 1))


def _box(x):
    return ((-0.5 <= x) & (x < 0.5)) * 1.0


def _lanczos(n):
    def f(x):
        eps = np.finfo(np.float32).eps
        return ((np.sin(math.pi * x) * np.sin(math.pi * x / n) + eps) / ((math.pi ** 2 * x ** 2 / n) + eps)) * (abs(x) < n)
    return f


_KERNELS = {"cubic": (_cubic, 4.0), None: (_cubic, 4.0), "linear": (_linear, 2.0), "box": (_box, 1.0),
            "lanczos2": (_lanczos(2), 4.0), "lanczos3": (_lanczos(3), 6.0)}


def resizer_band(in_len: int, out_len: int, scale: float, kernel=None, antialiasing: bool = True):
    """(fov, weights), both (taps, out_len): out[j] = Σ_k weights[k, j] · in[fov[k, j]]."""
    fn, width = _KERNELS[kernel]
    if antialiasing and scale < 1:
        base = fn
        fn = lambda t: scale * base(scale * t)  # noqa: E731  stretched low-pass kernel
        width = width / scale
    centre = (np.arange(1, out_len + 1) - (out_len - in_len * scale) / 2) / scale + 0.5 * (1 - 1 / scale)
    left = np.floor(centre - width / 2)
    taps = int(math.ceil(width)) + 2
    fov = np.int16(left[:, None] + np.arange(taps) - 1)  # int16 like the reference (wrap-around included)
    w = fn(1.0 * centre[:, None] - fov - 1)
    norm = w.sum(axis=1)
    norm[norm == 0] = 1.0
    w = 1.0 * w / norm[:, None]
    mirror = np.uint(np.concatenate((np.arange(in_len), np.arange(in_len - 1, -1, -1))))
    fov = mirror[np.mod(fov, mirror.shape[0])]
    keep = np.nonzero(np.any(w, axis=0))
    w = np.squeeze(w[:, keep])
    fov = np.squeeze(fov[:, keep])
    return np.ascontiguousarray(fov.T.astype(np.int32)), np.ascontiguousarray(w.T.astype(np.float32))


def resizer_tables(in_shape, scale_factor: float, kernel=None, antialiasing: bool = True):
    """Bands for the H (dim 2) and W (dim 3) axes of an (N,C,H,W) input scaled by `scale_factor` < 1."""
    H, W = int(in_shape[-2]), int(in_shape[-1])
    s = float(scale_factor)
    out_h, out_w = int(np.uint(np.ceil(H * s))), int(np.uint(np.ceil(W * s)))
    fov_h, w_h = resizer_band(H, out_h, s, kernel, antialiasing)
    fov_w, w_w = resizer_band(W, out_w, s, kernel, antialiasing)
    return (fov_h, w_h), (fov_w, w_w), (out_h, out_w)


# ------------------------------------------------------------------------------------------------
# inpainting masks
# ------------------------------------------------------------------------------------------------
class MaskGenerator:
    """mask_generator of the reference (util/img_utils.py:184-235): same numpy draws in the same
    order, so under the same np.random.seed the mask is bit-identical.  Returns a float32 numpy
    array shaped like `img` ((B,C,H,W), values in {0,1})."""

    def __init__(self, mask_type, mask_len_range=None, mask_prob_range=None, image_size=256, margin=(16, 16)):
        if mask_type not in ("box", "random", "both", "extreme"):
            raise AssertionError(f"unknown mask type {mask_type}")
        self.mask_type = mask_type
        self.mask_len_range = mask_len_range
        self.mask_prob_range = mask_prob_range
        self.image_size = image_size
        self.margin = margin

    def _box(self, shape):
        lo, hi = (int(v) for v in self.mask_len_range)
        mask_h = np.random.randint(lo, hi)
        mask_w = np.random.randint(lo, hi)
        mh, mw = self.margin
        top = np.random.randint(mh, self.image_size - mh - mask_h)
        left = np.random.randint(mw, self.image_size - mw - mask_w)
        mask = np.ones(shape, dtype=np.float32)
        mask[..., top:top + mask_h, left:left + mask_w] = 0
        return mask

    def _random(self, shape):
        size = self.image_size
        lo, hi = self.mask_prob_range
        prob = np.random.uniform(lo, hi)
        flat = np.ones(size * size, dtype=np.float32)
        gone = np.random.choice(size * size, int(size * size * prob), replace=False)
        flat[gone] = 0
        mask = np.empty(shape, dtype=np.float32)
        mask[...] = flat.reshape(1, 1, size, size)
        return mask

    def __call__(self, img):
        shape = tuple(img.shape)
        if self.mask_type == "random":
            return self._random(shape)
        if self.mask_type == "box":
            return self._box(shape)
        if self.mask_type == "extreme":
            return 1.0 - self._box(shape)
        return None  # 'both' returns None in the reference as well (img_utils.py:225-235)
