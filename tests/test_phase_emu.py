"""Index logic of the register-resident phase column kernel (dps_ttc_b200/csrc/phase_colsreg.cuh) on the CPU.

The kernel is written as barrier-free per-thread phases; tests/emu/phase_cols_emu.cpp compiles the SAME headers for the host
and runs the phases thread by thread.  Compared here with a plain numpy DFT of the same column step: |F|/L, the residual at
both Hermitian-mirrored output positions, the partial sums and the second transform of the symmetrised cotangent × unit
phase.  (Test infrastructure; the product path is the CUDA kernel, checked against the oracle by the -m gpu tests.)"""
import ctypes as C
import os
import shutil
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _build(r3, tmp_path, which="cols", packed=0):
    if shutil.which("g++") is None:
        pytest.skip("g++ not available")
    so = str(tmp_path / f"phase_{which}_emu_{r3}_{packed}.so")
    subprocess.run(["g++", "-O1", "-ffp-contract=off", "-shared", "-fPIC", f"-DPHASE_R3={r3}", f"-DPHASE_PACKED={packed}", "-I",
                    os.path.join(ROOT, "dps_ttc_b200", "csrc"), "-o", so,
                    os.path.join(ROOT, "tests", "emu", f"phase_{which}_emu.cpp")], check=True)
    return C.CDLL(so)


@pytest.mark.parametrize("packed", [0, 1, 2])
@pytest.mark.parametrize("r3,want_r", [(6, True), (4, True), (3, True), (6, False)])
def test_register_column_kernel_matches_numpy_dft(r3, want_r, packed, tmp_path):
    lib = _build(r3, tmp_path, packed=packed)
    dims = (C.c_int * 4)()
    lib.emu_dims(dims)
    L, img, half, groups = list(dims)
    assert L == 64 * r3 and img == L - 128 and half == L // 2 + 1
    rng = np.random.default_rng(7 + r3)
    rt = (rng.standard_normal((half, img)) + 1j * rng.standard_normal((half, img))).astype(np.complex64) * 8
    with_y = True   # dps_operator_guidance always has the measurement
    y = (rng.random((L, L)) * 1.5).astype(np.float32)
    r_out = np.full((L, L), np.nan, np.float32) if want_r else None
    t = np.full((img, half), np.nan + 0j, np.complex64)
    partials = np.zeros((groups, 2), np.float32)
    fp = lambda a: a.ctypes.data_as(C.c_void_p) if a is not None else None
    assert lib.emu_cols(fp(rt), fp(y), fp(r_out), fp(t), fp(partials)) == 0

    # plain restatement: column k2 of the half spectrum, rows zero-padded by 64
    pad = np.zeros((half, L), np.complex128)
    pad[:, 64:64 + img] = rt
    F = np.fft.fft(pad, axis=1)                       # F[k2, k1]
    amp = np.abs(F) / L
    sh = lambda k: (k + L // 2) % L
    k1 = np.arange(L)
    k2 = np.arange(half)
    ya = y.astype(np.float64) if with_y else None
    pos1 = (sh(k1)[None, :], sh(k2)[:, None])                          # (row, col) of the direct output
    pos2 = (sh((L - k1) % L)[None, :], sh((L - k2) % L)[:, None])      # mirrored output
    r1 = (ya[pos1] - amp) if with_y else amp
    r2 = (ya[pos2] - amp) if with_y else amp
    mir = ((k2 > 0) & (k2 < L // 2))[:, None] & np.ones((1, L), bool)
    ref_out = np.full((L, L), np.nan)
    ref_out[pos1[0].repeat(half, 0), pos1[1].repeat(L, 1)] = r1
    m_r, m_c = np.broadcast_to(pos2[0], (half, L))[mir], np.broadcast_to(pos2[1], (half, L))[mir]
    ref_out[m_r, m_c] = r2[mir]
    assert not np.isnan(ref_out).any(), "the two position sets must tile the L×L plane"
    scale = max(1.0, np.abs(ref_out).max())
    if want_r:
        assert np.abs(r_out - ref_out).max() <= 2e-5 * scale
    # partial sums per column group
    sq = np.where(mir, r1 ** 2 + r2 ** 2, r1 ** 2).sum(1)
    ab = np.where(mir, np.abs(r1) + np.abs(r2), np.abs(r1)).sum(1)
    for g in range(groups):
        cols = slice(8 * g, min(8 * g + 8, half))
        assert abs(partials[g, 0] - sq[cols].sum()) <= 1e-4 * sq[cols].sum()
        assert abs(partials[g, 1] - ab[cols].sum()) <= 1e-4 * ab[cols].sum()
    # cotangent: ½(r(k) + r(−k)) · conj(F)/|F|, second (forward) transform over k1, rows 64..64+img
    with np.errstate(invalid="ignore", divide="ignore"):
        unit = np.where(np.abs(F) > 0, np.conj(F) / np.abs(F), 0)
    Hs = 0.5 * (r1 + r2) * unit
    T = np.fft.fft(Hs, axis=1)[:, 64:64 + img].T      # T[row, k2]
    assert not np.isnan(t.view(np.float32)).any()
    assert np.abs(t - T).max() <= 2e-5 * np.abs(T).max()


@pytest.mark.parametrize("packed", [0, 1, 2])
@pytest.mark.parametrize("r3,clip", [(6, True), (4, True), (3, False)])
def test_register_row_kernels_match_numpy_dft(r3, clip, packed, tmp_path):
    """K1 (x, ε → x̂₀ → half spectrum of every image row, two rows per complex FFT, + clamp-pass bytes) and K3 (Hermitian half
    rows → real rows × coefficient × clamp mask) of phase_rowsreg.cuh, emulated thread by thread."""
    lib = _build(r3, tmp_path, "rows", packed=packed)
    dims = (C.c_int * 4)()
    lib.emu_dims(dims)
    L, img, half, groups = list(dims)
    assert L == 64 * r3 and img * 0 + groups == img // 16
    rng = np.random.default_rng(11 + r3)
    c1, c2 = np.float32(1.7), np.float32(0.9)
    x = rng.standard_normal((img, img)).astype(np.float32)
    eps = rng.standard_normal((img, img)).astype(np.float32)
    rt = np.full((half, img), np.nan + 0j, np.complex64)
    maskb = np.full((img, img), 7, np.uint8)
    fp = lambda a: a.ctypes.data_as(C.c_void_p)
    lib.emu_rows_fwd.argtypes = [C.c_void_p, C.c_void_p, C.c_float, C.c_float, C.c_int, C.c_void_p, C.c_void_p]
    assert lib.emu_rows_fwd(fp(x), fp(eps), float(c1), float(c2), int(clip), fp(rt), fp(maskb)) == 0
    pre = c1 * x - c2 * eps                                   # float32, every product rounded on its own
    x0 = np.clip(pre, -1, 1) if clip else pre
    pad = np.zeros((img, L)); pad[:, 64:64 + img] = x0
    ref = np.fft.fft(pad, axis=1)[:, :half].T                 # Rt[k2][row]
    assert not np.isnan(rt.view(np.float32)).any()
    assert np.abs(rt - ref).max() <= 2e-5 * np.abs(ref).max()
    want_mask = ((pre >= -1) & (pre <= 1)).astype(np.uint8) if clip else np.ones((img, img), np.uint8)
    assert np.array_equal(maskb, want_mask)

    # forward path of operator.forward(x): no ε (x̂₀ = x, never clamped), no mask bytes
    rt2 = np.full((half, img), np.nan + 0j, np.complex64)
    assert lib.emu_rows_fwd(fp(x), None, float(c1), float(c2), int(clip), fp(rt2), None) == 0
    pad2 = np.zeros((img, L)); pad2[:, 64:64 + img] = x
    ref2 = np.fft.fft(pad2, axis=1)[:, :half].T
    assert not np.isnan(rt2.view(np.float32)).any()
    assert np.abs(rt2 - ref2).max() <= 2e-5 * np.abs(ref2).max()

    t = (rng.standard_normal((img, half)) + 1j * rng.standard_normal((img, half))).astype(np.complex64) * 4
    coef = 1.0 / L
    lib.emu_rows_adj.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_float, C.c_float, C.c_void_p, C.c_float, C.c_void_p]
    # X[k] = T1[k] + i T2[k] for k <= L/2, conj-mirrored above (row pairs); z = FFT(X); Re -> even row, Im -> odd row
    k = np.arange(L)
    kk = np.where(k < half, k, L - k)
    full = np.where((k < half)[None, :], t[:, kk], np.conj(t[:, kk])).astype(np.complex128)
    X = full[0::2] + 1j * full[1::2]
    z = np.fft.fft(X, axis=1)[:, 64:64 + img]
    rows = np.empty((img, img))
    rows[0::2], rows[1::2] = z.real, z.imag
    extra = (rng.standard_normal((img, img)) * 0.1).astype(np.float32)
    # (a) fused path: mask bytes, no extra  (b) two-kernel path: mask recomputed from x and eps, extra term, another coefficient
    # (c) no mask at all
    for maskb, mx, me, ex, cf, want in ((want_mask, None, None, None, coef, rows * coef * want_mask),
                                        (None, x, eps, extra, -0.37 * coef, (rows * (-0.37 * coef) + extra) * ((pre >= -1) & (pre <= 1))),
                                        (None, None, None, None, coef, rows * coef)):
        g = np.full((img, img), np.nan, np.float32)
        assert lib.emu_rows_adj(fp(t), fp(maskb) if maskb is not None else None, fp(mx) if mx is not None else None,
                                fp(me) if me is not None else None, float(c1), float(c2),
                                fp(ex) if ex is not None else None, cf, fp(g)) == 0
        assert not np.isnan(g).any()
        assert np.abs(g - want).max() <= 2e-5 * np.abs(want).max()


@pytest.mark.parametrize("r3,with_y", [(6, True), (6, False), (4, True), (3, False)])
def test_register_forward_column_kernel_matches_numpy_dft(r3, with_y, tmp_path):
    """phase_cols_fwd_reg (forward pass of the two-kernel path): y − |F|/L (or |F|/L) at both mirrored positions, partial sums
    and the unit phase conj(F)/|F| laid out [k2][k1] for the adjoint."""
    lib = _build(r3, tmp_path)
    dims = (C.c_int * 4)()
    lib.emu_dims(dims)
    L, img, half, groups = list(dims)
    rng = np.random.default_rng(23 + r3)
    rt = (rng.standard_normal((half, img)) + 1j * rng.standard_normal((half, img))).astype(np.complex64) * 8
    y = (rng.random((L, L)) * 1.5).astype(np.float32) if with_y else None
    out = np.full((L, L), np.nan, np.float32)
    ph = np.full((half, L), np.nan + 0j, np.complex64)
    partials = np.zeros((groups, 2), np.float32)
    fp = lambda a: a.ctypes.data_as(C.c_void_p) if a is not None else None
    assert lib.emu_cols_fwd(fp(rt), fp(y), fp(out), fp(ph), fp(partials)) == 0
    pad = np.zeros((half, L), np.complex128)
    pad[:, 64:64 + img] = rt
    F = np.fft.fft(pad, axis=1)
    amp = np.abs(F) / L
    sh = lambda k: (k + L // 2) % L
    k1, k2 = np.arange(L), np.arange(half)
    ref = np.full((L, L), np.nan)
    r1 = (y[sh(k1)[None, :], sh(k2)[:, None]] - amp) if with_y else amp
    r2 = (y[sh((L - k1) % L)[None, :], sh((L - k2) % L)[:, None]] - amp) if with_y else amp
    mir = ((k2 > 0) & (k2 < L // 2))[:, None] & np.ones((1, L), bool)
    ref[np.broadcast_to(sh(k1)[None, :], (half, L)), np.broadcast_to(sh(k2)[:, None], (half, L))] = r1
    ref[np.broadcast_to(sh((L - k1) % L)[None, :], (half, L))[mir], np.broadcast_to(sh((L - k2) % L)[:, None], (half, L))[mir]] = r2[mir]
    assert not np.isnan(ref).any() and not np.isnan(out).any()
    assert np.abs(out - ref).max() <= 2e-5 * max(1.0, np.abs(ref).max())
    unit = np.conj(F) / np.abs(F)
    assert not np.isnan(ph.view(np.float32)).any()
    assert np.abs(ph - unit).max() <= 1e-4                      # |unit| = 1; bins with tiny |F| lose relative accuracy
    sq = np.where(mir, r1 ** 2 + r2 ** 2, r1 ** 2).sum(1)
    for g in range(groups):
        cols = slice(8 * g, min(8 * g + 8, half))
        assert abs(partials[g, 0] - sq[cols].sum()) <= 1e-4 * sq[cols].sum()


@pytest.mark.parametrize("r3", [6, 4, 3])
def test_register_adjoint_column_kernel_matches_numpy_dft(r3, tmp_path):
    """phase_cols_adj_reg (adjoint of the two-kernel path): symmetrised cotangent × stored unit phase → column transform → T."""
    lib = _build(r3, tmp_path)
    dims = (C.c_int * 4)()
    lib.emu_dims(dims)
    L, img, half, groups = list(dims)
    rng = np.random.default_rng(31 + r3)
    rplane = rng.standard_normal((L, L)).astype(np.float32)
    ph = np.exp(1j * rng.uniform(0, 2 * np.pi, (half, L))).astype(np.complex64)
    t = np.full((img, half), np.nan + 0j, np.complex64)
    fp = lambda a: a.ctypes.data_as(C.c_void_p)
    assert lib.emu_cols_adj(fp(rplane), fp(ph), fp(t)) == 0
    sh = lambda k: (k + L // 2) % L
    k1, k2 = np.arange(L), np.arange(half)
    g1 = rplane[sh(k1)[None, :], sh(k2)[:, None]].astype(np.float64)
    g2 = rplane[sh((L - k1) % L)[None, :], sh((L - k2) % L)[:, None]].astype(np.float64)
    Hs = 0.5 * (g1 + g2) * ph.astype(np.complex128)
    T = np.fft.fft(Hs, axis=1)[:, 64:64 + img].T
    assert not np.isnan(t.view(np.float32)).any()
    assert np.abs(t - T).max() <= 2e-5 * np.abs(T).max()
