// Register-resident column kernel of the fused phase-retrieval guidance (included by phase_impl.cuh once per transform length).
//
// phase_cols_fused keeps every intermediate of its two column transforms in shared memory: staging store, three stages that
// each read and write the whole sequence, an |F| pass and a residual pass that read and write it again — about 10 reads and
// 10 writes of every 8-byte element plus 1.7 twiddle loads per element and transform, at 75 % thread efficiency (ncu:
// 44.8 M warp instructions for 96 planes, short-scoreboard + MIO-throttle stalls 44 %, 8.6 M bank conflicts in the residual
// pass).  Here a thread OWNS a butterfly: its inputs arrive in registers, the radix kernel runs in registers, and shared
// memory is only the exchange between two stages (ping-pong buffers A / B, one barrier per exchange):
//
//   global Rt ─► regs ─dft8─► A ─► regs ─tw·dft8─► B ─► regs ─tw·dft_last─► F(k1) in regs
//        |F|, unit phase, y at both Hermitian-mirrored positions, residual, Σr², Σ|r|, cotangent × unit phase ─► A
//   A ─► regs ─dft8─► B ─► regs ─tw·dft8─► A ─► regs ─tw·dft_last─► rows 64..64+H of T (global)
//
// 5 exchanges (one write + one read each) instead of 20 passes; the radix-8 twiddles W64^{k·r} come from a 64-entry table
// indexed [r][k] (conflict-free) and the last-stage twiddles from a full L-entry table (no half-table sign logic; its values
// are built from the half table with the rule of twid(), so both column kernels multiply by the same numbers).
// Thread roles (512 threads, 8 spectrum columns per CTA):
//   J role (radix-8 stages, threads 0 .. 8·L/8 − 1; whole warps beyond that idle):  column f = tid / (L/8), butterfly j = tid % (L/8)
//          → global loads of a warp are 256 consecutive bytes of one column;
//   F role (last stage + epilogue / final store, all 512 threads): column f = tid & 7, butterfly j = tid >> 3
//          → 8 consecutive lanes touch 8 consecutive k2: 32-byte runs of y / r_out, 64-byte runs of T (as before).
// Sequence stride kLQ ≡ 2 (mod 16) elements: both roles are bank-conflict-free on 64-bit accesses (J: strides 9 and 1 inside a
// sequence; F: 2·f + {P(j), P(j)+1} are 16 distinct 8-byte banks per half-warp).
//
// The arithmetic (dft8 / dft_last / cmul, twiddle values, Stockham index maps) is that of fft_batch, so spectra agree with the
// shared-memory kernel to the last bit up to FMA contraction; the partial sums are accumulated in a different order.
// Every function below is one barrier-free phase of the kernel taking the thread index as an argument:
// tests/emu/phase_cols_emu.cpp runs the same phases thread by thread on the CPU against a plain DFT (index logic check).
#ifndef PHASE_PACKED
#define PHASE_PACKED 0  // 0: scalar fp32 butterflies.  1: complex adds / subtracts as FADD2 (same values), twiddle products scalar.
#endif                  // 2: also the twiddle products packed (FMUL2 + FFMA2 on (w, −i·w) quadruples; measured slower)
#if PHASE_PACKED == 2
typedef float4 tw_t;  // (w.x, w.y, −w.y, w.x)
DPS_DEV float2 cr_cmul(float2 a, float4 w) { return cmul_tw(a, w); }
DPS_DEV void cr_dft8(float2* v) { dft8p(v); }
DPS_DEV void cr_dft_last(float2* v) { dft_last_p<true>(v); }
#elif PHASE_PACKED == 1
typedef float2 tw_t;
DPS_DEV float2 cr_cmul(float2 a, float2 w) { return cmul(a, w); }
DPS_DEV void cr_dft8(float2* v) { dft8p(v); }
DPS_DEV void cr_dft_last(float2* v) { dft_last_p<false>(v); }
#else
typedef float2 tw_t;
DPS_DEV float2 cr_cmul(float2 a, float2 w) { return cmul(a, w); }
DPS_DEV void cr_dft8(float2* v) { dft8(v); }
DPS_DEV void cr_dft_last(float2* v) { dft_last(v); }
#endif
constexpr int kT2 = 512;
constexpr int kSeq = 8;
constexpr int kJ = kSeq * kL8;                                   // 384 / 256 / 192 threads in the J role
constexpr int kLQ = ((kL + kL / 8 - 2 + 15) / 16) * 16 + 2;     // 434 / 290 / 226
static_assert(kSeq == kColsPerCta && kJ <= kT2 && kJ % 32 == 0 && kSeq * 64 == kT2, "thread roles");
static_assert(kLQ % 16 == 2 && kLQ > kL - 1 + (kL - 1) / 8, "sequence stride");

struct ColsCtx {
  float2* A;
  float2* B;
  const tw_t* tw;     // exp(−2πi j/L), j < L (full table: the upper half is −tw[j − L/2], the rule of twid())
  const tw_t* w64;    // W64^{k·r} at [r·8 + k]
  const float2* rt;   // Rt[k2][row] of this plane
  const float* y;     // measurement plane (L×L), never null (dps_operator_guidance requires it)
  float* outp;        // residual plane (L×L) or null
  float2* t;          // T[row][k2] of this plane
  int k20, ncols;
};
struct ColsRegs {
  float2 v[8];
  float sq, ab;
};
struct ColsY {
  float y1[kR3], y2[kR3];
};

// J role: the 8 inputs j + (L/8)·r of the first radix-8 stage straight from the row kernel's scratch; rows outside
// [64, 64 + H) are the zero padding (whole r-slices of it are dropped at compile time).
DPS_DEV void cr_load(int tid, ColsRegs& R, const ColsCtx& c) {
  if (tid >= kJ) return;
  const int f = tid / kL8, j = tid - f * kL8;
  if (f >= c.ncols) return;  // (the last column group holds ONE column: its other sequences are never touched)
  const float2* src = c.rt + (int64_t)(c.k20 + f) * kImg;
#pragma unroll
  for (int r = 0; r < 8; ++r) {
    if (kL8 * r + kL8 - 1 < kPad || kL8 * r >= kPad + kImg) {
      R.v[r] = make_float2(0.f, 0.f);
    } else {
      const int row = j + kL8 * r - kPad;
      R.v[r] = (row >= 0 && row < kImg) ? ldg_stream2(src + row) : make_float2(0.f, 0.f);
    }
  }
}
// J role: the inputs of the first stage from a buffer in natural order (second transform)
DPS_DEV void cr_read_a(int tid, ColsRegs& R, const float2* buf, int nseq = kSeq) {
  if (tid >= kJ) return;
  const int f = tid / kL8, j = tid - f * kL8;
  if (f >= nseq) return;
  const float2* src = buf + f * kLQ + P(j);  // P(j + (L/8)·r) = P(j) + kS8·r
#pragma unroll
  for (int r = 0; r < 8; ++r) R.v[r] = src[kS8 * r];
}
// J role, stage 1 (R = 8, Ns = 1, no twiddles): out[8j + r] = DFT8(in[j + (L/8)·r]);  P(8j + r) = 9j + r
DPS_DEV void cr_stage_a(int tid, ColsRegs& R, float2* dstbuf, int nseq = kSeq) {
  if (tid >= kJ) return;
  const int f = tid / kL8, j = tid - f * kL8;
  if (f >= nseq) return;
  cr_dft8(R.v);
  float2* dst = dstbuf + f * kLQ + 9 * j;
#pragma unroll
  for (int r = 0; r < 8; ++r) dst[r] = R.v[r];
}
// J role, stage 2 (R = 8, Ns = 8): twiddle W64^{k·r}, k = j & 7;  out[64·(j>>3) + k + 8r];  P(·) = 72·(j>>3) + k + 9r
DPS_DEV void cr_stage_b(int tid, ColsRegs& R, const float2* srcbuf, float2* dstbuf, const tw_t* w64, int nseq = kSeq) {
  if (tid >= kJ) return;
  const int f = tid / kL8, j = tid - f * kL8;
  if (f >= nseq) return;
  const int k = j & 7;
  const float2* src = srcbuf + f * kLQ + P(j);
#pragma unroll
  for (int r = 0; r < 8; ++r) {
    R.v[r] = src[kS8 * r];
    if (r) R.v[r] = cr_cmul(R.v[r], w64[8 * r + k]);
  }
  cr_dft8(R.v);
  float2* dst = dstbuf + f * kLQ + 72 * (j >> 3) + k;
#pragma unroll
  for (int r = 0; r < 8; ++r) dst[9 * r] = R.v[r];
}
// F role, stage 3 (R = L/64, Ns = 64): twiddle exp(−2πi·j·r/L); the thread ends up with bins j + 64r in natural order
DPS_DEV void cr_stage_c(int tid, float2* v, const float2* srcbuf, const tw_t* tw) {
  const int f = tid & 7, j = tid >> 3;
  const float2* src = srcbuf + f * kLQ + P(j);  // P(j + 64r) = P(j) + 72r
#pragma unroll
  for (int r = 0; r < kR3; ++r) {
    v[r] = src[72 * r];
    if (r) v[r] = cr_cmul(v[r], tw[j * r]);
  }
  cr_dft_last(v);
}
// F role: the measurement at the direct output position shift(k1, k2) and at the mirror shift(−k1, −k2) of the thread's bins
DPS_DEV void cr_yload(int tid, ColsY& Y, const ColsCtx& c) {
  const int f = tid & 7, j = tid >> 3;
  const int k2 = c.k20 + f;
  const bool act = f < c.ncols;
  const int c1 = shift_idx(k2), c2 = shift_idx(k2 ? kL - k2 : 0);
#pragma unroll
  for (int r = 0; r < kR3; ++r) {
    const int k1 = j + 64 * r;
    // (volatile loads: the compiler otherwise sinks them below the barrier they are meant to be in flight across)
    Y.y1[r] = act ? ldg_ro_pinned(c.y + shift_idx(k1) * kL + c1) : 0.f;
    Y.y2[r] = act ? ldg_ro_pinned(c.y + shift_idx(k1 ? kL - k1 : 0) * kL + c2) : 0.f;
  }
}
// (Inactive sequences of the last column group hold whatever was in shared memory; every use is guarded by `act`.)
// F role: last stage of the first transform, then everything that is local to a bin: |F|/L, the unit phase conj(F)/|F|, the
// residual at both mirrored output positions (partial sums; r itself only if asked for) and the symmetrised cotangent
// ½(r(k) + r(−k)) × unit phase → dstbuf in natural order (the input of the second transform).
template <bool kOut>
DPS_DEV void cr_epilogue(int tid, ColsRegs& R, const ColsY& Y, const ColsCtx& c) {
  cr_stage_c(tid, R.v, c.B, c.tw);
  const int f = tid & 7, j = tid >> 3;
  const int k2 = c.k20 + f;
  const bool act = f < c.ncols, mir = act && k2 > 0 && k2 < kL / 2;
  const int c1 = shift_idx(k2), c2 = shift_idx(k2 ? kL - k2 : 0);
  const float inv_l = 1.0f / (float)kL;
  float2* dst = c.A + f * kLQ + P(j);
#pragma unroll
  for (int r = 0; r < kR3; ++r) {
    const int k1 = j + 64 * r;
    const float2 F = R.v[r];
    // |F| and 1/|F| without the slow paths of sqrtf and the division: rsqrt.approx refined by one Newton step, then
    // |F| = m2·r corrected by its residual (as in phase_cols_fused)
    const float m2 = fmaf(F.x, F.x, F.y * F.y);
    float q = rsqrtf(m2);
    q = fmaf(q, fmaf(-0.5f * m2 * q, q, 0.5f), q);
    float mag = m2 * q;
    mag = fmaf(fmaf(-mag, mag, m2), 0.5f * q, mag);
    const bool nz = m2 > 0.f;
    const float inv = nz ? q : 0.f;
    const float a = nz ? mag * inv_l : 0.f;
    const float r1 = __fsub_rn(Y.y1[r], a);
    const float r2 = __fsub_rn(Y.y2[r], a);
    if (act) {
      R.sq = fmaf(r1, r1, R.sq);
      R.ab += fabsf(r1);
      if constexpr (kOut) stg_stream(c.outp + shift_idx(k1) * kL + c1, r1);
    }
    if (mir) {  // for the self-conjugate columns k2 = 0, L/2 the mirrored output is another bin of the same column
      R.sq = fmaf(r2, r2, R.sq);
      R.ab += fabsf(r2);
      if constexpr (kOut) stg_stream(c.outp + shift_idx(k1 ? kL - k1 : 0) * kL + c2, r2);
    }
    const float gs = act ? 0.5f * (r1 + r2) : 0.f;
    dst[72 * r] = make_float2(gs * (F.x * inv), gs * (-F.y * inv));  // the association of phase_cols_fused: same bits
  }
}
// F role, forward (two-kernel) path: last stage, |F|/L, the output y − |F|/L (or |F|/L itself) at both mirrored positions, partial
// sums, and the unit phase conj(F)/|F| → ph[k2][k1] for the adjoint's column kernel (32-byte runs: 8 lanes = 8 columns, 4 bins each)
template <bool kHasY>
DPS_DEV void cr_fwd_epilogue(int tid, ColsRegs& R, const ColsY& Y, const ColsCtx& c, float2* ph) {
  cr_stage_c(tid, R.v, c.B, c.tw);
  const int f = tid & 7, j = tid >> 3;
  const int k2 = c.k20 + f;
  const bool act = f < c.ncols, mir = act && k2 > 0 && k2 < kL / 2;
  const int c1 = shift_idx(k2), c2 = shift_idx(k2 ? kL - k2 : 0);
  const float inv_l = 1.0f / (float)kL;
  float2* php = ph + (int64_t)k2 * kL + j;
#pragma unroll
  for (int r = 0; r < kR3; ++r) {
    const int k1 = j + 64 * r;
    const float2 F = R.v[r];
    const float m2 = fmaf(F.x, F.x, F.y * F.y);
    float q = rsqrtf(m2);
    q = fmaf(q, fmaf(-0.5f * m2 * q, q, 0.5f), q);
    float mag = m2 * q;
    mag = fmaf(fmaf(-mag, mag, m2), 0.5f * q, mag);
    const bool nz = m2 > 0.f;
    const float inv = nz ? q : 0.f;
    const float a = nz ? mag * inv_l : 0.f;
    const float r1 = kHasY ? __fsub_rn(Y.y1[r], a) : a;
    const float r2 = kHasY ? __fsub_rn(Y.y2[r], a) : a;
    if (act) {
      stg_stream2(reinterpret_cast<float*>(php + 64 * r), make_float2(F.x * inv, -F.y * inv));
      R.sq = fmaf(r1, r1, R.sq);
      R.ab += fabsf(r1);
      if (c.outp) stg_stream(c.outp + shift_idx(k1) * kL + c1, r1);
    }
    if (mir) {
      R.sq = fmaf(r2, r2, R.sq);
      R.ab += fabsf(r2);
      if (c.outp) stg_stream(c.outp + shift_idx(k1 ? kL - k1 : 0) * kL + c2, r2);
    }
  }
}
// F role, adjoint (two-kernel) path: the symmetrised cotangent ½(g(k) + g(−k)) of the thread's bins (32-byte runs of the residual
// plane, as the measurement loads above) × the unit phase the forward pass left in ph[k2][k1] → dstbuf in natural order
DPS_DEV void ca_load(int tid, const float* rplane, const float2* ph, const ColsCtx& c) {
  const int f = tid & 7, j = tid >> 3;
  const int k2 = c.k20 + f;
  const bool act = f < c.ncols;
  const int c1 = shift_idx(k2), c2 = shift_idx(k2 ? kL - k2 : 0);
  const float2* php = ph + (int64_t)k2 * kL + j;
  float g1[kR3], g2[kR3];
  float2 u[kR3];
#pragma unroll
  for (int r = 0; r < kR3; ++r) {
    const int k1 = j + 64 * r;
    g1[r] = act ? ldg_stream(rplane + shift_idx(k1) * kL + c1) : 0.f;
    g2[r] = act ? ldg_stream(rplane + shift_idx(k1 ? kL - k1 : 0) * kL + c2) : 0.f;
    u[r] = act ? ldg_stream2(php + 64 * r) : make_float2(0.f, 0.f);
  }
  float2* dst = c.A + f * kLQ + P(j);
#pragma unroll
  for (int r = 0; r < kR3; ++r) {
    const float g = 0.5f * (g1[r] + g2[r]);
    dst[72 * r] = make_float2(g * u[r].x, g * u[r].y);
  }
}
// F role: last stage of the second transform; padded rows 64 .. 64 + H − 1 go to T[row][k2] (row stride L/2 + 1)
DPS_DEV void cr_store(int tid, ColsRegs& R, const ColsCtx& c) {
  cr_stage_c(tid, R.v, c.A, c.tw);
  const int f = tid & 7, j = tid >> 3;
  if (f >= c.ncols) return;
  float2* dst = c.t + c.k20 + f;
#pragma unroll
  for (int r = 0; r < kR3; ++r) {
    if (64 * r >= kPad && 64 * r + 63 < kPad + kImg)  // whole 64-row slices: kPad = 64 and H is a multiple of 64
      stg_stream2(reinterpret_cast<float*>(dst + (int64_t)(j + 64 * r - kPad) * kHalf), R.v[r]);
  }
}
static_assert(kPad == 64 && kImg % 64 == 0, "cr_store keeps whole 64-row slices");
