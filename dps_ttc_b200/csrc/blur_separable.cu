// Separable blur (Gaussian): ReflectionPad2d(k/2) + depthwise cross-correlation with a rank-1 kernel
// (measurements.py:129-149, util/img_utils.py:268-308), and its exact adjoint.
//
// Formulation.  In 1-D, "reflect-pad then correlate" is the banded matrix
//     A[i][m] = Σ_{d : reflect(i+d) = m} w[d+r],   |i − m| ≤ r,
// which equals the plain taps w[m−i+r] except next to each border, where the mirrored taps fold back
// onto the band (rows i < r of A; rows m ≤ r of Aᵀ — column r of A still receives w[0] from row 0).
// Forward applies A_v ⊗ A_h, the adjoint A_vᵀ ⊗ A_hᵀ: the SAME kernel with different tap tables — interior
// taps as kernel parameters (constant bank → uniform registers), border rows from a small table in shared
// memory.  No padded image is ever materialised.
//
// One CTA = one (particle, channel, strip of 32 output rows), 256 threads:
//   0. stage the strip + r halo rows in shared memory, 128-bit loads, x̂₀ = clamp(c1·x − c2·ε) applied on the
//      fly (forward) — the only global read of the particle;
//   1. vertical pass: a thread owns TWO adjacent columns and 16 rows; every FMA is a packed FFMA2
//      (fma.rn.f32x2, new on sm_100) on a (colA, colB) pair with the tap broadcast from a uniform register —
//      32 LDS.64 feed 200 FFMA2 = 400 FMAs.  Results are stored row-pair interleaved;
//   2. horizontal pass: a thread owns a ROW PAIR and 4 adjacent columns, again FFMA2 on (rowA, rowB) pairs,
//      (4+2r)/2 LDS.128 per 100 FFMA2.  The 2·(r/4+1) border quads of a row take a separate, warp-uniform
//      phase with table taps, so no warp executes both paths;
//   3. epilogue in registers: residual y − A x̂₀ + per-CTA Σr², Σ|r|   (forward)
//                            clamp mask ⊙ (coef·Aᵀr + extra)          (adjoint).
// Roofline: 2·(2r+1) FMA per pixel (50 for σ=3) against 12-16 B per pixel.  With FFMA2 the issue-slot floor
// (≈0.2 µs/particle) drops below the HBM floor (≈0.36 µs/particle): HBM-bound by design, see DESIGN.md.
// Radii above 16 (σ > 4) fall back to the scalar-FMA variant of the same algorithm (sep1_kernel).
#include <vector>

#include "operator.cuh"

namespace {

constexpr int kThreads = 256;
constexpr int kRows = 32;  // output rows per CTA
constexpr int kGroup = 8;  // vertical outputs per register block

template <int R>
struct SepParams {
  float wv[2 * R + 1];  // interior vertical taps, index e+R multiplies input row i+e
  float wh[2 * R + 1];
  const float* bv;  // border rows: (2(R+4), 2R+1); rows [0,R+4) top, then rows [L-R-4, L) bottom
  const float* bh;
  int C, H, W;
  int strips;  // ceil(H / kRows)
};

struct SepSet {
  std::vector<float> wv, wh;  // (2R+1)
  float* bv = nullptr;        // device
  float* bh = nullptr;
};

}  // namespace

struct SepTables {
  int R = 0;  // template radius (multiple of 4, >= true radius)
  SepSet fwd, adj;
};

namespace {

template <int R>
DPS_DEV const float* border_row(const float* table, int i, int L) {  // i is a border index: i < R+4 or i >= L-R-4
  return table + (i < R + 4 ? i : (R + 4) + (i - (L - R - 4))) * (2 * R + 1);
}

// ---- epilogue shared by every variant: 4 adjacent outputs of one row ------------------------------
template <bool kAdjoint>
DPS_DEV void sep_epilogue(float4 o, int n, int64_t off, int64_t nchw, const FwdArgs& fa, const AdjArgs& aa,
                          float& sq, float& ab) {
  if (!kAdjoint) {
    float4 res = o;
    if (fa.y) {
      const float4 yv = ldg_ro4(fa.y + n * fa.y_stride + off);
      res = make_float4(__fsub_rn(yv.x, o.x), __fsub_rn(yv.y, o.y), __fsub_rn(yv.z, o.z), __fsub_rn(yv.w, o.w));
    }
    stg_stream4(fa.out + n * nchw + off, res);
    sq += res.x * res.x + res.y * res.y + res.z * res.z + res.w * res.w;
    ab += fabsf(res.x) + fabsf(res.y) + fabsf(res.z) + fabsf(res.w);
  } else {
    const float coef = aa.coef ? aa.coef[n] : 1.0f;
    float4 res = make_float4(coef * o.x, coef * o.y, coef * o.z, coef * o.w);
    if (aa.extra) {
      const float4 e = ldg_stream4(aa.extra + n * aa.extra_stride + off);
      res.x += e.x; res.y += e.y; res.z += e.z; res.w += e.w;
    }
    const float4 pass = mask_load4(aa.mask_src, aa.has_mask, n, off);
    res.x *= pass.x; res.y *= pass.y; res.z *= pass.z; res.w *= pass.w;
    stg_stream4(aa.g + n * aa.g_stride + off, res);
  }
}

// Same epilogue split in two so that its global loads (y | extra, x, ε) are issued BEFORE the FMA block of an
// item and their latency hides under ~100 FFMA2 instead of stalling the thread at the end of every item.
struct RowIO {
  float4 a, mx, me;  // a: y (forward) or extra (adjoint); mx, me: clamp-mask sources
};
template <bool kAdjoint>
DPS_DEV RowIO row_prefetch(bool live, int n, int64_t off, const FwdArgs& fa, const AdjArgs& aa) {
  RowIO io;
  const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
  io.a = io.mx = io.me = z;
  if (!live) return io;
  if (!kAdjoint) {
    if (fa.y) io.a = ldg_ro4(fa.y + n * fa.y_stride + off);
  } else {
    if (aa.extra) io.a = ldg_stream4(aa.extra + n * aa.extra_stride + off);
    if (aa.has_mask && aa.mask_src.eps && aa.mask_src.clip) {
      io.mx = ldg_stream4(aa.mask_src.x + n * aa.mask_src.x_stride + off);
      io.me = ldg_stream4(aa.mask_src.eps + n * aa.mask_src.eps_stride + off);
    }
  }
  return io;
}
template <bool kAdjoint>
DPS_DEV void row_finish(float4 o, const RowIO& io, bool live, int n, int64_t off, int64_t nchw, const FwdArgs& fa,
                        const AdjArgs& aa, float& sq, float& ab) {
  if (!live) return;
  if (!kAdjoint) {
    float4 res = o;
    if (fa.y) res = make_float4(__fsub_rn(io.a.x, o.x), __fsub_rn(io.a.y, o.y), __fsub_rn(io.a.z, o.z), __fsub_rn(io.a.w, o.w));
    stg_stream4(fa.out + n * nchw + off, res);
    sq += res.x * res.x + res.y * res.y + res.z * res.z + res.w * res.w;
    ab += fabsf(res.x) + fabsf(res.y) + fabsf(res.z) + fabsf(res.w);
  } else {
    const float coef = aa.coef ? aa.coef[n] : 1.0f;
    float4 res = make_float4(coef * o.x + io.a.x, coef * o.y + io.a.y, coef * o.z + io.a.z, coef * o.w + io.a.w);
    if (aa.has_mask && aa.mask_src.eps && aa.mask_src.clip) {
      const float c1 = aa.mask_src.c1, c2 = aa.mask_src.c2;
      res.x *= clamp_pass(x0_pre(io.mx.x, io.me.x, c1, c2)); res.y *= clamp_pass(x0_pre(io.mx.y, io.me.y, c1, c2));
      res.z *= clamp_pass(x0_pre(io.mx.z, io.me.z, c1, c2)); res.w *= clamp_pass(x0_pre(io.mx.w, io.me.w, c1, c2));
    }
    stg_stream4(aa.g + n * aa.g_stride + off, res);
  }
}

// ---- staging shared by both variants: rows [r0-R, r0+kRows+R) of one plane, zero outside the image ---
template <int R, bool kAdjoint>
DPS_DEV void sep_stage(float* tile, int row_stride, int col_off, int r0, int n, int64_t plane, int C, int H, int W,
                       const FwdArgs& fa, const AdjArgs& aa) {
  const float* x;
  const float* eps = nullptr;
  float c1 = 1.f, c2 = 0.f;
  int clip = 0;
  if (kAdjoint) {
    x = aa.r + (int64_t)n * C * H * W + plane;
  } else {
    x = fa.src.x + n * fa.src.x_stride + plane;
    if (fa.src.eps) eps = fa.src.eps + n * fa.src.eps_stride + plane;
    c1 = fa.src.c1; c2 = fa.src.c2; clip = fa.src.clip;
  }
  // Batches of kB float4 per thread: every global load of a batch is issued before the first use, so a
  // thread keeps 2·kB 16-byte requests in flight (the CTA ≈ 57 KB) instead of one dependent pair at a time.
  const int w4 = W / 4;
  const int total = (kRows + 2 * R) * w4;
  constexpr int kB = 7;
  const int dtr = kThreads / w4, dq = kThreads - dtr * w4;  // (row, quad) advance per kThreads items, no division
  int tr = threadIdx.x / w4, q = threadIdx.x - tr * w4;
#pragma unroll 1
  for (int i0 = threadIdx.x; i0 < total; i0 += kB * kThreads) {
    float4 xv[kB], ev[kB];
    int dst[kB];
#pragma unroll
    for (int b = 0; b < kB; ++b) {
      const int row = r0 - R + tr;
      const bool live = (i0 + b * kThreads < total);
      const bool inside = live && row >= 0 && row < H;
      dst[b] = live ? tr * row_stride + col_off + q * 4 : -1;
      const int64_t off = (int64_t)(inside ? row : 0) * W + q * 4;
      xv[b] = inside ? ldg_stream4(x + off) : make_float4(0.f, 0.f, 0.f, 0.f);
      ev[b] = (inside && eps) ? ldg_stream4(eps + off) : make_float4(0.f, 0.f, 0.f, 0.f);
      q += dq; tr += dtr;
      if (q >= w4) { q -= w4; ++tr; }
    }
#pragma unroll
    for (int b = 0; b < kB; ++b) {
      if (dst[b] < 0) continue;
      float4 v = xv[b];
      if (eps) {
        v.x = x0_of(xv[b].x, ev[b].x, c1, c2, clip); v.y = x0_of(xv[b].y, ev[b].y, c1, c2, clip);
        v.z = x0_of(xv[b].z, ev[b].z, c1, c2, clip); v.w = x0_of(xv[b].w, ev[b].w, c1, c2, clip);
      }
      *reinterpret_cast<float4*>(tile + dst[b]) = v;
    }
  }
}

// =================================================================================================
// FFMA2 variant (R ≤ 16)
// =================================================================================================
template <int R, int WT, bool kAdjoint>  // WT: compile-time image width (0 = runtime p.W) so that tile offsets fold into immediates
__global__ void __launch_bounds__(kThreads, 2) sep2_kernel(const SepParams<R> p, const FwdArgs fa, const AdjArgs aa) {
  extern __shared__ __align__(16) float smem[];
  const int H = p.H, W = WT ? WT : p.W;
  const int W2 = W / 2;                  // float2 per tile row
  const int VW = W + 2 * R;              // float2 per V2 row (column halo of R each side)
  constexpr int kBorder = 2 * (R + 4) * (2 * R + 1);
  float* tile = smem;                                                    // (kRows+2R, W) floats
  float2* V2 = reinterpret_cast<float2*>(tile + (kRows + 2 * R) * W);    // (kRows/2, VW) row-pair interleaved
  float* bvs = reinterpret_cast<float*>(V2 + (kRows / 2) * VW);
  float* bhs = bvs + kBorder;
  float* red = bhs + kBorder;  // 64 floats

  const int strip = blockIdx.x % p.strips;
  const int c = blockIdx.x / p.strips;
  const int n = blockIdx.y;
  const int r0 = strip * kRows;
  const int tid = threadIdx.x;
  const int64_t plane = (int64_t)c * H * W;
  const int64_t nchw = (int64_t)p.C * H * W;

  stage_async(bvs, p.bv, kBorder, tid, kThreads);  // border tables: asynchronous, awaited with the tile
  stage_async(bhs, p.bh, kBorder, tid, kThreads);
  sep_stage<R, kAdjoint>(tile, W, 0, r0, n, plane, p.C, H, W, fa, aa);
  for (int i = tid; i < (kRows / 2) * 2 * R; i += kThreads) {  // zero the column halos of V2
    const int rp = i / (2 * R), q = i - rp * (2 * R);
    V2[rp * VW + (q < R ? q : W + q)] = make_float2(0.f, 0.f);
  }
  stage_wait();
  __syncthreads();

  // ---- phase 1: vertical pass, two columns per thread, 16 rows per thread ------------------------
  {
    const int half = tid >> 7;  // rows [16·half, 16·half+16)
    const float2* tile2 = reinterpret_cast<const float2*>(tile);
    for (int c2 = tid & 127; c2 < W2; c2 += 128) {
#pragma unroll 1
      for (int g = 0; g < 2; ++g) {
        const int g0 = half * (kRows / 2) + g * kGroup;
        float2 in2[kGroup + 2 * R];
#pragma unroll
        for (int s = 0; s < kGroup + 2 * R; ++s) in2[s] = tile2[(g0 + s) * W2 + c2];
        float2 acc[kGroup];
        const int row0 = r0 + g0;  // uniform across the CTA half
        if (row0 >= R + 4 && row0 + kGroup - 1 < H - R - 4) {  // whole group interior: taps from uniform registers
#pragma unroll
          for (int o = 0; o < kGroup; ++o) {
            float2 a = make_float2(0.f, 0.f);
#pragma unroll
            for (int k = 0; k <= 2 * R; ++k) a = __ffma2_rn(make_float2(p.wv[k], p.wv[k]), in2[o + k], a);
            acc[o] = a;
          }
        } else {
#pragma unroll
          for (int o = 0; o < kGroup; ++o) {
            const int row = row0 + o;
            float2 a = make_float2(0.f, 0.f);
            if (row < H) {
              const bool interior = row >= R + 4 && row < H - R - 4;
              const float* bt = interior ? nullptr : border_row<R>(bvs, row, H);
#pragma unroll
              for (int k = 0; k <= 2 * R; ++k) {
                const float w = interior ? p.wv[k] : bt[k];
                a = __ffma2_rn(make_float2(w, w), in2[o + k], a);
              }
            }
            acc[o] = a;
          }
        }
        // store row-pair interleaved: V2[rp][col] = (row 2rp, row 2rp+1); two adjacent columns = one 16-byte store
#pragma unroll
        for (int o = 0; o < kGroup; o += 2) {
          float4 st = make_float4(acc[o].x, acc[o + 1].x, acc[o].y, acc[o + 1].y);
          *reinterpret_cast<float4*>(V2 + ((g0 + o) >> 1) * VW + R + 2 * c2) = st;
        }
      }
    }
  }
  __syncthreads();

  // ---- phase 2: horizontal pass, a row pair × 4 columns per item ------------------------------
  float sq = 0.f, ab = 0.f;
  int q_lo = R / 4 + 1, q_hi = (W - 4 - R) / 4;  // interior quads [q_lo, q_hi)
  if (q_hi < q_lo) q_lo = q_hi = W / 4;          // narrow image: every column is within R+4 of a border
  const int nq_int = q_hi - q_lo;
  if (nq_int > 0) {
    const int drp = kThreads / nq_int, dqi = kThreads - drp * nq_int;  // item advance per kThreads, no division in the loop
    int rp = tid / nq_int, qi = tid - rp * nq_int;
    for (; rp < kRows / 2; rp += drp, qi += dqi) {
      if (qi >= nq_int) { qi -= nq_int; if (++rp >= kRows / 2) break; }
      const int col = (q_lo + qi) * 4;
      const int row = r0 + 2 * rp;
      const int64_t off = plane + (int64_t)row * W + col;
      const RowIO ioA = row_prefetch<kAdjoint>(row < H, n, off, fa, aa);
      const RowIO ioB = row_prefetch<kAdjoint>(row + 1 < H, n, off + W, fa, aa);
      const float2* rowp = V2 + rp * VW + col;  // window starts at image column col−R
      float2 in2[4 + 2 * R];
#pragma unroll
      for (int s = 0; s < (4 + 2 * R) / 2; ++s) {
        const float4 v = *reinterpret_cast<const float4*>(rowp + 2 * s);
        in2[2 * s] = make_float2(v.x, v.y);
        in2[2 * s + 1] = make_float2(v.z, v.w);
      }
      float2 o[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float2 a = make_float2(0.f, 0.f);
#pragma unroll
        for (int k = 0; k <= 2 * R; ++k) a = __ffma2_rn(make_float2(p.wh[k], p.wh[k]), in2[j + k], a);
        o[j] = a;
      }
      row_finish<kAdjoint>(make_float4(o[0].x, o[1].x, o[2].x, o[3].x), ioA, row < H, n, off, nchw, fa, aa, sq, ab);
      row_finish<kAdjoint>(make_float4(o[0].y, o[1].y, o[2].y, o[3].y), ioB, row + 1 < H, n, off + W, nchw, fa, aa, sq, ab);
    }
  }
  // border quads: [0, q_lo) and [q_hi, W/4) — table taps, one warp-uniform phase
  const int nq_b = q_lo + (W / 4 - q_hi);
  for (int i = tid; i < (kRows / 2) * nq_b; i += kThreads) {
    const int rp = i / nq_b, bq = i - rp * nq_b;
    const int q = bq < q_lo ? bq : q_hi + (bq - q_lo);
    const int col = q * 4;
    const int row = r0 + 2 * rp;
    const int64_t off = plane + (int64_t)row * W + col;
    const RowIO ioA = row_prefetch<kAdjoint>(row < H, n, off, fa, aa);
    const RowIO ioB = row_prefetch<kAdjoint>(row + 1 < H, n, off + W, fa, aa);
    const float2* rowp = V2 + rp * VW + col;
    float2 o[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float* bt = border_row<R>(bhs, col + j, W);
      float2 a = make_float2(0.f, 0.f);
#pragma unroll
      for (int k = 0; k <= 2 * R; ++k) {
        const float w = bt[k];
        a = __ffma2_rn(make_float2(w, w), rowp[j + k], a);
      }
      o[j] = a;
    }
    row_finish<kAdjoint>(make_float4(o[0].x, o[1].x, o[2].x, o[3].x), ioA, row < H, n, off, nchw, fa, aa, sq, ab);
    row_finish<kAdjoint>(make_float4(o[0].y, o[1].y, o[2].y, o[3].y), ioB, row + 1 < H, n, off + W, nchw, fa, aa, sq, ab);
  }
  if (!kAdjoint && fa.partials) {
    block_sum2(sq, ab, red);
    if (tid == 0) {
      float* pp = fa.partials + ((int64_t)n * (p.C * p.strips) + blockIdx.x) * 2;
      pp[0] = sq;
      pp[1] = ab;
    }
  }
}

template <int R>
size_t sep2_smem_bytes(int W) {
  return sizeof(float) * ((size_t)(kRows + 2 * R) * W + (size_t)(kRows / 2) * (W + 2 * R) * 2 +
                          2 * (size_t)(2 * (R + 4)) * (2 * R + 1) + 64);
}

// =================================================================================================
// streaming FFMA2 variant (R ≤ 16, W = 256, H % 32 == 0) — the one the 256×256 configs run
//
// The vertical pass no longer stages the input: a thread owns one column PAIR and walks its 32+2R rows straight
// from global memory (64-bit loads, a warp reads 256 contiguous bytes per row and tensor), keeping the last
// 2R+1 rows in registers.  Loads of the next 8 rows are in flight while the current 8 outputs are computed, so
// HBM streams during the FFMA2 work instead of before it, and nothing is loaded twice inside a CTA.  Border rows
// need no table: the mirrored taps are added from the same uniform registers, statically (≈R²/2 extra FFMA2
// in the first and last strip only).
// =================================================================================================
constexpr int kT3 = 128;  // threads = column pairs of a 256-wide image
#ifndef SEP3_LB
#define SEP3_LB 8
#endif
#ifndef SEP3_MINB
#define SEP3_MINB 4
#endif
constexpr int kLB = SEP3_LB;  // rows per load batch = outputs per compute batch

// Mirrored-tap correction of output row `o` of a border strip (all indices compile-time after unrolling).
// a / b = distance of the input / output row from the border, a = κ − R + b with κ the window position counted
// from the border side; forward: a ≥ 1, b ≥ 0; adjoint: a ≥ 0, b ≥ 1; a + b ≤ R;
// tap index in this direction's own array: R − (a+b) for top-forward / bottom-adjoint, R + (a+b) otherwise.
template <int R, bool kAdjoint, bool kBottom, int O>
DPS_DEV float2 border_fix(const float (&w)[2 * R + 1], const float2* win, float2 acc) {
  constexpr int b = kBottom ? (kRows - 1 - O) : O;
#pragma unroll
  for (int k = 0; k <= 2 * R; ++k) {
    const int kap = kBottom ? 2 * R - k : k;
    const int a = kap - R + b;
    const bool hit = (a >= (kAdjoint ? 0 : 1)) && (b >= (kAdjoint ? 1 : 0)) && (a + b <= R);
    const int idx = (kBottom != kAdjoint) ? R + a + b : R - a - b;
    if (hit) acc = __ffma2_rn(make_float2(w[idx], w[idx]), win[k], acc);
  }
  return acc;
}

template <int R, bool kAdjoint, int OB>
struct BatchOut {  // the 8 outputs of compute batch OB (rows 8·OB … 8·OB+7), fully static
  template <int J>
  static DPS_DEV void run(const SepParams<R>& p, const float2* in, bool is_top, bool is_bot, float2* acc) {
    constexpr int O = OB * kLB + J;
    float2 a = make_float2(0.f, 0.f);
#pragma unroll
    for (int k = 0; k <= 2 * R; ++k) a = __ffma2_rn(make_float2(p.wv[k], p.wv[k]), in[O + k], a);
    if (O <= R && is_top) a = border_fix<R, kAdjoint, false, O>(p.wv, in + O, a);
    if (kRows - 1 - O <= R && is_bot) a = border_fix<R, kAdjoint, true, O>(p.wv, in + O, a);
    acc[J] = a;
    if constexpr (J + 1 < kLB) run<J + 1>(p, in, is_top, is_bot, acc);
  }
};

template <int R, bool kAdjoint>
__global__ void __launch_bounds__(kT3, SEP3_MINB) sep3_kernel(const SepParams<R> p, const FwdArgs fa, const AdjArgs aa) {
  constexpr int W = 256, W2 = W / 2, VW = W + 2 * R;
  constexpr int NR = kRows + 2 * R;  // input rows walked by a thread
  constexpr int NB = NR / kLB;       // load batches
  constexpr int kLag = (kLB - 1 + 2 * R) / kLB;  // output batch ob reads in[kLB·ob … kLB·ob+kLB−1+2R]: input batches ≤ ob + kLag
  static_assert(NB == kRows / kLB + kLag, "every output batch must be produced inside the load loop");
  constexpr int kBorder = 2 * (R + 4) * (2 * R + 1);
  static_assert(NR % kLB == 0 && R % 4 == 0 && (kLB == 4 || kLB == 8), "radius must be a multiple of 4");
  extern __shared__ __align__(16) float smem[];
  float2* V2 = reinterpret_cast<float2*>(smem);  // (kRows/2, VW) row-pair interleaved
  float* bhs = reinterpret_cast<float*>(V2 + (kRows / 2) * VW);
  float* red = bhs + kBorder;

  const int H = p.H;
  const int strip = blockIdx.x % p.strips;
  const int c = blockIdx.x / p.strips;
  const int n = blockIdx.y;
  const int r0 = strip * kRows;
  const int tid = threadIdx.x;
  const int64_t plane = (int64_t)c * H * W;
  const int64_t nchw = (int64_t)p.C * H * W;
  const bool is_top = strip == 0, is_bot = strip == p.strips - 1;

  stage_async(bhs, p.bh, kBorder, tid, kT3);  // horizontal border table: lands while the vertical pass streams
  for (int i = tid; i < (kRows / 2) * 2 * R; i += kT3) {
    const int rp = i / (2 * R), q = i - rp * (2 * R);
    V2[rp * VW + (q < R ? q : W + q)] = make_float2(0.f, 0.f);
  }

  // ---- phase 1: vertical pass straight from global memory -----------------------------------------
  {
    const float2* xs;
    const float2* es = nullptr;
    float c1 = 1.f, c2 = 0.f;
    int clip = 0;
    if (kAdjoint) {
      xs = reinterpret_cast<const float2*>(aa.r + (int64_t)n * nchw + plane);
    } else {
      xs = reinterpret_cast<const float2*>(fa.src.x + n * fa.src.x_stride + plane);
      if (fa.src.eps) es = reinterpret_cast<const float2*>(fa.src.eps + n * fa.src.eps_stride + plane);
      c1 = fa.src.c1; c2 = fa.src.c2; clip = fa.src.clip;
    }
    const float2 zero2 = make_float2(0.f, 0.f);
    float2 in[NR];
    float2 rx[kLB], re[kLB];
    auto issue = [&](int lb) {
#pragma unroll
      for (int j = 0; j < kLB; ++j) {
        const int row = r0 - R + lb * kLB + j;
        const bool inside = row >= 0 && row < H;
        rx[j] = inside ? ldg_stream2(xs + row * W2 + tid) : zero2;
        re[j] = (inside && es) ? ldg_stream2(es + row * W2 + tid) : zero2;
      }
    };
    issue(0);
#pragma unroll
    for (int lb = 0; lb < NB; ++lb) {
      float2 cx[kLB], ce[kLB];
#pragma unroll
      for (int j = 0; j < kLB; ++j) { cx[j] = rx[j]; ce[j] = re[j]; }
      if (lb + 1 < NB) issue(lb + 1);  // next batch in flight while this one is consumed
#pragma unroll
      for (int j = 0; j < kLB; ++j) in[lb * kLB + j] = es ? x0_pair(cx[j], ce[j], c1, c2, clip) : cx[j];
      if (lb >= kLag) {
        float2 acc[kLB];
        switch (lb - kLag) {  // lb is a compile-time constant after unrolling: one case survives
          case 0: BatchOut<R, kAdjoint, 0>::template run<0>(p, in, is_top, is_bot, acc); break;
          case 1: BatchOut<R, kAdjoint, 1>::template run<0>(p, in, is_top, is_bot, acc); break;
          case 2: BatchOut<R, kAdjoint, 2>::template run<0>(p, in, is_top, is_bot, acc); break;
          case 3: BatchOut<R, kAdjoint, 3>::template run<0>(p, in, is_top, is_bot, acc); break;
#if SEP3_LB < 8
          case 4: BatchOut<R, kAdjoint, 4>::template run<0>(p, in, is_top, is_bot, acc); break;
          case 5: BatchOut<R, kAdjoint, 5>::template run<0>(p, in, is_top, is_bot, acc); break;
          case 6: BatchOut<R, kAdjoint, 6>::template run<0>(p, in, is_top, is_bot, acc); break;
          case 7: BatchOut<R, kAdjoint, 7>::template run<0>(p, in, is_top, is_bot, acc); break;
#endif
          default: break;
        }
        const int o0 = (lb - kLag) * kLB;
#pragma unroll
        for (int j = 0; j < kLB; j += 2)
          *reinterpret_cast<float4*>(V2 + ((o0 + j) >> 1) * VW + R + 2 * tid) =
              make_float4(acc[j].x, acc[j + 1].x, acc[j].y, acc[j + 1].y);
      }
    }
  }
  stage_wait();
  __syncthreads();

  // ---- phase 2: horizontal pass, a row pair × 4 columns per item (as sep2_kernel) -----------------
  float sq = 0.f, ab = 0.f;
  constexpr int q_lo = R / 4 + 1, q_hi = (W - 4 - R) / 4, nq_int = q_hi - q_lo;
  {
    constexpr int drp = kT3 / nq_int, dqi = kT3 - drp * nq_int;
    int rp = tid / nq_int, qi = tid - rp * nq_int;
    for (; rp < kRows / 2; rp += drp, qi += dqi) {
      if (qi >= nq_int) { qi -= nq_int; if (++rp >= kRows / 2) break; }
      const int col = (q_lo + qi) * 4;
      const int row = r0 + 2 * rp;
      const int64_t off = plane + (int64_t)row * W + col;
      const RowIO ioA = row_prefetch<kAdjoint>(true, n, off, fa, aa);
      const RowIO ioB = row_prefetch<kAdjoint>(true, n, off + W, fa, aa);
      const float2* rowp = V2 + rp * VW + col;
      float2 in2[4 + 2 * R];
#pragma unroll
      for (int s = 0; s < (4 + 2 * R) / 2; ++s) {
        const float4 v = *reinterpret_cast<const float4*>(rowp + 2 * s);
        in2[2 * s] = make_float2(v.x, v.y);
        in2[2 * s + 1] = make_float2(v.z, v.w);
      }
      float2 o[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float2 a = make_float2(0.f, 0.f);
#pragma unroll
        for (int k = 0; k <= 2 * R; ++k) a = __ffma2_rn(make_float2(p.wh[k], p.wh[k]), in2[j + k], a);
        o[j] = a;
      }
      row_finish<kAdjoint>(make_float4(o[0].x, o[1].x, o[2].x, o[3].x), ioA, true, n, off, nchw, fa, aa, sq, ab);
      row_finish<kAdjoint>(make_float4(o[0].y, o[1].y, o[2].y, o[3].y), ioB, true, n, off + W, nchw, fa, aa, sq, ab);
    }
  }
  constexpr int nq_b = q_lo + (W / 4 - q_hi);
  for (int i = tid; i < (kRows / 2) * nq_b; i += kT3) {
    const int rp = i / nq_b, bq = i - rp * nq_b;
    const int q = bq < q_lo ? bq : q_hi + (bq - q_lo);
    const int col = q * 4;
    const int row = r0 + 2 * rp;
    const int64_t off = plane + (int64_t)row * W + col;
    const RowIO ioA = row_prefetch<kAdjoint>(true, n, off, fa, aa);
    const RowIO ioB = row_prefetch<kAdjoint>(true, n, off + W, fa, aa);
    const float2* rowp = V2 + rp * VW + col;
    float2 o[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float* bt = border_row<R>(bhs, col + j, W);
      float2 a = make_float2(0.f, 0.f);
#pragma unroll
      for (int k = 0; k <= 2 * R; ++k) {
        const float w = bt[k];
        a = __ffma2_rn(make_float2(w, w), rowp[j + k], a);
      }
      o[j] = a;
    }
    row_finish<kAdjoint>(make_float4(o[0].x, o[1].x, o[2].x, o[3].x), ioA, true, n, off, nchw, fa, aa, sq, ab);
    row_finish<kAdjoint>(make_float4(o[0].y, o[1].y, o[2].y, o[3].y), ioB, true, n, off + W, nchw, fa, aa, sq, ab);
  }
  if (!kAdjoint && fa.partials) {
    block_sum2(sq, ab, red);
    if (tid == 0) {
      float* pp = fa.partials + ((int64_t)n * (p.C * p.strips) + blockIdx.x) * 2;
      pp[0] = sq;
      pp[1] = ab;
    }
  }
}

template <int R>
constexpr size_t sep3_smem_bytes() {
  return sizeof(float) * ((size_t)(kRows / 2) * (256 + 2 * R) * 2 + (size_t)(2 * (R + 4)) * (2 * R + 1) + 64);
}

// =================================================================================================
// scalar-FMA variant (R = 24, 32): same algorithm, one column / one row per thread item
// =================================================================================================
template <int R, bool kAdjoint>
__global__ void __launch_bounds__(kThreads) sep1_kernel(const SepParams<R> p, const FwdArgs fa, const AdjArgs aa) {
  extern __shared__ __align__(16) float smem[];
  const int H = p.H, W = p.W;
  const int SW = W + 2 * R;  // tile row stride (multiple of 4)
  constexpr int kBorder = 2 * (R + 4) * (2 * R + 1);
  constexpr int kG = 4;
  float* tile = smem;
  float* bvs = tile + (kRows + 2 * R) * SW;
  float* bhs = bvs + kBorder;
  float* red = bhs + kBorder;
  const int strip = blockIdx.x % p.strips;
  const int c = blockIdx.x / p.strips;
  const int n = blockIdx.y;
  const int r0 = strip * kRows;
  const int tid = threadIdx.x;
  const int64_t plane = (int64_t)c * H * W;
  const int64_t nchw = (int64_t)p.C * H * W;
  stage_async(bvs, p.bv, kBorder, tid, kThreads);  // border tables: asynchronous, awaited with the tile
  stage_async(bhs, p.bh, kBorder, tid, kThreads);
  sep_stage<R, kAdjoint>(tile, SW, R, r0, n, plane, p.C, H, W, fa, aa);
  for (int i = tid; i < (kRows + 2 * R) * 2 * R; i += kThreads) {
    const int tr = i / (2 * R), q = i - tr * (2 * R);
    tile[tr * SW + (q < R ? q : W + q)] = 0.f;
  }
  stage_wait();
  __syncthreads();
  for (int col = tid; col < W; col += kThreads) {  // vertical, in place (a column is touched by one thread)
    float* colp = tile + R + col;
#pragma unroll 1
    for (int g0 = 0; g0 < kRows; g0 += kG) {
      float in[kG + 2 * R];
#pragma unroll
      for (int s = 0; s < kG + 2 * R; ++s) in[s] = colp[(g0 + s) * SW];
#pragma unroll
      for (int o = 0; o < kG; ++o) {
        const int row = r0 + g0 + o;
        float acc = 0.f;
        if (row >= R + 4 && row < H - R - 4) {
#pragma unroll
          for (int k = 0; k <= 2 * R; ++k) acc = fmaf(p.wv[k], in[o + k], acc);
        } else if (row < H) {
          const float* bt = border_row<R>(bvs, row, H);
#pragma unroll
          for (int k = 0; k <= 2 * R; ++k) acc = fmaf(bt[k], in[o + k], acc);
        }
        colp[(g0 + o) * SW] = acc;
      }
    }
  }
  __syncthreads();
  float sq = 0.f, ab = 0.f;
  const int w4 = W / 4;
  for (int i = tid; i < kRows * w4; i += kThreads) {
    const int tr = i / w4, q = i - tr * w4;
    const int row = r0 + tr;
    if (row >= H) continue;
    const int col = q * 4;
    const float* rowp = tile + tr * SW + col;
    float o[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int cc = col + j;
      float acc = 0.f;
      if (cc >= R + 4 && cc < W - R - 4) {
#pragma unroll
        for (int k = 0; k <= 2 * R; ++k) acc = fmaf(p.wh[k], rowp[j + k], acc);
      } else {
        const float* bt = border_row<R>(bhs, cc, W);
#pragma unroll
        for (int k = 0; k <= 2 * R; ++k) acc = fmaf(bt[k], rowp[j + k], acc);
      }
      o[j] = acc;
    }
    sep_epilogue<kAdjoint>(make_float4(o[0], o[1], o[2], o[3]), n, plane + (int64_t)row * W + col, nchw, fa, aa, sq, ab);
  }
  if (!kAdjoint && fa.partials) {
    block_sum2(sq, ab, red);
    if (tid == 0) {
      float* pp = fa.partials + ((int64_t)n * (p.C * p.strips) + blockIdx.x) * 2;
      pp[0] = sq;
      pp[1] = ab;
    }
  }
}

template <int R>
size_t sep1_smem_bytes(int W) {
  return sizeof(float) * ((size_t)(kRows + 2 * R) * (W + 2 * R) + 2 * (size_t)(2 * (R + 4)) * (2 * R + 1) + 64);
}

size_t sep_smem_for(int R, int W) {
  switch (R) {
    case 4: return sep2_smem_bytes<4>(W);
    case 8: return sep2_smem_bytes<8>(W);
    case 12: return sep2_smem_bytes<12>(W);
    case 16: return sep2_smem_bytes<16>(W);
    case 24: return sep1_smem_bytes<24>(W);
    case 32: return sep1_smem_bytes<32>(W);
  }
  return (size_t)-1;
}

template <int R, bool kAdjoint>
int sep_launch(const dps_operator* op, const SepSet& set, const FwdArgs& fa, const AdjArgs& aa, int n, cudaStream_t st) {
  SepParams<R> p;
  for (int k = 0; k <= 2 * R; ++k) {
    p.wv[k] = set.wv[k];
    p.wh[k] = set.wh[k];
  }
  p.bv = set.bv;
  p.bh = set.bh;
  p.C = op->C; p.H = op->H; p.W = op->W;
  p.strips = (op->H + kRows - 1) / kRows;
  const size_t smem = sep_smem_for(R, op->W);
  dim3 grid((unsigned)(p.C * p.strips), (unsigned)n);
  if constexpr (R <= 16) {
    DPS_SMEM_OPTIN((sep2_kernel<R, 256, kAdjoint>), 227 * 1024, op->device);
    DPS_SMEM_OPTIN((sep2_kernel<R, 0, kAdjoint>), 227 * 1024, op->device);
    if (op->W == 256 && op->H % kRows == 0) {
      DPS_SMEM_OPTIN((sep3_kernel<R, kAdjoint>), sep3_smem_bytes<R>(), op->device);
      sep3_kernel<R, kAdjoint><<<grid, kT3, sep3_smem_bytes<R>(), st>>>(p, fa, aa);
    } else if (op->W == 256)
      sep2_kernel<R, 256, kAdjoint><<<grid, kThreads, smem, st>>>(p, fa, aa);
    else
      sep2_kernel<R, 0, kAdjoint><<<grid, kThreads, smem, st>>>(p, fa, aa);
  } else {
    DPS_SMEM_OPTIN((sep1_kernel<R, kAdjoint>), 227 * 1024, op->device);
    sep1_kernel<R, kAdjoint><<<grid, kThreads, smem, st>>>(p, fa, aa);
  }
  DPS_LAUNCH_CHECK(kAdjoint ? "sep_blur_adjoint" : "sep_blur_forward");
  return DPS_OK;
}

template <bool kAdjoint>
int sep_dispatch(const dps_operator* op, const FwdArgs& fa, const AdjArgs& aa, int n, cudaStream_t st) {
  const SepTables* t = op->sep;
  const SepSet& set = kAdjoint ? t->adj : t->fwd;
  switch (t->R) {
    case 4: return sep_launch<4, kAdjoint>(op, set, fa, aa, n, st);
    case 8: return sep_launch<8, kAdjoint>(op, set, fa, aa, n, st);
    case 12: return sep_launch<12, kAdjoint>(op, set, fa, aa, n, st);
    case 16: return sep_launch<16, kAdjoint>(op, set, fa, aa, n, st);
    case 24: return sep_launch<24, kAdjoint>(op, set, fa, aa, n, st);
    case 32: return sep_launch<32, kAdjoint>(op, set, fa, aa, n, st);
  }
  dps_set_error("separable blur: unsupported radius %d", t->R);
  return DPS_ERR_UNSUPPORTED;
}

// Dense 1-D operator matrix of "reflect-pad r then correlate with w (2r+1 taps)" on length L.
std::vector<double> band_matrix(const std::vector<double>& w, int r, int L) {
  std::vector<double> A((size_t)L * L, 0.0);
  for (int i = 0; i < L; ++i)
    for (int d = -r; d <= r; ++d) {
      int m = i + d;
      if (m < 0) m = -m;
      if (m >= L) m = 2 * (L - 1) - m;
      A[(size_t)i * L + m] += w[d + r];
    }
  return A;
}

// interior + border tap tables of A (transpose=false) or Aᵀ (transpose=true), zero-padded to radius R
int build_set(const std::vector<double>& wv, int rv, int H, const std::vector<double>& wh, int rh, int W, int R,
              bool transpose, SepSet* out) {
  auto build = [&](const std::vector<double>& w, int r, int L, std::vector<float>* interior, float** border_dev) -> int {
    std::vector<double> A = band_matrix(w, r, L);
    auto at = [&](int i, int m) -> double {  // operator entry: output i, input m
      if (m < 0 || m >= L) return 0.0;
      return transpose ? A[(size_t)m * L + i] : A[(size_t)i * L + m];
    };
    interior->assign(2 * R + 1, 0.f);  // plain taps: w[e+r] for A, the flipped w[−e+r] for Aᵀ
    for (int e = -r; e <= r; ++e) (*interior)[e + R] = (float)w[(transpose ? -e : e) + r];
    const int nb = R + 4;
    std::vector<float> border((size_t)2 * nb * (2 * R + 1), 0.f);
    for (int b = 0; b < 2 * nb; ++b) {
      const int i = b < nb ? b : (L - nb) + (b - nb);
      for (int e = -R; e <= R; ++e) border[(size_t)b * (2 * R + 1) + e + R] = (float)at(i, i + e);
    }
    DPS_CUDA(cudaMalloc(border_dev, border.size() * sizeof(float)));
    DPS_CUDA(cudaMemcpy(*border_dev, border.data(), border.size() * sizeof(float), cudaMemcpyHostToDevice));
    return DPS_OK;
  };
  if (int rc = build(wv, rv, H, &out->wv, &out->bv)) return rc;
  if (int rc = build(wh, rh, W, &out->wh, &out->bh)) return rc;
  return DPS_OK;
}

}  // namespace

// taps1d_v / taps1d_h: (2rv+1) / (2rh+1) cross-correlation taps, index d+r multiplies x[i+d]
int sep_create(dps_operator* op, const float* taps1d_v, const float* taps1d_h, int rv, int rh) {
  const int r = rv > rh ? rv : rh;
  int R = 0;
  for (int cand : {4, 8, 12, 16, 24, 32})
    if (cand >= r) { R = cand; break; }
  DPS_REQUIRE(R > 0, DPS_ERR_UNSUPPORTED, "separable blur: radius %d > 32", r);
  DPS_REQUIRE(op->H >= R + 5 && op->W >= R + 8 && r < op->H && r < op->W && op->W % 4 == 0 && op->H % 2 == 0,
              DPS_ERR_UNSUPPORTED, "separable blur: image %dx%d too small (or odd) for radius %d", op->H, op->W, R);
  const size_t smem = sep_smem_for(R, op->W);
  DPS_REQUIRE(smem <= 227 * 1024, DPS_ERR_UNSUPPORTED, "separable blur: tile of %zu bytes exceeds shared memory", smem);
  std::vector<double> wv(taps1d_v, taps1d_v + 2 * rv + 1), wh(taps1d_h, taps1d_h + 2 * rh + 1);
  SepTables* t = new SepTables();
  t->R = R;
  op->sep = t;
  if (int rc = build_set(wv, rv, op->H, wh, rh, op->W, R, false, &t->fwd)) return rc;
  if (int rc = build_set(wv, rv, op->H, wh, rh, op->W, R, true, &t->adj)) return rc;
  op->P = op->C * ((op->H + kRows - 1) / kRows);
  op->taps = 2 * r + 1;
  return sep_fused_create(op, taps1d_v, rv, taps1d_h, rh);  // fused residual + cotangent kernel where the shape allows
}

void sep_destroy(dps_operator* op) {
  if (!op->sep) return;
  for (SepSet* s : {&op->sep->fwd, &op->sep->adj}) {
    cudaFree(s->bv);
    cudaFree(s->bh);
  }
  delete op->sep;
  op->sep = nullptr;
}

int sep_forward(const dps_operator* op, const FwdArgs& a, cudaStream_t st) {
  AdjArgs dummy = {};
  return sep_dispatch<false>(op, a, dummy, a.n, st);
}
int sep_adjoint(const dps_operator* op, const AdjArgs& a, cudaStream_t st) {
  FwdArgs dummy = {};
  return sep_dispatch<true>(op, dummy, a, a.n, st);
}
