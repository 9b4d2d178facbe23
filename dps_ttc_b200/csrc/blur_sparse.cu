// Non-separable blur (motion kernels): ReflectionPad2d(k/2) + depthwise cross-correlation with an
// arbitrary (k,k) kernel (measurements.py:93-126, util/img_utils.py:268-308) and its exact adjoint.
//
// The motion kernels of the reference are thin paths: a few hundred non-zero taps of 3 721.  The plan
// keeps only the non-zero taps (dy, dx, w).
//
// forward : one CTA stages a strip of x̂₀ with its halo in shared memory, reflect-filled, and every thread
//           accumulates 16 vertically adjacent outputs of TWO columns on packed FFMA2.  The non-zero taps are
//           covered by vertical CHUNKS of 4 (same dx, dy0 … dy0+3, absent taps weigh 0): a chunk loads the 19 tile
//           value pairs under it once and feeds 64 FFMA2 (128 FMAs) from registers (a motion path is 2-5 taps thick
//           in every column, so a chunk is rarely more than half empty).  A warp owns a 64-column block and lane L the
//           columns L and L + 32 of it: every tile load is 32 consecutive words for any dx (round 2: adjacent column
//           pairs read with 64-bit loads were conflict-free only for even dx — a third of all shared-memory wavefronts
//           were conflict replays; 377 → 345 µs at N = 128).  The plan still sorts even-dx chunks first: the summation
//           order, hence every bit of the result, is the one of the earlier kernels.
// adjoint : A = C·P (P = reflect pad, C = valid correlation) ⇒ Aᵀ = Pᵀ·Cᵀ, done literally in two kernels:
//           (1) t = Cᵀu on the PADDED domain (H+2Ry, W+2Rx) — the same gather kernel with negated offsets over
//               a zero-filled tile, no border cases at all — into the operator's workspace (stays in L2);
//           (2) fold: g[m] = Σ_{p : reflect(p) = m} t[p]  (1, 2 or 4 terms per pixel), fused with
//               coef / extra / clamp mask.
// Roofline: T_nz taps → 2·T_nz flop per pixel per direction against 12-16 B: for T_nz ≳ 20 this is bound by the
// shared-memory / fp32 issue rate of the SM, NOT by HBM — stated as such in DESIGN.md.
#include <algorithm>
#include <cstdlib>
#include <vector>

#include "operator.cuh"

namespace {
constexpr int kRows = 32;   // rows per CTA of the standard variant; small grids use 16 (template parameter kR)
#ifndef DPS_SPARSE_GROUP
#define DPS_SPARSE_GROUP 16
#endif
constexpr int kGroup = DPS_SPARSE_GROUP;  // vertically adjacent outputs per thread and pass
constexpr int kMaxThreads = 320 * 16 / kGroup;  // 32-row CTAs: 5 column blocks x 2 row groups; 16-row CTAs launch half of that
constexpr int kSWFixed = 384;   // compile-time tile row strides of the 256-wide fast path: halo ≤ 64 columns per side,
constexpr int kSWFixedS = 320;  // or ≤ 32 (smaller tile, one more CTA per SM)

constexpr int kChunk = 4;  // taps per vertical chunk

struct Tap {    // global: one chunk
  int dydx;     // (dy0 << 16) | (dx & 0xffff)
  float w[kChunk];
};
struct __align__(16) TapOff {  // shared: the same with the tile offset resolved
  float w[kChunk];
  int off;      // signed offset in the staged tile of the chunk's first row
  int pad[3];
};
}  // namespace

struct SparseTables {
  int ntaps = 0;
  int n_even = 0;  // chunks [0, n_even) have an even dx
  int Ry = 0, Rx = 0;  // halo (Rx rounded up to a multiple of 4)
  Tap* taps_dev = nullptr;
};

namespace {

DPS_DEV int tap_dy(int v) { return v >> 16; }
DPS_DEV int tap_dx(int v) { return (int)(short)(v & 0xffff); }

// kAdjoint = false: out rows/cols = image;   tile = rows [r0−Ry, r0+32+Ry) × cols [−Rx, W+Rx), reflect fill
// kAdjoint = true : out rows/cols = padded t; tile = u rows [p0−2Ry, p0+32) (image coords p0−Ry …) zero fill,
//                   cols [−2Rx, W+2Rx).  blockDim.x ≥ number of output columns is NOT required (columns loop).
// kSW > 0: the tile row stride is the compile-time constant kSW (every LDS of the tap loop gets an immediate offset,
// no address arithmetic); kSW = 0: stride W + 2·halo computed at run time.
// kR: output rows per CTA — 32, or 16 when the 32-row grid would leave most SMs with one CTA (N ≲ 12 particles: a CTA's tap
// loop is a ≈40 µs dependent chain, so the launch time IS one CTA's time; twice as many half-size CTAs halve it).  The
// residual partial sums are kept per 16-ROW GROUP in both variants (one warp per (group, 64-column block), warp sums added in
// block order), so a launch's variant never changes a norm: P = C · 2 · ⌈H/32⌉ pairs per particle either way.
template <bool kAdjoint, int kSW, int kR>
__global__ void __launch_bounds__(kMaxThreads * kR / 32) sparse_kernel(const Tap* __restrict__ taps_g, int ntaps, int n_even, int Ry, int Rx, int C,
                                                     int H, int W, int strips, const FwdArgs fa, const AdjArgs aa,
                                                     float* __restrict__ t_out) {
  extern __shared__ __align__(16) float smem[];
  const int OW = kAdjoint ? W + 2 * Rx : W;      // output columns
  const int OH = kAdjoint ? H + 2 * Ry : H;      // output rows
  const int halo_x = kAdjoint ? 2 * Rx : Rx;     // tile columns left of image column 0
  const int SW = kSW > 0 ? kSW : W + 2 * halo_x;
  const int tile_rows = kR + 2 * Ry;
  float* tile = smem;
  TapOff* taps = reinterpret_cast<TapOff*>(tile + ((tile_rows * SW + 3) & ~3));
  float* red = reinterpret_cast<float*>(taps + ntaps);

  const int strip = blockIdx.x % strips;
  const int c = blockIdx.x / strips;
  const int n = blockIdx.y;
  const int o0 = strip * kR;                     // first output row of this CTA (padded coords for the adjoint)
  const int img_row0 = kAdjoint ? o0 - 2 * Ry : o0 - Ry;  // image row held by tile row 0
  const int tid = threadIdx.x, nthreads = blockDim.x;
  const int64_t plane = (int64_t)c * H * W;

  for (int i = tid; i < ntaps; i += nthreads) {
    const Tap tp = taps_g[i];
    const int dy0 = tap_dy(tp.dydx), dx = tap_dx(tp.dydx);
    // adjoint: t[p] = Σ w_i·u[p − d_i]; the run p − (dy0+3) … p − dy0 read top-down carries the weights reversed
    taps[i].off = kAdjoint ? -((dy0 + kChunk - 1) * SW + dx) : dy0 * SW + dx;
#pragma unroll
    for (int k = 0; k < kChunk; ++k) taps[i].w[k] = tp.w[kAdjoint ? kChunk - 1 - k : k];
  }
  {
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
    const int w4 = W / 4;
    // kSB independent 16-byte requests (twice that with ε) per thread in flight before the first shared store: at N = 8
    // (one or two CTAs per SM) every batch is a full DRAM / L2 round trip on the launch's critical path
    // — the small-grid variant (kR = 16) therefore keeps 7 (14) in flight, the full-machine variant 4 (8) and 64 registers.
    constexpr int kSB = kR == 16 ? 7 : 4;
    for (int i0 = tid; i0 < tile_rows * w4; i0 += kSB * nthreads) {
      float4 v[kSB];
#pragma unroll
      for (int b = 0; b < kSB; ++b) {
        const int i = i0 + b * nthreads;
        const int tr = i / w4, q = i - tr * w4;
        int row = img_row0 + tr;
        v[b] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (i < tile_rows * w4) {
          if (kAdjoint) {
            if (row >= 0 && row < H) v[b] = ldg_stream4(x + (int64_t)row * W + q * 4);
          } else {
            row = reflect_idx(row, H);
            v[b] = src_load4(x, eps, (int64_t)row * W + q * 4, c1, c2, clip);
          }
        }
      }
#pragma unroll
      for (int b = 0; b < kSB; ++b) {
        const int i = i0 + b * nthreads;
        const int tr = i / w4, q = i - tr * w4;
        if (i < tile_rows * w4) *reinterpret_cast<float4*>(tile + tr * SW + halo_x + q * 4) = v[b];
      }
    }
    __syncthreads();
    // column halos: reflect (forward) from the staged interior, zero (adjoint)
    for (int i = tid; i < tile_rows * 2 * halo_x; i += nthreads) {
      const int tr = i / (2 * halo_x), q = i - tr * (2 * halo_x);
      const int colimg = q < halo_x ? q - halo_x : W + (q - halo_x);
      float v = 0.f;
      if (!kAdjoint) v = tile[tr * SW + halo_x + reflect_idx(colimg, W)];
      tile[tr * SW + halo_x + colimg] = v;
    }
  }
  __syncthreads();

  // items = (64-column block, group of kGroup rows), one per WARP: lane L owns columns 64·blk + L and 64·blk + 32 + L.
  // Every tile load of the tap loop is then a 32-bit access of 32 CONSECUTIVE words whatever the tap's dx: one wavefront,
  // no bank conflict.  (Adjacent column pairs — one 64-bit load per pair — are conflict-free only for even dx; for odd dx
  // the two 32-bit loads of a pair have stride 2 and cost two wavefronts each: ncu counted 28 M excessive wavefronts of
  // 85 M at N = 128, profiles/r3c_ncu_motion_fwd_n128.csv, and shared-memory bandwidth is what bounds this kernel.)
  const int nblk = (OW + 63) >> 6;
  constexpr int kGroups = kR / kGroup;
  const int lane = tid & 31;
  for (int item = tid >> 5; item < nblk * kGroups; item += nthreads >> 5) {
    const int grp = item / nblk, blk = item - grp * nblk;
    const int g0 = grp * kGroup;
    float sq = 0.f, ab = 0.f;
    const int col_lo = blk * 64 + lane;
    const bool hi_ok = col_lo + 32 < OW, lo_ok = col_lo < OW;
    const int col_hi = hi_ok ? col_lo + 32 : col_lo;  // (a block that sticks out of the row recomputes its low half)
    // tile element under output (o0+g0, col) with zero tap offset:
    //   forward: image (o0+g0, col)            → tile row g0+Ry,      tile col halo_x+col
    //   adjoint: padded (p, q) = image (p−Ry, q−Rx) → tile row g0+Ry, tile col halo_x+col−Rx
    const float* base = tile + (Ry + g0) * SW + halo_x + (lo_ok ? col_lo : 0) - (kAdjoint ? Rx : 0);
    const int dhi = col_hi - (lo_ok ? col_lo : 0);
    float2 acc[kGroup];
#pragma unroll
    for (int j = 0; j < kGroup; ++j) acc[j] = make_float2(0.f, 0.f);
#pragma unroll 2
    for (int t = 0; t < ntaps; ++t) {
      const float4 w = *reinterpret_cast<const float4*>(taps[t].w);
      const float* p = base + taps[t].off;
      const float* ph = p + dhi;
      float2 v[kGroup + kChunk - 1];
#pragma unroll
      for (int i = 0; i < kGroup + kChunk - 1; ++i) v[i] = make_float2(p[i * SW], ph[i * SW]);
#pragma unroll
      for (int j = 0; j < kGroup; ++j) {
        acc[j] = __ffma2_rn(make_float2(w.x, w.x), v[j], acc[j]);
        acc[j] = __ffma2_rn(make_float2(w.y, w.y), v[j + 1], acc[j]);
        acc[j] = __ffma2_rn(make_float2(w.z, w.z), v[j + 2], acc[j]);
        acc[j] = __ffma2_rn(make_float2(w.w, w.w), v[j + 3], acc[j]);
      }
    }
    constexpr bool kYBatch = !kAdjoint && kR == 16;  // small grids: the epilogue's y loads are a serial chain of L2 round trips
    float2 yv[kYBatch ? kGroup : 1];
    if (kYBatch && fa.y) {  // all measurement values of the group in flight at once (one round trip, not kGroup)
#pragma unroll
      for (int j = 0; j < kGroup; ++j) {
        const int row = o0 + g0 + j;
        const float* yr = fa.y + n * fa.y_stride + plane + (int64_t)row * W;
        yv[j] = make_float2((row < OH && lo_ok) ? __ldg(yr + col_lo) : 0.f, (row < OH && hi_ok) ? __ldg(yr + col_hi) : 0.f);
      }
    }
#pragma unroll
    for (int j = 0; j < kGroup; ++j) {
      const int row = o0 + g0 + j;
      if (row >= OH) continue;
      if (!kAdjoint) {
        const int64_t off = plane + (int64_t)row * W;
        float2 res = acc[j];
        if (fa.y) {
          if (kYBatch) {
            res = make_float2(__fsub_rn(yv[j].x, res.x), __fsub_rn(yv[j].y, res.y));
          } else {
            const float* yr = fa.y + n * fa.y_stride + off;
            res = make_float2(__fsub_rn(lo_ok ? __ldg(yr + col_lo) : 0.f, res.x), __fsub_rn(hi_ok ? __ldg(yr + col_hi) : 0.f, res.y));
          }
        }
        float* orow = fa.out + (int64_t)n * C * H * W + off;
        if (lo_ok) { stg_stream(orow + col_lo, res.x); sq = fmaf(res.x, res.x, sq); ab += fabsf(res.x); }
        if (hi_ok) { stg_stream(orow + col_hi, res.y); sq = fmaf(res.y, res.y, sq); ab += fabsf(res.y); }
      } else {
        float* trow = t_out + (((int64_t)n * C + c) * OH + row) * OW;
        if (lo_ok) stg_stream(trow + col_lo, acc[j].x);
        if (hi_ok) stg_stream(trow + col_hi, acc[j].y);
      }
    }
    if (!kAdjoint && fa.partials) {  // this item's sums → red[item]; added per row group, in block order, below
      sq = warp_sum(sq);
      ab = warp_sum(ab);
      if (lane == 0) {
        red[item] = sq;
        red[32 + item] = ab;
      }
    }
  }
  if (!kAdjoint && fa.partials) {
    __syncthreads();
    if (tid < kGroups) {
      float sq = 0.f, ab = 0.f;
      for (int blk = 0; blk < nblk; ++blk) {
        sq += red[tid * nblk + blk];
        ab += red[32 + tid * nblk + blk];
      }
      // slot of the 16-row group (o0 + 16·tid)/16 of channel c: the same for the 32-row and the 16-row variant
      const int groups_per_plane = strips * kGroups;
      float* pp = fa.partials + ((int64_t)n * (C * groups_per_plane) + c * groups_per_plane + strip * kGroups + tid) * 2;
      pp[0] = sq;
      pp[1] = ab;
    }
  }
}

// g[my][mx] = Σ_{rows p : reflect(p−Ry)=my} Σ_{cols q : reflect(q−Rx)=mx} t[p][q], then the cotangent epilogue.
// A thread owns four adjacent pixels: the direct terms, the clamp-mask sources and the store are 128-bit accesses (five to
// seven independent 16-byte requests per thread in flight; the scalar one-pixel-per-thread version reached a third of the HBM
// rate: 185 µs of the 605 µs adjoint at N = 128); the mirrored column terms of border pixels stay scalar.  Every pixel sums
// its terms in the order of the scalar version (rows: direct, top, bottom; columns: direct, left, right): same bits.
__global__ void __launch_bounds__(256) sparse_fold_kernel(const float* __restrict__ t, int Ry, int Rx, int C, int H,
                                                          int W, const AdjArgs aa) {
  const int n = blockIdx.z, c = blockIdx.y;
  const int OH = H + 2 * Ry, OW = W + 2 * Rx;
  const float* tp = t + ((int64_t)n * C + c) * OH * OW;
  const int idx4 = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx4 * 4 >= H * W) return;
  const int my = (idx4 * 4) / W, mx0 = idx4 * 4 - my * W;  // W % 4 == 0: the four pixels share the row
  int rows[3], nr = 0;
  rows[nr++] = my + Ry;
  if (my >= 1 && my <= Ry) rows[nr++] = Ry - my;
  if (my <= H - 2 && my >= H - 1 - Ry) rows[nr++] = 2 * (H - 1) - my + Ry;
  const bool border_cols = mx0 <= Rx || mx0 + 3 >= W - 1 - Rx;
  float s[4] = {0.f, 0.f, 0.f, 0.f};
  for (int a = 0; a < nr; ++a) {
    const float* tr = tp + (int64_t)rows[a] * OW;
    const float4 d = ldg_stream4(tr + mx0 + Rx);  // Rx % 4 == 0, OW % 4 == 0: aligned
    const float dv[4] = {d.x, d.y, d.z, d.w};
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      s[e] += dv[e];
      if (border_cols) {
        const int mx = mx0 + e;
        if (mx >= 1 && mx <= Rx) s[e] += tr[Rx - mx];
        if (mx <= W - 2 && mx >= W - 1 - Rx) s[e] += tr[2 * (W - 1) - mx + Rx];
      }
    }
  }
  const int64_t off = (int64_t)c * H * W + (int64_t)idx4 * 4;
  const float cf = aa.coef ? aa.coef[n] : 1.0f;
  float4 res = make_float4(cf * s[0], cf * s[1], cf * s[2], cf * s[3]);
  if (aa.extra) {
    const float4 ex = ldg_stream4(aa.extra + n * aa.extra_stride + off);
    res.x += ex.x; res.y += ex.y; res.z += ex.z; res.w += ex.w;
  }
  const float4 m = mask_load4(aa.mask_src, aa.has_mask, n, off);
  res.x *= m.x; res.y *= m.y; res.z *= m.z; res.w *= m.w;
  stg_stream4(aa.g + n * aa.g_stride + off, res);
}

// tile row stride: the compile-time kSWFixed when the image is 256 wide and the halo fits, else W + 2·halo
int sparse_stride(const dps_operator* op, bool adjoint, bool* fixed) {
  const SparseTables* t = op->sparse;
  const int halo_x = adjoint ? 2 * t->Rx : t->Rx;
  *fixed = op->W == 256 && op->W + 2 * halo_x <= kSWFixed;
  return !*fixed ? op->W + 2 * halo_x : (op->W + 2 * halo_x <= kSWFixedS ? kSWFixedS : kSWFixed);
}

size_t sparse_smem(const dps_operator* op, bool adjoint, int rows = kRows) {
  const SparseTables* t = op->sparse;
  bool fixed;
  const int SW = sparse_stride(op, adjoint, &fixed);
  size_t bytes = sizeof(float) * ((((size_t)(rows + 2 * t->Ry) * SW + 3) & ~(size_t)3) + 64) + sizeof(TapOff) * (size_t)t->ntaps;
  if (fixed && bytes > 227 * 1024) {  // tall kernels: fall back to the tight run-time stride
    bytes = sizeof(float) * ((((size_t)(rows + 2 * t->Ry) * (op->W + 2 * (adjoint ? 2 * t->Rx : t->Rx)) + 3) & ~(size_t)3) + 64) +
            sizeof(TapOff) * (size_t)t->ntaps;
  }
  return bytes;
}
bool sparse_fixed(const dps_operator* op, bool adjoint) {
  const SparseTables* t = op->sparse;
  bool fixed;
  const int SW = sparse_stride(op, adjoint, &fixed);
  const size_t bytes = sizeof(float) * ((((size_t)(kRows + 2 * t->Ry) * SW + 3) & ~(size_t)3) + 64) + sizeof(TapOff) * (size_t)t->ntaps;
  return fixed && bytes <= 227 * 1024;
}

}  // namespace

namespace {
template <bool kA, int kSW, int kR>
int sparse_optin_one() {
  DPS_CUDA(cudaFuncSetAttribute(sparse_kernel<kA, kSW, kR>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
  return DPS_OK;
}
int sparse_optin_all() {  // (per current device: called from every plan creation)
  int rc = 0;
  rc |= sparse_optin_one<false, 0, kRows>();         rc |= sparse_optin_one<true, 0, kRows>();
  rc |= sparse_optin_one<false, kSWFixed, kRows>();  rc |= sparse_optin_one<true, kSWFixed, kRows>();
  rc |= sparse_optin_one<false, kSWFixedS, kRows>(); rc |= sparse_optin_one<true, kSWFixedS, kRows>();
  rc |= sparse_optin_one<false, 0, 16>();            rc |= sparse_optin_one<true, 0, 16>();
  rc |= sparse_optin_one<false, kSWFixed, 16>();     rc |= sparse_optin_one<true, kSWFixed, 16>();
  rc |= sparse_optin_one<false, kSWFixedS, 16>();    rc |= sparse_optin_one<true, kSWFixedS, 16>();
  return rc ? DPS_ERR_CUDA : DPS_OK;
}
}  // namespace

int sparse_create(dps_operator* op, const float* kernel, int ksize) {
  const int r0 = ksize / 2;
  int Ry = 0, Rx = 0, nnz = 0;
  for (int a = 0; a < ksize; ++a)
    for (int b = 0; b < ksize; ++b) {
      if (kernel[a * ksize + b] == 0.0f) continue;
      ++nnz;
      Ry = std::max(Ry, abs(a - r0));
      Rx = std::max(Rx, abs(b - r0));
    }
  Ry = std::max(Ry, 2);  // a chunk of 4 rows must fit inside the halo window [−Ry, Ry]
  // cover the non-zero taps of every column with chunks of kChunk rows; a chunk never leaves [−Ry, Ry]
  std::vector<Tap> taps;
  for (int b = 0; b < ksize; ++b) {
    int a = 0;
    while (a < ksize) {
      if (kernel[a * ksize + b] == 0.0f) { ++a; continue; }
      const int dy_first = a - r0;
      const int dy0 = std::min(dy_first, Ry - (kChunk - 1));
      Tap tp;
      tp.dydx = (int)(((unsigned)dy0 << 16) | ((unsigned)(b - r0) & 0xffffu));
      for (int k = 0; k < kChunk; ++k) {
        const int aa = dy0 + k + r0;  // rows before dy_first belong to an earlier chunk (or are zero)
        tp.w[k] = (aa >= a && aa < ksize) ? kernel[aa * ksize + b] : 0.0f;
      }
      taps.push_back(tp);
      a = dy0 + kChunk + r0;
    }
  }
  DPS_REQUIRE(!taps.empty(), DPS_ERR_INVALID, "blur: kernel is all zero");
  // even-dx chunks first (aligned 64-bit pair loads), odd-dx chunks after them
  std::stable_partition(taps.begin(), taps.end(), [](const Tap& tp) { return ((tp.dydx & 0xffff) & 1) == 0; });
  int n_even = 0;
  for (const Tap& tp : taps) n_even += ((tp.dydx & 0xffff) & 1) == 0;
  DPS_REQUIRE(Ry < op->H && Rx < op->W, DPS_ERR_UNSUPPORTED, "sparse blur: kernel radius (%d,%d) reaches the image size", Ry, Rx);
  Rx = (Rx + 3) / 4 * 4;
  if (Rx == 0) Rx = 4;
  DPS_REQUIRE(op->W % 4 == 0, DPS_ERR_UNSUPPORTED, "sparse blur: W must be a multiple of 4");
  SparseTables* t = new SparseTables();
  t->ntaps = (int)taps.size();
  t->n_even = n_even;
  t->Ry = Ry;
  t->Rx = Rx;
  op->sparse = t;
  DPS_REQUIRE(sparse_smem(op, true) <= 227 * 1024, DPS_ERR_UNSUPPORTED, "sparse blur: tile exceeds shared memory");
  DPS_CUDA(cudaMalloc(&t->taps_dev, taps.size() * sizeof(Tap)));
  DPS_CUDA(cudaMemcpy(t->taps_dev, taps.data(), taps.size() * sizeof(Tap), cudaMemcpyHostToDevice));
  if (int rc = sparse_optin_all()) return rc;
  op->P = op->C * 2 * ((op->H + kRows - 1) / kRows);  // one pair per 16-row group
  op->taps = nnz;
  op->aux_floats = (int64_t)op->C * (op->H + 2 * Ry) * (op->W + 2 * Rx);  // padded t of the adjoint
  return DPS_OK;
}

void sparse_destroy(dps_operator* op) {
  if (!op->sparse) return;
  cudaFree(op->sparse->taps_dev);
  delete op->sparse;
  op->sparse = nullptr;
}

int sparse_sm_count(int device) {
  static int sms[64] = {};
  if (device < 0 || device >= 64) return 148;
  if (!sms[device]) cudaDeviceGetAttribute(&sms[device], cudaDevAttrMultiProcessorCount, device);
  return sms[device] > 0 ? sms[device] : 148;
}

// rows per CTA: 16 while the 32-row grid would leave the machine under two CTAs per SM
int sparse_rows(const dps_operator* op, int strips32, int n) {
  return (int64_t)op->C * strips32 * n < 2 * (int64_t)sparse_sm_count(op->device) ? 16 : kRows;
}

template <bool kAdjoint, int kSW, int kR>
void sparse_launch(const dps_operator* op, int out_h, int out_w, int n, const FwdArgs& fa, const AdjArgs& aa, float* t_out, cudaStream_t st) {
  const SparseTables* t = op->sparse;
  const int strips = ((out_h + kRows - 1) / kRows) * (kRows / kR);  // the 16-row grid covers the same 32-row strips (slots of the partial sums)
  const int threads = std::max(kAdjoint ? 32 : 128, std::min(kMaxThreads * kR / 32, 32 * ((out_w + 63) / 64) * (kR / kGroup)));
  dim3 grid((unsigned)(op->C * strips), (unsigned)n);
  sparse_kernel<kAdjoint, kSW, kR><<<grid, threads, sparse_smem(op, kAdjoint, kR), st>>>(t->taps_dev, t->ntaps, t->n_even, t->Ry, t->Rx, op->C, op->H,
                                                                                      op->W, strips, fa, aa, t_out);
}
template <bool kAdjoint>
void sparse_dispatch(const dps_operator* op, int out_h, int out_w, int n, const FwdArgs& fa, const AdjArgs& aa, float* t_out, cudaStream_t st) {
  bool fx;
  const int SW = sparse_stride(op, kAdjoint, &fx);
  const int rows = sparse_rows(op, (out_h + kRows - 1) / kRows, n);
  const int sel = sparse_fixed(op, kAdjoint) ? (SW == kSWFixedS ? 1 : 2) : 0;
#define DPS_SPARSE_CASE(S, SWV)                                                                          \
  if (sel == S) {                                                                                        \
    if (rows == 16) sparse_launch<kAdjoint, SWV, 16>(op, out_h, out_w, n, fa, aa, t_out, st);            \
    else sparse_launch<kAdjoint, SWV, kRows>(op, out_h, out_w, n, fa, aa, t_out, st);                    \
  }
  DPS_SPARSE_CASE(1, kSWFixedS)
  DPS_SPARSE_CASE(2, kSWFixed)
  DPS_SPARSE_CASE(0, 0)
#undef DPS_SPARSE_CASE
}

int sparse_forward(const dps_operator* op, const FwdArgs& a, cudaStream_t st) {
  AdjArgs dummy = {};
  sparse_dispatch<false>(op, op->H, op->W, a.n, a, dummy, nullptr, st);
  DPS_LAUNCH_CHECK("sparse_blur_forward");
  return DPS_OK;
}

int sparse_adjoint(const dps_operator* op, const AdjArgs& a, cudaStream_t st) {
  const SparseTables* t = op->sparse;
  DPS_REQUIRE(a.aux, DPS_ERR_INVALID, "sparse blur adjoint needs the aux workspace (%lld floats per particle)",
              (long long)op->aux_floats);
  float* scratch = const_cast<float*>(a.aux);
  FwdArgs dummy = {};
  const int OH = op->H + 2 * t->Ry, OW = op->W + 2 * t->Rx;
  sparse_dispatch<true>(op, OH, OW, a.n, dummy, a, scratch, st);
  DPS_LAUNCH_CHECK("sparse_blur_adjoint_t");
  dim3 fgrid((unsigned)((op->H * op->W / 4 + 255) / 256), (unsigned)op->C, (unsigned)a.n);
  sparse_fold_kernel<<<fgrid, 256, 0, st>>>(scratch, t->Ry, t->Rx, op->C, op->H, op->W, a);
  DPS_LAUNCH_CHECK("sparse_blur_adjoint_fold");
  return DPS_OK;
}
