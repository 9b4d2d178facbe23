// Super-resolution operator: the Shocher Resizer (util/resizer.py:55-74) used by
// SuperResolutionOperator.forward (measurements.py:84-85) and its exact adjoint.
//
// The Resizer's per-dimension "x[fov] * w summed over taps" is the banded matrix
//   A[j][m] = Σ_k w[k][j]·[fov[k][j] == m]      (out_len × in_len; reflection folds duplicate indices)
// so  y = A_h · X · A_wᵀ  per channel and  Aᵀ: G ↦ A_hᵀ · G · A_w  (SURVEY.md App. A.4).  The plan turns
// the reference's own (fov, weights) tables into strip-local dense blocks of A_h and row/column
// windows of A_w; nothing 16×-sized is ever materialised (the reference's gather builds a
// (taps, out, N, C, W) intermediate).
//
// forward : CTA = (particle, channel, 8 output rows).  H pass straight from global memory — one
//           thread per column, every input row of the strip's window read once, coalesced — into a
//           (8, W) shared tile, then the W pass from shared memory, then residual + Σr², Σ|r|.
//           The reference resizes W first and H second; the order only changes fp32 rounding.
// adjoint : CTA = (particle, channel, 32 input rows).  The (out_h, out_w) residual plane (16 KB at ×4)
//           is staged whole; E = G·A_w for the ≤16 measurement rows the strip touches lives in
//           registers (one thread per input column), then out = A_hᵀ·E, fused with coef/extra/clamp mask.
// Roofline: HBM-bound.  forward 2T + M, adjoint M + 3T per particle (T = particle, M = measurement).
#include <algorithm>
#include <vector>

#include "operator.cuh"

namespace {
constexpr int kThreads = 256;
constexpr int kRO = 8;     // forward: output rows per CTA
constexpr int kRA = 32;    // adjoint: input-space rows per CTA
constexpr int kJMax = 16;  // adjoint: measurement rows a strip may touch
constexpr int kKTMax = 12; // adjoint: measurement columns touching one input column
}  // namespace

struct ResizeTables {
  // forward
  int fstrips = 0;        // ceil(out_h / kRO)
  int span = 0;           // max input-row window of a strip
  int* f_rmin = nullptr;  // (fstrips)
  int* f_rcnt = nullptr;  // (fstrips)
  float* f_dh = nullptr;  // (fstrips, span, kRO): A_h[strip*kRO + j][rmin + rr]
  int kw = 0;             // max column window of an output column
  int* f_cstart = nullptr;  // (out_w)
  float* f_ww = nullptr;    // (out_w, kw): A_w[j][cstart[j] + k]
  // adjoint
  int astrips = 0;        // ceil(H / kRA)
  int* a_jmin = nullptr;  // (astrips)
  int* a_jcnt = nullptr;  // (astrips)
  float* a_dht = nullptr; // (astrips, kRA, kJMax): A_h[jmin + jj][strip*kRA + i]
  int kt = 0;             // max measurement-column window of an input column
  int* a_jstart = nullptr;  // (W)
  float* a_wt = nullptr;    // (W, kt): A_w[jstart[m] + k][m]
};

namespace {

__global__ void __launch_bounds__(kThreads) resize_fwd_kernel(const ResizeTables t, int C, int H, int W,
                                                              int oH, int oW, const FwdArgs a) {
  extern __shared__ __align__(16) float smem[];
  float* dh = smem;                  // (span, kRO)
  float* V = dh + t.span * kRO;      // (kRO, W)
  float* red = V + kRO * W;          // 64
  float* wws = red + 64;             // (kw, oW): W-pass weights, tap-major so that a warp reads consecutive words
  int* css = reinterpret_cast<int*>(wws + t.kw * oW);  // (oW)
  const int strip = blockIdx.x % t.fstrips;
  const int c = blockIdx.x / t.fstrips;
  const int n = blockIdx.y;
  const int tid = threadIdx.x;
  const int rmin = t.f_rmin[strip], rcnt = t.f_rcnt[strip];
  for (int i = tid; i < rcnt * kRO; i += kThreads) dh[i] = t.f_dh[(int64_t)strip * t.span * kRO + i];
  for (int i = tid; i < t.kw * oW; i += kThreads) {  // coalesced read of (oW, kw), transposed store
    const int jc = i / t.kw, k = i - jc * t.kw;
    wws[k * oW + jc] = t.f_ww[i];
  }
  for (int i = tid; i < oW; i += kThreads) css[i] = t.f_cstart[i];
  __syncthreads();

  const int64_t plane = (int64_t)c * H * W;
  const float* x = a.src.x + n * a.src.x_stride + plane;
  const float* eps = a.src.eps ? a.src.eps + n * a.src.eps_stride + plane : nullptr;
  // ---- H pass: V[j][col] = Σ_rows A_h[j0+j][row]·x̂₀[row][col] -----------------------------------
  for (int col = tid; col < W; col += kThreads) {
    float acc[kRO];
#pragma unroll
    for (int j = 0; j < kRO; ++j) acc[j] = 0.f;
    // rows in batches of kBatch: all loads of a batch are issued before the first use (2·kBatch requests in
    // flight per thread), which is what hides the HBM latency in this otherwise serial walk down the column
    constexpr int kBatch = 12;
#pragma unroll 1
    for (int rr0 = 0; rr0 < rcnt; rr0 += kBatch) {
      float xv[kBatch], ev[kBatch];
#pragma unroll
      for (int b = 0; b < kBatch; ++b) {
        const int rr = rr0 + b < rcnt ? rr0 + b : rcnt - 1;
        xv[b] = ldg_stream(x + (int64_t)(rmin + rr) * W + col);
        ev[b] = eps ? ldg_stream(eps + (int64_t)(rmin + rr) * W + col) : 0.f;
      }
#pragma unroll
      for (int b = 0; b < kBatch; ++b) {
        if (rr0 + b < rcnt) {
          const float v = eps ? x0_of(xv[b], ev[b], a.src.c1, a.src.c2, a.src.clip) : xv[b];
          const float4 w0 = *reinterpret_cast<const float4*>(dh + (rr0 + b) * kRO);
          const float4 w1 = *reinterpret_cast<const float4*>(dh + (rr0 + b) * kRO + 4);
          acc[0] = fmaf(w0.x, v, acc[0]); acc[1] = fmaf(w0.y, v, acc[1]);
          acc[2] = fmaf(w0.z, v, acc[2]); acc[3] = fmaf(w0.w, v, acc[3]);
          acc[4] = fmaf(w1.x, v, acc[4]); acc[5] = fmaf(w1.y, v, acc[5]);
          acc[6] = fmaf(w1.z, v, acc[6]); acc[7] = fmaf(w1.w, v, acc[7]);
        }
      }
    }
#pragma unroll
    for (int j = 0; j < kRO; ++j) V[j * W + col] = acc[j];
  }
  __syncthreads();
  // ---- W pass + epilogue -----------------------------------------------------------------------
  float sq = 0.f, ab = 0.f;
  const int64_t oplane = ((int64_t)n * C + c) * oH * oW;
  const int64_t yplane = (int64_t)c * oH * oW;
  for (int i = tid; i < kRO * oW; i += kThreads) {
    const int j = i / oW, jc = i - j * oW;
    const int orow = strip * kRO + j;
    if (orow >= oH) continue;
    const int cs = css[jc];
    const float* vr = V + j * W + cs;
    float acc = 0.f;
    for (int k = 0; k < t.kw; ++k) acc = fmaf(wws[k * oW + jc], (cs + k < W) ? vr[k] : 0.f, acc);
    float res = acc;
    const int64_t o = (int64_t)orow * oW + jc;
    if (a.y) res = __fsub_rn(ldg_ro(a.y + n * a.y_stride + yplane + o), res);
    stg_stream(a.out + oplane + o, res);
    sq += res * res;
    ab += fabsf(res);
  }
  if (a.partials) {
    block_sum2(sq, ab, red);
    if (tid == 0) {
      float* pp = a.partials + ((int64_t)n * (C * t.fstrips) + blockIdx.x) * 2;
      pp[0] = sq;
      pp[1] = ab;
    }
  }
}

// Pair variant for W = 256: 256 threads = 128 column pairs × 2 row halves.  64-bit loads, x̂₀ and the accumulation on
// packed FFMA2 — half the load and FMA instructions of the scalar kernel at the same occupancy.  The two row-half
// partial sums are combined through shared memory in a fixed order.
__global__ void __launch_bounds__(kThreads) resize_fwd_pair_kernel(const ResizeTables t, int C, int H, int oH, int oW,
                                                                   const FwdArgs a) {
  constexpr int W = 256, W2 = 128, kBatch = 12;
  extern __shared__ __align__(16) float smem[];
  float* dh = smem;                    // (span, kRO)
  float* V = dh + t.span * kRO;        // (kRO, W)   half 0 partial, then the sum
  float* V1 = V + kRO * W;             // (kRO, W)   half 1 partial
  float* red = V1 + kRO * W;           // 64
  float* wws = red + 64;               // (kw, oW)
  int* css = reinterpret_cast<int*>(wws + t.kw * oW);
  const int strip = blockIdx.x % t.fstrips;
  const int c = blockIdx.x / t.fstrips;
  const int n = blockIdx.y;
  const int tid = threadIdx.x;
  const int rmin = t.f_rmin[strip], rcnt = t.f_rcnt[strip];
  for (int i = tid; i < rcnt * kRO; i += kThreads) dh[i] = t.f_dh[(int64_t)strip * t.span * kRO + i];
  for (int i = tid; i < t.kw * oW; i += kThreads) {
    const int jc = i / t.kw, k = i - jc * t.kw;
    wws[k * oW + jc] = t.f_ww[i];
  }
  for (int i = tid; i < oW; i += kThreads) css[i] = t.f_cstart[i];
  __syncthreads();
  const int half = tid >> 7, cp = tid & 127;
  const int hrows = (rcnt + 1) >> 1;             // rows per half
  const int rr_lo = half * hrows, rr_hi = min(rcnt, rr_lo + hrows);
  const int64_t plane = (int64_t)c * H * W;
  const float2* x2 = reinterpret_cast<const float2*>(a.src.x + n * a.src.x_stride + plane);
  const float2* e2 = a.src.eps ? reinterpret_cast<const float2*>(a.src.eps + n * a.src.eps_stride + plane) : nullptr;
  float2 acc[kRO];
#pragma unroll
  for (int j = 0; j < kRO; ++j) acc[j] = make_float2(0.f, 0.f);
#pragma unroll 1
  for (int rr0 = rr_lo; rr0 < rr_hi; rr0 += kBatch) {
    float2 xv[kBatch], ev[kBatch];
#pragma unroll
    for (int b = 0; b < kBatch; ++b) {
      const int rr = rr0 + b < rr_hi ? rr0 + b : rr_hi - 1;
      xv[b] = ldg_stream2(x2 + (rmin + rr) * W2 + cp);
      ev[b] = e2 ? ldg_stream2(e2 + (rmin + rr) * W2 + cp) : make_float2(0.f, 0.f);
    }
#pragma unroll
    for (int b = 0; b < kBatch; ++b) {
      if (rr0 + b < rr_hi) {
        const float2 v = e2 ? x0_pair(xv[b], ev[b], a.src.c1, a.src.c2, a.src.clip) : xv[b];
        const float4 w0 = *reinterpret_cast<const float4*>(dh + (rr0 + b) * kRO);
        const float4 w1 = *reinterpret_cast<const float4*>(dh + (rr0 + b) * kRO + 4);
        acc[0] = __ffma2_rn(make_float2(w0.x, w0.x), v, acc[0]); acc[1] = __ffma2_rn(make_float2(w0.y, w0.y), v, acc[1]);
        acc[2] = __ffma2_rn(make_float2(w0.z, w0.z), v, acc[2]); acc[3] = __ffma2_rn(make_float2(w0.w, w0.w), v, acc[3]);
        acc[4] = __ffma2_rn(make_float2(w1.x, w1.x), v, acc[4]); acc[5] = __ffma2_rn(make_float2(w1.y, w1.y), v, acc[5]);
        acc[6] = __ffma2_rn(make_float2(w1.z, w1.z), v, acc[6]); acc[7] = __ffma2_rn(make_float2(w1.w, w1.w), v, acc[7]);
      }
    }
  }
  {
    float* dstp = half ? V1 : V;
#pragma unroll
    for (int j = 0; j < kRO; ++j) *reinterpret_cast<float2*>(dstp + j * W + 2 * cp) = acc[j];
  }
  __syncthreads();
  for (int i = tid; i < kRO * W / 4; i += kThreads) {  // V += V1 (fixed order)
    float4 s0 = *reinterpret_cast<const float4*>(V + i * 4);
    const float4 s1 = *reinterpret_cast<const float4*>(V1 + i * 4);
    s0.x += s1.x; s0.y += s1.y; s0.z += s1.z; s0.w += s1.w;
    *reinterpret_cast<float4*>(V + i * 4) = s0;
  }
  __syncthreads();
  float sq = 0.f, ab = 0.f;
  const int64_t oplane = ((int64_t)n * C + c) * oH * oW;
  const int64_t yplane = (int64_t)c * oH * oW;
  for (int i = tid; i < kRO * oW; i += kThreads) {
    const int j = i / oW, jc = i - j * oW;
    const int orow = strip * kRO + j;
    if (orow >= oH) continue;
    const int64_t o = (int64_t)orow * oW + jc;
    const float yv = a.y ? ldg_ro(a.y + n * a.y_stride + yplane + o) : 0.f;
    const int cs = css[jc];
    const float* vr = V + j * W + cs;
    float s = 0.f;
    for (int k = 0; k < t.kw; ++k) s = fmaf(wws[k * oW + jc], (cs + k < W) ? vr[k] : 0.f, s);
    const float res = a.y ? __fsub_rn(yv, s) : s;
    stg_stream(a.out + oplane + o, res);
    sq += res * res;
    ab += fabsf(res);
  }
  if (a.partials) {
    block_sum2(sq, ab, red);
    if (tid == 0) {
      float* pp = a.partials + ((int64_t)n * (C * t.fstrips) + blockIdx.x) * 2;
      pp[0] = sq;
      pp[1] = ab;
    }
  }
}

__global__ void __launch_bounds__(kThreads) resize_adj_kernel(const ResizeTables t, int C, int H, int W,
                                                              int oH, int oW, const AdjArgs a) {
  extern __shared__ __align__(16) float smem[];
  float* G = smem;               // (oH, oW) whole residual plane of this (n, c)
  float* dht = G + ((oH * oW + 3) & ~3);  // (kRA, kJMax), 16-byte aligned
  const int strip = blockIdx.x % t.astrips;
  const int c = blockIdx.x / t.astrips;
  const int n = blockIdx.y;
  const int tid = threadIdx.x;
  const int jmin = t.a_jmin[strip], jcnt = t.a_jcnt[strip];
  const float* r = a.r + ((int64_t)n * C + c) * oH * oW;
  for (int i = tid; i < oH * oW; i += kThreads) G[i] = ldg_stream(r + i);
  for (int i = tid; i < kRA * kJMax; i += kThreads) dht[i] = t.a_dht[(int64_t)strip * kRA * kJMax + i];
  float* wts = dht + kRA * kJMax;  // (kt, W): column weights, tap-major (coalesced global read, conflict-free use)
  for (int i = tid; i < t.kt * W; i += kThreads) {
    const int m = i / t.kt, k = i - m * t.kt;
    wts[k * W + m] = t.a_wt[i];
  }
  __syncthreads();

  const float coef = a.coef ? a.coef[n] : 1.0f;
  const int64_t plane = (int64_t)c * H * W;
  for (int m = tid; m < W; m += kThreads) {
    // E[jj] = Σ_k A_w[jstart+k][m]·G[jmin+jj][jstart+k]
    const int js = t.a_jstart[m];
    float wt[kKTMax];
#pragma unroll
    for (int k = 0; k < kKTMax; ++k) wt[k] = k < t.kt ? wts[k * W + m] : 0.f;
    float e[kJMax];
#pragma unroll
    for (int jj = 0; jj < kJMax; ++jj) {
      float s = 0.f;
      if (jj < jcnt) {
        const float* gr = G + (jmin + jj) * oW + js;
#pragma unroll
        for (int k = 0; k < kKTMax; ++k)
          if (k < t.kt && js + k < oW) s = fmaf(wt[k], gr[k], s);
      }
      e[jj] = s;
    }
    // out[i][m] = Σ_jj A_h[jmin+jj][i]·E[jj]
    // rows in batches of 8: the clamp-mask sources (x, ε) and `extra` of a batch are loaded before any use
    constexpr int kB = 8;
    const bool masked = a.has_mask && a.mask_src.eps && a.mask_src.clip;
    const float* mx = masked ? a.mask_src.x + n * a.mask_src.x_stride : nullptr;
    const float* me = masked ? a.mask_src.eps + n * a.mask_src.eps_stride : nullptr;
    const float* ex = a.extra ? a.extra + n * a.extra_stride : nullptr;
#pragma unroll 1
    for (int i0 = 0; i0 < kRA; i0 += kB) {
      float xv[kB], ev[kB], xt[kB];
#pragma unroll
      for (int b = 0; b < kB; ++b) {
        const int row = min(strip * kRA + i0 + b, H - 1);
        const int64_t off = plane + (int64_t)row * W + m;
        xv[b] = masked ? ldg_stream(mx + off) : 0.f;
        ev[b] = masked ? ldg_stream(me + off) : 0.f;
        xt[b] = ex ? ldg_stream(ex + off) : 0.f;
      }
#pragma unroll
      for (int b = 0; b < kB; ++b) {
        const int row = strip * kRA + i0 + b;
        if (row < H) {
          const float4* dr = reinterpret_cast<const float4*>(dht + (i0 + b) * kJMax);
          float s = 0.f;
#pragma unroll
          for (int q = 0; q < kJMax / 4; ++q) {
            const float4 w = dr[q];
            s = fmaf(w.x, e[4 * q + 0], s); s = fmaf(w.y, e[4 * q + 1], s);
            s = fmaf(w.z, e[4 * q + 2], s); s = fmaf(w.w, e[4 * q + 3], s);
          }
          float res = coef * s + xt[b];
          if (masked) res *= clamp_pass(x0_pre(xv[b], ev[b], a.mask_src.c1, a.mask_src.c2));
          stg_stream(a.g + n * a.g_stride + plane + (int64_t)row * W + m, res);
        }
      }
    }
  }
}

template <typename T>
int upload(const std::vector<T>& h, T** d) {
  DPS_CUDA(cudaMalloc(d, std::max<size_t>(1, h.size()) * sizeof(T)));
  if (!h.empty()) DPS_CUDA(cudaMemcpy(*d, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice));
  return DPS_OK;
}

// dense (out_len, in_len) matrix from the Resizer tables fov/weights (taps, out_len)
std::vector<double> dense_from_tables(const int32_t* fov, const float* w, int taps, int out_len, int in_len,
                                      bool* ok) {
  std::vector<double> A((size_t)out_len * in_len, 0.0);
  *ok = true;
  for (int k = 0; k < taps; ++k)
    for (int j = 0; j < out_len; ++j) {
      const int m = fov[(size_t)k * out_len + j];
      if (m < 0 || m >= in_len) { *ok = false; return A; }
      A[(size_t)j * in_len + m] += (double)w[(size_t)k * out_len + j];
    }
  return A;
}

size_t fwd_smem(const ResizeTables& t, int W, int oW) {
  return sizeof(float) * ((size_t)t.span * kRO + (size_t)kRO * W + 64 + (size_t)t.kw * oW + oW);
}
size_t adj_smem(int oH, int oW, int kt, int W) {
  return sizeof(float) * ((size_t)((oH * oW + 3) & ~3) + (size_t)kRA * kJMax + (size_t)kt * W);
}

}  // namespace

int resize_create(dps_operator* op, const int32_t* fov_h, const float* w_h, int taps_h, int out_h,
                  const int32_t* fov_w, const float* w_w, int taps_w, int out_w) {
  const int H = op->H, W = op->W;
  DPS_REQUIRE(out_h > 0 && out_w > 0 && out_h <= H && out_w <= W, DPS_ERR_UNSUPPORTED,
              "resize: only down-scaling is supported (%dx%d -> %dx%d)", H, W, out_h, out_w);
  bool ok_h, ok_w;
  std::vector<double> Ah = dense_from_tables(fov_h, w_h, taps_h, out_h, H, &ok_h);
  std::vector<double> Aw = dense_from_tables(fov_w, w_w, taps_w, out_w, W, &ok_w);
  DPS_REQUIRE(ok_h && ok_w, DPS_ERR_INVALID, "resize: field-of-view index out of range");
  ResizeTables* t = new ResizeTables();
  op->resize = t;
  // ---- forward H-pass blocks ----
  t->fstrips = (out_h + kRO - 1) / kRO;
  std::vector<int> rmin(t->fstrips), rcnt(t->fstrips);
  int span = 1;
  for (int s = 0; s < t->fstrips; ++s) {
    int lo = H, hi = -1;
    for (int j = s * kRO; j < std::min(out_h, (s + 1) * kRO); ++j)
      for (int m = 0; m < H; ++m)
        if (Ah[(size_t)j * H + m] != 0.0) { lo = std::min(lo, m); hi = std::max(hi, m); }
    if (hi < lo) { lo = 0; hi = 0; }
    rmin[s] = lo;
    rcnt[s] = hi - lo + 1;
    span = std::max(span, rcnt[s]);
  }
  t->span = span;
  std::vector<float> dh((size_t)t->fstrips * span * kRO, 0.f);
  for (int s = 0; s < t->fstrips; ++s)
    for (int rr = 0; rr < rcnt[s]; ++rr)
      for (int j = 0; j < kRO; ++j) {
        const int jo = s * kRO + j;
        if (jo < out_h) dh[((size_t)s * span + rr) * kRO + j] = (float)Ah[(size_t)jo * H + rmin[s] + rr];
      }
  // ---- forward W-pass windows ----
  std::vector<int> cstart(out_w);
  int kw = 1;
  for (int j = 0; j < out_w; ++j) {
    int lo = W, hi = -1;
    for (int m = 0; m < W; ++m)
      if (Aw[(size_t)j * W + m] != 0.0) { lo = std::min(lo, m); hi = std::max(hi, m); }
    if (hi < lo) { lo = 0; hi = 0; }
    cstart[j] = lo;
    kw = std::max(kw, hi - lo + 1);
  }
  t->kw = kw;
  std::vector<float> ww((size_t)out_w * kw, 0.f);
  for (int j = 0; j < out_w; ++j)
    for (int k = 0; k < kw; ++k)
      if (cstart[j] + k < W) ww[(size_t)j * kw + k] = (float)Aw[(size_t)j * W + cstart[j] + k];
  // ---- adjoint: A_hᵀ blocks per input-row strip ----
  t->astrips = (H + kRA - 1) / kRA;
  std::vector<int> jmin(t->astrips), jcnt(t->astrips);
  for (int s = 0; s < t->astrips; ++s) {
    int lo = out_h, hi = -1;
    for (int i = s * kRA; i < std::min(H, (s + 1) * kRA); ++i)
      for (int j = 0; j < out_h; ++j)
        if (Ah[(size_t)j * H + i] != 0.0) { lo = std::min(lo, j); hi = std::max(hi, j); }
    if (hi < lo) { lo = 0; hi = 0; }
    jmin[s] = lo;
    jcnt[s] = hi - lo + 1;
    DPS_REQUIRE(jcnt[s] <= kJMax, DPS_ERR_UNSUPPORTED, "resize: a %d-row strip touches %d measurement rows (> %d)",
                kRA, jcnt[s], kJMax);
  }
  std::vector<float> dht((size_t)t->astrips * kRA * kJMax, 0.f);
  for (int s = 0; s < t->astrips; ++s)
    for (int ii = 0; ii < kRA; ++ii)
      for (int jj = 0; jj < jcnt[s]; ++jj) {
        const int i = s * kRA + ii;
        if (i < H) dht[((size_t)s * kRA + ii) * kJMax + jj] = (float)Ah[(size_t)(jmin[s] + jj) * H + i];
      }
  // ---- adjoint: A_w column windows ----
  std::vector<int> jstart(W);
  int kt = 1;
  for (int m = 0; m < W; ++m) {
    int lo = out_w, hi = -1;
    for (int j = 0; j < out_w; ++j)
      if (Aw[(size_t)j * W + m] != 0.0) { lo = std::min(lo, j); hi = std::max(hi, j); }
    if (hi < lo) { lo = 0; hi = 0; }
    jstart[m] = lo;
    kt = std::max(kt, hi - lo + 1);
  }
  DPS_REQUIRE(kt <= kKTMax, DPS_ERR_UNSUPPORTED, "resize: an input column feeds %d measurement columns (> %d)", kt,
              kKTMax);
  t->kt = kt;
  std::vector<float> wt((size_t)W * kt, 0.f);
  for (int m = 0; m < W; ++m)
    for (int k = 0; k < kt; ++k)
      if (jstart[m] + k < out_w) wt[(size_t)m * kt + k] = (float)Aw[(size_t)(jstart[m] + k) * W + m];

  DPS_REQUIRE(fwd_smem(*t, W, out_w) <= 227 * 1024 && adj_smem(out_h, out_w, kt, W) <= 227 * 1024, DPS_ERR_UNSUPPORTED,
              "resize: tiles exceed shared memory");
  if (int rc = upload(rmin, &t->f_rmin)) return rc;
  if (int rc = upload(rcnt, &t->f_rcnt)) return rc;
  if (int rc = upload(dh, &t->f_dh)) return rc;
  if (int rc = upload(cstart, &t->f_cstart)) return rc;
  if (int rc = upload(ww, &t->f_ww)) return rc;
  if (int rc = upload(jmin, &t->a_jmin)) return rc;
  if (int rc = upload(jcnt, &t->a_jcnt)) return rc;
  if (int rc = upload(dht, &t->a_dht)) return rc;
  if (int rc = upload(jstart, &t->a_jstart)) return rc;
  if (int rc = upload(wt, &t->a_wt)) return rc;
  op->oC = op->C;
  op->oH = out_h;
  op->oW = out_w;
  op->P = op->C * t->fstrips;
  op->taps = taps_h;
  return DPS_OK;
}

void resize_destroy(dps_operator* op) {
  ResizeTables* t = op->resize;
  if (!t) return;
  cudaFree(t->f_rmin); cudaFree(t->f_rcnt); cudaFree(t->f_dh); cudaFree(t->f_cstart); cudaFree(t->f_ww);
  cudaFree(t->a_jmin); cudaFree(t->a_jcnt); cudaFree(t->a_dht); cudaFree(t->a_jstart); cudaFree(t->a_wt);
  delete t;
  op->resize = nullptr;
}

int resize_forward(const dps_operator* op, const FwdArgs& a, cudaStream_t st) {
  const ResizeTables& t = *op->resize;
  const size_t smem = fwd_smem(t, op->W, op->oW);
  static bool attr_set = false;
  if (!attr_set) {
    DPS_CUDA(cudaFuncSetAttribute(resize_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    DPS_CUDA(cudaFuncSetAttribute(resize_fwd_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    attr_set = true;
  }
  dim3 grid((unsigned)(op->C * t.fstrips), (unsigned)a.n);
  if (op->W == 256) {
    resize_fwd_pair_kernel<<<grid, kThreads, smem + sizeof(float) * kRO * 256, st>>>(t, op->C, op->H, op->oH, op->oW, a);
    DPS_LAUNCH_CHECK("resize_forward");
    return DPS_OK;
  }
  resize_fwd_kernel<<<grid, kThreads, smem, st>>>(t, op->C, op->H, op->W, op->oH, op->oW, a);
  DPS_LAUNCH_CHECK("resize_forward");
  return DPS_OK;
}

int resize_adjoint(const dps_operator* op, const AdjArgs& a, cudaStream_t st) {
  const ResizeTables& t = *op->resize;
  const size_t smem = adj_smem(op->oH, op->oW, t.kt, op->W);
  static bool attr_set = false;
  if (!attr_set) {
    DPS_CUDA(cudaFuncSetAttribute(resize_adj_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    attr_set = true;
  }
  dim3 grid((unsigned)(op->C * t.astrips), (unsigned)a.n);
  resize_adj_kernel<<<grid, kThreads, smem, st>>>(t, op->C, op->H, op->W, op->oH, op->oW, a);
  DPS_LAUNCH_CHECK("resize_adjoint");
  return DPS_OK;
}
