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
#include <cstdlib>
#include <vector>

#include "operator.cuh"

namespace {
constexpr int kThreads = 256;
constexpr int kRO = 8;     // forward: output rows per CTA
constexpr int kRA = 32;    // adjoint: input-space rows per CTA
constexpr int kJMax = 16;  // adjoint: measurement rows a strip may touch
constexpr int kRAs = 8;    // adjoint, small-grid variant: rows per CTA
constexpr int kJs = 8;     //   and the measurement rows such a strip may touch
constexpr int kKTMax = 12; // adjoint: measurement columns touching one input column
}  // namespace

// Strip windows travel in the kernel parameters (constant bank): a CTA knows its window without a dependent
// global load in front of its first data load.
constexpr int kMaxStrips = 128;
struct StripMeta {
  short lo[kMaxStrips];
  short cnt[kMaxStrips];
};

struct FwdTables {
  int fstrips = 0;        // ceil(out_h / kRO)
  int span = 0;           // max input-row window of a strip
  int kw = 0;             // max column window of an output column
  StripMeta rows;         // input-row window of every strip
  float* dh = nullptr;    // (fstrips, span, kRO): A_h[strip*kRO + j][lo + rr]
  int* cstart = nullptr;  // (out_w)
  float* wwt = nullptr;   // (kw, out_w) tap-major: A_w[j][cstart[j] + k]
};

// Streaming forward (resize_fwd_stream_kernel): input rows arrive in absolute 8-row chunks; chunk Q feeds the output rows
// jb(Q) … jb(Q)+kWO−1 with jb(Q) = D·Q − B, and the rows jb(Q) … jb(Q)+D−1 are complete once chunk Q has been added.
constexpr int kWO = 6;              // sliding window of output-row accumulators
constexpr int kStripsPerUnit = 4;   // a work unit = 4 output strips (32 measurement rows) of one plane
struct FwdStream {
  int ok = 0;
  int D = 0, B = 0;
  int upp = 0;              // units per plane
  short q_lo[8], q_hi[8];   // chunk range of every unit of a plane (q_hi may be a virtual chunk past the image: flush only)
  float* wq = nullptr;      // (H, kWO): weight of input row r for output row jb(r/8)+t
};

struct AdjStrips {
  int strips = 0;         // ceil(H / RA); 0: variant not available
  StripMeta rows;         // measurement-row window of every strip
  float* dht = nullptr;   // (strips, RA, KJ): A_h[lo + jj][strip*RA + i]
};

struct AdjCols {
  int kt = 0;             // max measurement-column window of an input column
  int* jstart = nullptr;  // (W)
  float* wtt = nullptr;   // (kt, W) tap-major: A_w[jstart[m] + k][m]
};

int resize_fused_create(dps_operator* op, const std::vector<double>& Ah, const std::vector<double>& Aw, int out_h, int out_w);

struct ResizeTables {
  FwdTables f;
  FwdStream fs;
  AdjStrips big;    // strips of kRA rows, ≤ kJMax measurement rows each
  AdjStrips small;  // strips of kRAs rows, ≤ kJs measurement rows each
  AdjCols cols;
};

namespace {

// shared-memory carve-up of the forward kernels
struct FwdSmem {
  float* dh;   // (span, kRO)
  float* V;    // (parts, kRO, W)
  float* red;  // 64
  float* wws;  // (kw, oW)
  int* css;    // (oW)
};
DPS_DEV FwdSmem fwd_carve(float* smem, const FwdTables& t, int parts, int W, int oW) {
  FwdSmem m;
  m.dh = smem;
  m.V = m.dh + t.span * kRO;
  m.red = m.V + parts * kRO * W;
  m.wws = m.red + 64;
  m.css = reinterpret_cast<int*>(m.wws + ((t.kw * oW + 3) & ~3));
  return m;
}
DPS_DEV void fwd_stage(const FwdSmem& m, const FwdTables& t, int strip, int rcnt, int oW, int tid, int nt) {
  stage_async(m.dh, t.dh + (int64_t)strip * t.span * kRO, rcnt * kRO, tid, nt);
  stage_async(m.wws, t.wwt, t.kw * oW, tid, nt);
  stage_async(reinterpret_cast<float*>(m.css), reinterpret_cast<const float*>(t.cstart), oW, tid, nt);
}

// The measurement values a thread needs in the W pass are fetched early (before the H-pass barrier) so that their
// DRAM latency is not a serial step at the end of the CTA: the first kYPre outputs of every thread.
constexpr int kYPre = 2;
struct YPre {
  float v[kYPre];
};
DPS_DEV YPre fwd_y_prefetch(int oH, int oW, int strip, int c, int n, int tid, int nt, const FwdArgs& a) {
  YPre y;
  const float* yp = a.y ? a.y + n * a.y_stride + (int64_t)c * oH * oW : nullptr;
#pragma unroll
  for (int u = 0; u < kYPre; ++u) {
    const int i = tid + u * nt;
    const int orow = strip * kRO + i / oW;
    y.v[u] = (yp && i < kRO * oW && orow < oH) ? ldg_ro(yp + (int64_t)strip * kRO * oW + i) : 0.f;
  }
  return y;
}

// One output of the W pass: Σ_k A_w[jc][cs+k]·V[j][cs+k], residual, store, Σr², Σ|r|.  i = j·oW + jc inside the strip.
DPS_DEV void wpass_one(const FwdSmem& m, const FwdTables& t, int W, int oH, int oW, int strip, int i, const float* yp,
                       bool have, float yv, float* outp, float& sq, float& ab) {
  const int j = i / oW, jc = i - j * oW;
  const int orow = strip * kRO + j;
  if (orow >= oH) return;
  const int64_t o = (int64_t)orow * oW + jc;
  if (!have) yv = yp ? ldg_ro(yp + o) : 0.f;
  const int cs = m.css[jc];
  const float* vr = m.V + j * W + cs;
  float s = 0.f;
#pragma unroll 8
  for (int k = 0; k < t.kw; ++k) s = fmaf(m.wws[k * oW + jc], (cs + k < W) ? vr[k] : 0.f, s);
  const float res = yp ? __fsub_rn(yv, s) : s;
  stg_stream(outp + o, res);
  sq += res * res;
  ab += fabsf(res);
}

// W pass + residual + Σr², Σ|r| from the (kRO, W) tile V
// `slot`: the strip's partial-sum slot c·fstrips + strip.
DPS_DEV void fwd_wpass(const FwdSmem& m, const FwdTables& t, int C, int W, int oH, int oW, int strip, int c, int n,
                       int tid, int nt, const FwdArgs& a, const YPre& ypre, int slot) {
  float sq = 0.f, ab = 0.f;
  float* outp = a.out + ((int64_t)n * C + c) * oH * oW;
  const float* yp = a.y ? a.y + n * a.y_stride + (int64_t)c * oH * oW : nullptr;
#pragma unroll
  for (int u = 0; u < kYPre; ++u)
    if (tid + u * nt < kRO * oW) wpass_one(m, t, W, oH, oW, strip, tid + u * nt, yp, true, ypre.v[u], outp, sq, ab);
  for (int i = tid + kYPre * nt; i < kRO * oW; i += nt) wpass_one(m, t, W, oH, oW, strip, i, yp, false, 0.f, outp, sq, ab);
  if (a.partials) {
    block_sum2(sq, ab, m.red);
    if (tid == 0) {
      float* pp = a.partials + ((int64_t)n * (C * t.fstrips) + slot) * 2;
      pp[0] = sq;
      pp[1] = ab;
    }
  }
}

__global__ void __launch_bounds__(kThreads) resize_fwd_kernel(const FwdTables t, int C, int H, int W, int oH, int oW,
                                                              const FwdArgs a) {
  extern __shared__ __align__(128) float smem[];
  const FwdSmem m = fwd_carve(smem, t, 1, W, oW);
  const int strip = blockIdx.x % t.fstrips;
  const int c = blockIdx.x / t.fstrips;
  const int n = blockIdx.y;
  const int tid = threadIdx.x;
  const int rmin = t.rows.lo[strip], rcnt = t.rows.cnt[strip];
  fwd_stage(m, t, strip, rcnt, oW, tid, kThreads);
  const YPre ypre = fwd_y_prefetch(oH, oW, strip, c, n, tid, kThreads, a);
  stage_wait();
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
          const float4 w0 = *reinterpret_cast<const float4*>(m.dh + (rr0 + b) * kRO);
          const float4 w1 = *reinterpret_cast<const float4*>(m.dh + (rr0 + b) * kRO + 4);
          acc[0] = fmaf(w0.x, v, acc[0]); acc[1] = fmaf(w0.y, v, acc[1]);
          acc[2] = fmaf(w0.z, v, acc[2]); acc[3] = fmaf(w0.w, v, acc[3]);
          acc[4] = fmaf(w1.x, v, acc[4]); acc[5] = fmaf(w1.y, v, acc[5]);
          acc[6] = fmaf(w1.z, v, acc[6]); acc[7] = fmaf(w1.w, v, acc[7]);
        }
      }
    }
#pragma unroll
    for (int j = 0; j < kRO; ++j) m.V[j * W + col] = acc[j];
  }
  __syncthreads();
  fwd_wpass(m, t, C, W, oH, oW, strip, c, n, tid, kThreads, a, ypre, blockIdx.x);
}

// Bulk-copy variant for W = 256 (the default): the strip's input window is fetched by the TMA engine — 1-D bulk copies
// of 8 image rows (8 KB per tensor, contiguous in a plane) into a 3-stage shared-memory ring, completion on mbarriers —
// so 48 KB per CTA (144 KB per SM at 3 CTAs) are in flight without holding a register, and the threads only ever wait
// on shared memory.  Compute is the pair kernel's: 128 column pairs × 2 row groups on packed FFMA2, then the W pass.
// Stage geometry of the streaming kernels.  tools/tma_stream_probe.cu (B200): persistent CTAs that each stream their own
// region reach 4.0-4.2 TB/s with 8 KB bulk copies, 5.5 TB/s with 16 KB and 6.0 TB/s with 32 KB (a grid-wide linear sweep:
// 6.0 TB/s at any size) — DRAM row locality wants large contiguous requests, so a stage holds 16 / 32 rows per tensor.
constexpr int kSaStages = 3;          // streaming adjoint: ring stages
constexpr int kSaRows = 16;           //   rows per stage and tensor (16 KB per bulk copy)
constexpr int kSaThreads = 256 + 32;  // streaming kernels: 256 consumers + one producer warp
constexpr int kCR = 8;          // rows per chunk
constexpr int kStagesStd = 3;   // chunks in flight (4 CTAs per SM)
constexpr int kStagesDeep = 6;  // very small grids: the whole ×4 window (≤ 6 chunks) in flight at once — one DRAM round
                                // trip per CTA instead of two, no recycle barrier; same arithmetic, same bits

template <int kStages>
__global__ void __launch_bounds__(256, kStages == kStagesStd ? 4 : 2) resize_fwd_bulk_kernel(const FwdTables t, int C, int H, int oH,
                                                                                    int oW, const FwdArgs a) {
  constexpr int W = 256, W2 = 128, kT = 256, kParts = 2;
  extern __shared__ __align__(128) float smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem);  // kStages mbarriers in the first 128 bytes
  float* ring = smem + 32;                             // (kStages, 2, kCR, W): x rows then ε rows of a chunk
  FwdSmem m = fwd_carve(ring + kStages * 2 * kCR * W, t, 0, W, oW);
  m.V = ring;  // the H-pass result (2, kRO, W) reuses the ring once every chunk has been consumed → 55 KB, 4 CTAs per SM
  static_assert(kStages * 2 * kCR >= kParts * kRO, "V must fit in the ring");
  const int strip = blockIdx.x % t.fstrips;
  const int c = blockIdx.x / t.fstrips;
  const int n = blockIdx.y;
  const int tid = threadIdx.x;
  const int rmin = t.rows.lo[strip], rcnt = t.rows.cnt[strip];
  const int nchunks = (rcnt + kCR - 1) / kCR;
  const int64_t plane = (int64_t)c * H * W;
  const float* x = a.src.x + n * a.src.x_stride + plane + (int64_t)rmin * W;
  const float* eps = a.src.eps ? a.src.eps + n * a.src.eps_stride + plane + (int64_t)rmin * W : nullptr;
  auto issue = [&](int k) {  // one thread: arm the stage's barrier, start the copies of chunk k
    const int stage = k % kStages;
    const unsigned bytes = (unsigned)(min(kCR, rcnt - k * kCR) * W * sizeof(float));
    float* dst = ring + stage * 2 * kCR * W;
    mbar_expect_tx(&bars[stage], eps ? 2 * bytes : bytes);
    bulk_load(dst, x + (int64_t)k * kCR * W, bytes, &bars[stage]);
    if (eps) bulk_load(dst + kCR * W, eps + (int64_t)k * kCR * W, bytes, &bars[stage]);
  };
  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < kStages; ++s) mbar_init(&bars[s], 1);
    mbar_init_fence();
    for (int k = 0; k < min(kStages, nchunks); ++k) issue(k);
  }
  fwd_stage(m, t, strip, rcnt, oW, tid, kT);
  const YPre ypre = fwd_y_prefetch(oH, oW, strip, c, n, tid, kT, a);
  stage_wait();
  __syncthreads();  // tables staged, barriers initialised
  const int part = tid >> 7, cp = tid & 127;
  float2 acc[kRO];
#pragma unroll
  for (int j = 0; j < kRO; ++j) acc[j] = make_float2(0.f, 0.f);
#pragma unroll 1
  for (int k = 0; k < nchunks; ++k) {
    const int stage = k % kStages;
    mbar_wait(&bars[stage], (unsigned)((k / kStages) & 1));
    const float2* xs = reinterpret_cast<const float2*>(ring + stage * 2 * kCR * W);
    const float2* es = xs + kCR * W2;
    const int rows = min(kCR, rcnt - k * kCR);
#pragma unroll
    for (int q = 0; q < kCR / kParts; ++q) {  // row group `part` takes rows part·4 … part·4+3 of the chunk
      const int rl = part * (kCR / kParts) + q;
      if (rl < rows) {
        const float2 xv = xs[rl * W2 + cp];
        const float2 v = eps ? x0_pair(xv, es[rl * W2 + cp], a.src.c1, a.src.c2, a.src.clip) : xv;
        const float* wr = m.dh + (k * kCR + rl) * kRO;
        const float4 w0 = *reinterpret_cast<const float4*>(wr);
        const float4 w1 = *reinterpret_cast<const float4*>(wr + 4);
        acc[0] = __ffma2_rn(make_float2(w0.x, w0.x), v, acc[0]); acc[1] = __ffma2_rn(make_float2(w0.y, w0.y), v, acc[1]);
        acc[2] = __ffma2_rn(make_float2(w0.z, w0.z), v, acc[2]); acc[3] = __ffma2_rn(make_float2(w0.w, w0.w), v, acc[3]);
        acc[4] = __ffma2_rn(make_float2(w1.x, w1.x), v, acc[4]); acc[5] = __ffma2_rn(make_float2(w1.y, w1.y), v, acc[5]);
        acc[6] = __ffma2_rn(make_float2(w1.z, w1.z), v, acc[6]); acc[7] = __ffma2_rn(make_float2(w1.w, w1.w), v, acc[7]);
      }
    }
    if (k + kStages < nchunks) {  // recycle the stage: every thread has read it, then one thread refills it
      __syncthreads();
      if (tid == 0) issue(k + kStages);
    }
  }
  __syncthreads();  // every chunk consumed: the ring becomes V
  {
    float* dstp = m.V + part * kRO * W;
#pragma unroll
    for (int j = 0; j < kRO; ++j) *reinterpret_cast<float2*>(dstp + j * W + 2 * cp) = acc[j];
  }
  __syncthreads();
  for (int i = tid; i < kRO * W / 4; i += kT) {  // V[0] += V[1], fixed order
    float4 s0 = *reinterpret_cast<const float4*>(m.V + i * 4);
    const float4 s1 = *reinterpret_cast<const float4*>(m.V + kRO * W + i * 4);
    s0.x += s1.x; s0.y += s1.y; s0.z += s1.z; s0.w += s1.w;
    *reinterpret_cast<float4*>(m.V + i * 4) = s0;
  }
  __syncthreads();
  fwd_wpass(m, t, C, W, oH, oW, strip, c, n, tid, kT, a, ypre, blockIdx.x);
}

// Lean strip forward — OPT-IN (DPSTTC_RESIZE_FWD_LEAN=1), prepared from the N = 8 ncu capture of the bulk kernel
// (profiles/r1l_resize_fwd_n8_sass.md): at bench size the kernel is bound by each warp's own instruction stream (2.7 warps
// per scheduler, one issue every 8.9 cycles, 12 % of the stall samples on instruction fetch), and only 12 % of its 2.46 M
// warp instructions are FFMA2 — the rest is per-row control flow (a BSSY/BSYNC-guarded `rl < rows` block and a uniform
// `eps` branch per row), runtime index arithmetic, and a W pass that divides by a runtime oW and walks a runtime tap count.
// Same arithmetic in the same order as resize_fwd_bulk_kernel (hence the same bits), with the control flow folded away:
//   * specialised for the DPS path proper: x̂₀ source with clipping, W = 256, oW = 64, KW taps at compile time;
//   * H pass: rows past the window's end contribute v = 0 through a select instead of a branch (acc + w·0 = acc exactly:
//     the weights row index is clamped into the staged block, so w is finite; acc is never −0), so the four rows of a
//     thread's chunk share one straight-line block and their shared-memory loads issue together;
//   * W pass: i → (j, jc) by shifts, taps fully unrolled.
// Not launched unless the environment variable is set; bit-identity with the default kernel is checked by
//   python tools/variant_check.py --op sr4 --n 8 --env DPSTTC_RESIZE_FWD_LEAN=0 --env DPSTTC_RESIZE_FWD_LEAN=1
template <int kStages, int KW>
__global__ void __launch_bounds__(256, kStages == kStagesStd ? 4 : 2) resize_fwd_lean_kernel(const FwdTables t, int C, int H, int oH,
                                                                                    const FwdArgs a) {
  constexpr int W = 256, W2 = 128, kT = 256, kParts = 2, oW = 64;
  static_assert(kYPre * kT == kRO * oW, "every W-pass output of a thread has its measurement value prefetched");
  extern __shared__ __align__(128) float smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem);
  float* ring = smem + 32;
  FwdSmem m = fwd_carve(ring + kStages * 2 * kCR * W, t, 0, W, oW);
  m.V = ring;
  static_assert(kStages * 2 * kCR >= kParts * kRO, "V must fit in the ring");
  const int strip = blockIdx.x % t.fstrips;
  const int c = blockIdx.x / t.fstrips;
  const int n = blockIdx.y;
  const int tid = threadIdx.x;
  const int rmin = t.rows.lo[strip], rcnt = t.rows.cnt[strip];
  const int nchunks = (rcnt + kCR - 1) / kCR;
  const int64_t plane = (int64_t)c * H * W;
  const float* x = a.src.x + n * a.src.x_stride + plane + (int64_t)rmin * W;
  const float* eps = a.src.eps + n * a.src.eps_stride + plane + (int64_t)rmin * W;
  auto issue = [&](int k) {
    const int stage = k % kStages;
    const unsigned bytes = (unsigned)(min(kCR, rcnt - k * kCR) * W * sizeof(float));
    float* dst = ring + stage * 2 * kCR * W;
    mbar_expect_tx(&bars[stage], 2 * bytes);
    bulk_load(dst, x + (int64_t)k * kCR * W, bytes, &bars[stage]);
    bulk_load(dst + kCR * W, eps + (int64_t)k * kCR * W, bytes, &bars[stage]);
  };
  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < kStages; ++s) mbar_init(&bars[s], 1);
    mbar_init_fence();
    for (int k = 0; k < min(kStages, nchunks); ++k) issue(k);
  }
  fwd_stage(m, t, strip, rcnt, oW, tid, kT);
  const YPre ypre = fwd_y_prefetch(oH, oW, strip, c, n, tid, kT, a);
  stage_wait();
  __syncthreads();
  const int part = tid >> 7, cp = tid & 127;
  const float c1 = a.src.c1, c2 = a.src.c2;
  float2 acc[kRO];
#pragma unroll
  for (int j = 0; j < kRO; ++j) acc[j] = make_float2(0.f, 0.f);
#pragma unroll 1
  for (int k = 0; k < nchunks; ++k) {
    const int stage = k % kStages;
    mbar_wait(&bars[stage], (unsigned)((k / kStages) & 1));
    const float2* xs = reinterpret_cast<const float2*>(ring + stage * 2 * kCR * W) + part * (kCR / kParts) * W2 + cp;
    const float2* es = xs + kCR * W2;
    const int left = rcnt - k * kCR - part * (kCR / kParts);  // rows of this thread's group that exist in the chunk
    float2 v[kCR / kParts];
#pragma unroll
    for (int q = 0; q < kCR / kParts; ++q) {
      const float2 p = x0_pair_bounds(xs[q * W2], es[q * W2], c1, c2, -1.0f, 1.0f);
      v[q] = make_float2(q < left ? p.x : 0.f, q < left ? p.y : 0.f);  // a select: stale rows (even NaN) never reach an FMA
    }
    const int wrow0 = k * kCR + part * (kCR / kParts);
#pragma unroll
    for (int q = 0; q < kCR / kParts; ++q) {
      const float* wr = m.dh + min(wrow0 + q, rcnt - 1) * kRO;
      const float4 w0 = *reinterpret_cast<const float4*>(wr);
      const float4 w1 = *reinterpret_cast<const float4*>(wr + 4);
      acc[0] = __ffma2_rn(make_float2(w0.x, w0.x), v[q], acc[0]); acc[1] = __ffma2_rn(make_float2(w0.y, w0.y), v[q], acc[1]);
      acc[2] = __ffma2_rn(make_float2(w0.z, w0.z), v[q], acc[2]); acc[3] = __ffma2_rn(make_float2(w0.w, w0.w), v[q], acc[3]);
      acc[4] = __ffma2_rn(make_float2(w1.x, w1.x), v[q], acc[4]); acc[5] = __ffma2_rn(make_float2(w1.y, w1.y), v[q], acc[5]);
      acc[6] = __ffma2_rn(make_float2(w1.z, w1.z), v[q], acc[6]); acc[7] = __ffma2_rn(make_float2(w1.w, w1.w), v[q], acc[7]);
    }
    if (k + kStages < nchunks) {
      __syncthreads();
      if (tid == 0) issue(k + kStages);
    }
  }
  __syncthreads();
  {
    float* dstp = m.V + part * kRO * W;
#pragma unroll
    for (int j = 0; j < kRO; ++j) *reinterpret_cast<float2*>(dstp + j * W + 2 * cp) = acc[j];
  }
  __syncthreads();
#pragma unroll
  for (int u = 0; u < kRO * W / 4 / kT; ++u) {  // V[0] += V[1], fixed order
    const int i = tid + u * kT;
    float4 s0 = *reinterpret_cast<const float4*>(m.V + i * 4);
    const float4 s1 = *reinterpret_cast<const float4*>(m.V + kRO * W + i * 4);
    s0.x += s1.x; s0.y += s1.y; s0.z += s1.z; s0.w += s1.w;
    *reinterpret_cast<float4*>(m.V + i * 4) = s0;
  }
  __syncthreads();
  // ---- W pass: the thread → output mapping, tap order and partial-sum order of fwd_wpass / wpass_one ----
  float sq = 0.f, ab = 0.f;
  float* outp = a.out + ((int64_t)n * C + c) * oH * oW;
  const bool has_y = a.y != nullptr;
#pragma unroll
  for (int u = 0; u < kYPre; ++u) {
    const int i = tid + u * kT;
    const int j = i >> 6, jc = i & (oW - 1);
    const int orow = strip * kRO + j;
    if (orow < oH) {
      const int cs = m.css[jc];
      const float* vr = m.V + j * W + cs;
      float s = 0.f;
#pragma unroll
      for (int k = 0; k < KW; ++k) s = fmaf(m.wws[k * oW + jc], (cs + k < W) ? vr[k] : 0.f, s);
      const float res = has_y ? __fsub_rn(ypre.v[u], s) : s;
      stg_stream(outp + (int64_t)orow * oW + jc, res);
      sq += res * res;
      ab += fabsf(res);
    }
  }
  if (a.partials) {
    block_sum2(sq, ab, m.red);
    if (tid == 0) {
      float* pp = a.partials + ((int64_t)n * (C * t.fstrips) + blockIdx.x) * 2;
      pp[0] = sq;
      pp[1] = ab;
    }
  }
}

// Streaming forward for full machines (W = 256; ×4 and ×8 bicubic): persistent CTAs of three roles.
//   producer (1 warp) : a work unit is 4 consecutive output strips of one plane; its input rows are streamed ONCE, in
//                       absolute 8-row chunks, through a ring of kSfStages TMA stages (x rows | ε rows) — the 1.5× halo
//                       re-read of the strip kernels disappears and loads never pause;
//   H warps (4)       : a thread owns a column pair and keeps a sliding window of kWO output-row accumulators (two partial
//                       sums: rows 0-3 and rows 4-7 of every chunk).  Chunk Q adds to rows jb(Q)…jb(Q)+5, after which the
//                       first D of them are complete and go to the V tile of their strip (double-buffered);
//   W warps (4)       : per completed strip the W pass, residual, store and partial sums — concurrently with the H warps'
//                       next strip.  The 128 threads play the 256 threads of the strip kernels (two virtual threads each,
//                       two virtual warps per warp), so the reduction tree and therefore the partial sums are bit-identical.
// Bit-identity with resize_fwd_bulk_kernel: chunks are aligned to absolute multiples of 8 rows in both, each partial sum
// adds its rows in ascending order, V = part0 + part1, and both use wpass_one.
// Variants that were measured and lost (N = 128, this kernel 50.7 µs): 16-row stages × 3 with a single V tile 64 µs,
// 32-row stages × 3 at one CTA per SM 66 µs, 8 H warps (one row half each) + W pass on the split tile 63 µs, 8 H warps
// with one column per thread (scalar FFMA) 56.8 µs, chunks issued in pairs 50.9 µs, in triples 53.1 µs.
// wpass_one for oW = 64 and KW taps at compile time (the form validated bit-identical in resize_fwd_lean_kernel): i → (j, jc) by
// shifts, taps unrolled.  Used by the streaming forward's W warps under DPSTTC_RESIZE_FWD_LEAN=1: with the generic wpass_one
// (≈180 instructions per output, 4 outputs per thread and strip) the four W warps, not the H warps, pace a CTA.
template <int KW>
DPS_DEV void wpass_one_lean(const FwdSmem& m, int oH, int strip, int i, bool has_y, float yv, float* outp, float& sq, float& ab) {
  constexpr int W = 256, oW = 64;
  const int j = i >> 6, jc = i & (oW - 1);
  const int orow = strip * kRO + j;
  if (orow >= oH) return;
  const int cs = m.css[jc];
  const float* vr = m.V + j * W + cs;
  float s = 0.f;
#pragma unroll
  for (int k = 0; k < KW; ++k) s = fmaf(m.wws[k * oW + jc], (cs + k < W) ? vr[k] : 0.f, s);
  const float res = has_y ? __fsub_rn(yv, s) : s;
  stg_stream(outp + (int64_t)orow * oW + jc, res);
  sq += res * res;
  ab += fabsf(res);
}

constexpr int kSfStages = 5;

template <int D, bool kLeanW = false>  // kLeanW: opt-in (oW = 64, 16 taps), see wpass_one_lean
__global__ void __launch_bounds__(kSaThreads, 2) resize_fwd_stream_kernel(const FwdTables t, const FwdStream fs, int C, int H,
                                                                            int oH, int oW, int units, const FwdArgs a) {
  constexpr int W = 256, W2 = 128, kStageFloats = 2 * 8 * W;
  extern __shared__ __align__(128) float smem[];
  uint64_t* full = reinterpret_cast<uint64_t*>(smem);
  uint64_t* empty = full + kSfStages;
  uint64_t* vfull = empty + kSfStages;  // [2]
  uint64_t* vempty = vfull + 2;         // [2]
  static_assert(2 * kSfStages + 4 <= 16, "barriers must fit in the first 128 bytes");
  float* ring = smem + 32;
  float* Vbuf = ring + kSfStages * kStageFloats;  // (2, kRO, W)
  float* wq = Vbuf + 2 * kRO * W;                 // (H, kWO)
  FwdSmem m;
  m.dh = nullptr;
  m.V = Vbuf;
  m.red = wq + H * kWO;                           // (2, 64)
  m.wws = m.red + 128;
  m.css = reinterpret_cast<int*>(m.wws + ((t.kw * oW + 3) & ~3));
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < kSfStages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 4); }
    mbar_init(&vfull[0], 4); mbar_init(&vfull[1], 4);
    mbar_init(&vempty[0], 4); mbar_init(&vempty[1], 4);
    mbar_init_fence();
  }
  stage_async(wq, fs.wq, H * kWO, tid, kSaThreads);
  stage_async(m.wws, t.wwt, t.kw * oW, tid, kSaThreads);
  stage_async(reinterpret_cast<float*>(m.css), reinterpret_cast<const float*>(t.cstart), oW, tid, kSaThreads);
  stage_wait();
  __syncthreads();
  const int upp = fs.upp, nchunks = H / 8;
  const bool has_eps = a.src.eps != nullptr;

  if (warp == 8) {  // ---------------- producer ----------------
    if (lane != 0) return;
    int it = 0;
    for (int u = blockIdx.x; u < units; u += gridDim.x) {
      const int k = u % upp, c = (u / upp) % C, n = u / (upp * C);
      const int64_t plane = (int64_t)c * H * W;
      const float* x = a.src.x + n * a.src.x_stride + plane;
      const float* eps = has_eps ? a.src.eps + n * a.src.eps_stride + plane : nullptr;
      const int q1 = min((int)fs.q_hi[k], nchunks - 1);
#pragma unroll 1
      for (int Q = fs.q_lo[k]; Q <= q1; ++Q, ++it) {
        const int s = it % kSfStages;
        mbar_wait_guarded(&empty[s], (unsigned)(((it / kSfStages) & 1) ^ 1));
        float* dst = ring + s * kStageFloats;
        const unsigned bytes = 8 * W * sizeof(float);
        mbar_expect_tx(&full[s], has_eps ? 2 * bytes : bytes);
        bulk_load(dst, x + (int64_t)Q * 8 * W, bytes, &full[s]);
        if (has_eps) bulk_load(dst + 8 * W, eps + (int64_t)Q * 8 * W, bytes, &full[s]);
      }
    }
    return;
  }

  if (warp < 4) {  // ---------------- H warps ----------------
    const int cp = tid;
    const float c1 = a.src.c1, c2 = a.src.c2;
    const float clamp_hi = a.src.clip ? 1.0f : __int_as_float(0x7f800000), clamp_lo = -clamp_hi;
    int it = 0, sc = 0;  // chunks consumed, strips produced
    for (int u = blockIdx.x; u < units; u += gridDim.x) {
      const int k = u % upp;
      const int j_lo = k * kStripsPerUnit * kRO, j_hi = min(oH, j_lo + kStripsPerUnit * kRO);
      float2 acc0[kWO], acc1[kWO];
#pragma unroll
      for (int tt = 0; tt < kWO; ++tt) acc0[tt] = acc1[tt] = make_float2(0.f, 0.f);
#pragma unroll 1
      for (int Q = fs.q_lo[k]; Q <= fs.q_hi[k]; ++Q) {
        if (Q < nchunks) {
          const int s = it % kSfStages;
          mbar_wait_guarded(&full[s], (unsigned)((it / kSfStages) & 1));
          const float2* xs = reinterpret_cast<const float2*>(ring + s * kStageFloats) + cp;
          const float2* es = xs + 8 * W2;
          const float2* wr = reinterpret_cast<const float2*>(wq + Q * 8 * kWO);
          float2 v[8];
          if (has_eps) {
#pragma unroll
            for (int rl = 0; rl < 8; ++rl) v[rl] = x0_pair_bounds(xs[rl * W2], es[rl * W2], c1, c2, clamp_lo, clamp_hi);
          } else {
#pragma unroll
            for (int rl = 0; rl < 8; ++rl) v[rl] = xs[rl * W2];
          }
#pragma unroll
          for (int rl = 0; rl < 8; ++rl) {
            const float2 wa = wr[rl * 3], wb = wr[rl * 3 + 1], wc = wr[rl * 3 + 2];
            float2* acc = rl < 4 ? acc0 : acc1;
            acc[0] = __ffma2_rn(make_float2(wa.x, wa.x), v[rl], acc[0]); acc[1] = __ffma2_rn(make_float2(wa.y, wa.y), v[rl], acc[1]);
            acc[2] = __ffma2_rn(make_float2(wb.x, wb.x), v[rl], acc[2]); acc[3] = __ffma2_rn(make_float2(wb.y, wb.y), v[rl], acc[3]);
            acc[4] = __ffma2_rn(make_float2(wc.x, wc.x), v[rl], acc[4]); acc[5] = __ffma2_rn(make_float2(wc.y, wc.y), v[rl], acc[5]);
          }
          __syncwarp();
          if (lane == 0) mbar_arrive(&empty[s]);
          ++it;
        }
        // rows jb … jb+D−1 are complete: V[row] = part0 + part1 into the strip's tile
        const int jb = D * Q - fs.B;
#pragma unroll
        for (int d = 0; d < D; ++d) {
          const int j = jb + d;
          if (j >= j_lo && j < j_hi) {
            const int buf = sc & 1;
            if ((j & (kRO - 1)) == 0) mbar_wait_guarded(&vempty[buf], (unsigned)(((sc >> 1) & 1) ^ 1));  // tile free?
            float2 sum = acc0[d];
            sum.x += acc1[d].x;
            sum.y += acc1[d].y;
            *reinterpret_cast<float2*>(Vbuf + (buf * kRO + (j & (kRO - 1))) * W + 2 * cp) = sum;
            if ((j & (kRO - 1)) == kRO - 1) {  // strip complete: hand the tile to the W warps
              __syncwarp();
              if (lane == 0) mbar_arrive(&vfull[buf]);
              ++sc;
            }
          }
        }
#pragma unroll
        for (int tt = 0; tt < kWO; ++tt) {
          acc0[tt] = tt + D < kWO ? acc0[tt + D] : make_float2(0.f, 0.f);
          acc1[tt] = tt + D < kWO ? acc1[tt + D] : make_float2(0.f, 0.f);
        }
      }
    }
    return;
  }

  // ---------------- W warps: thread wt plays the strip kernels' threads wt and wt + 128 ----------------
  const int wt = tid - 128, ww = wt >> 5;
  int sc = 0;
  for (int u = blockIdx.x; u < units; u += gridDim.x) {
    const int k = u % upp, c = (u / upp) % C, n = u / (upp * C);
    const int j_lo = k * kStripsPerUnit * kRO, j_hi = min(oH, j_lo + kStripsPerUnit * kRO);
    float* outp = a.out + ((int64_t)n * C + c) * oH * oW;
    const float* yp = a.y ? a.y + n * a.y_stride + (int64_t)c * oH * oW : nullptr;
#pragma unroll 1
    for (int strip = j_lo / kRO; strip < j_hi / kRO; ++strip, ++sc) {
      const int buf = sc & 1;
      // measurement values first: their latency hides behind the wait for the tile
      float yv[2][kYPre];
#pragma unroll
      for (int vtt = 0; vtt < 2; ++vtt)
#pragma unroll
        for (int uu = 0; uu < kYPre; ++uu) {
          const int i = wt + vtt * 128 + uu * 256;
          yv[vtt][uu] = (yp && i < kRO * oW) ? ldg_ro(yp + (int64_t)strip * kRO * oW + i) : 0.f;
        }
      mbar_wait_guarded(&vfull[buf], (unsigned)((sc >> 1) & 1));
      m.V = Vbuf + buf * kRO * W;
      float sq[2] = {0.f, 0.f}, ab[2] = {0.f, 0.f};
#pragma unroll
      for (int vtt = 0; vtt < 2; ++vtt) {
        const int vt = wt + vtt * 128;
#pragma unroll
        for (int uu = 0; uu < kYPre; ++uu) {
          if constexpr (kLeanW) {  // kRO·oW = 512 = kYPre·256: every output of a thread has its measurement value prefetched
            wpass_one_lean<16>(m, oH, strip, vt + uu * 256, yp != nullptr, yv[vtt][uu], outp, sq[vtt], ab[vtt]);
          } else {
            if (vt + uu * 256 < kRO * oW)
              wpass_one(m, t, W, oH, oW, strip, vt + uu * 256, yp, true, yv[vtt][uu], outp, sq[vtt], ab[vtt]);
          }
        }
        if constexpr (!kLeanW)
          for (int i = vt + kYPre * 256; i < kRO * oW; i += 256) wpass_one(m, t, W, oH, oW, strip, i, yp, false, 0.f, outp, sq[vtt], ab[vtt]);
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&vempty[buf]);  // the tile has been read
      if (a.partials) {  // block_sum2's tree for 8 (virtual) warps: virtual warp of (ww, vtt) = ww + 4·vtt
        float* red = m.red + buf * 64;
#pragma unroll
        for (int vtt = 0; vtt < 2; ++vtt) {
          const float s2 = warp_sum(sq[vtt]), a2 = warp_sum(ab[vtt]);
          if (lane == 0) {
            red[ww + 4 * vtt] = s2;
            red[32 + ww + 4 * vtt] = a2;
          }
        }
        named_bar_sync(2, 128);
        if (ww == 0) {
          float s2 = lane < 8 ? red[lane] : 0.0f;
          float a2 = lane < 8 ? red[32 + lane] : 0.0f;
          s2 = warp_sum(s2);
          a2 = warp_sum(a2);
          if (lane == 0) {
            float* pp = a.partials + ((int64_t)n * (C * t.fstrips) + c * t.fstrips + strip) * 2;
            pp[0] = s2;
            pp[1] = a2;
          }
        }
      }
    }
  }
}

// Adjoint.  RA input rows per CTA, KJ = the most measurement rows such a strip may touch.  Only the measurement
// rows the strip touches are staged.  The clamp-mask / extra loads of the first row batch are issued before the
// tiles are staged, so a CTA's dependent chain is one round trip.
// kBulk: the clamp-mask sources (x, ε rows of the strip, contiguous in a plane) are fetched by the TMA engine into shared
// memory in chunks of 8 rows, one mbarrier per chunk, all issued by one thread at kernel start — nothing is held in
// registers and no load is on a thread's dependent path.
template <int RA, int KJ, int KT, bool kBulk>
__global__ void __launch_bounds__(kThreads, RA == kRAs ? 4 : 3) resize_adj_kernel(const AdjStrips at, const AdjCols ac, int C, int H, int W,
                                                              int oH, int oW, const AdjArgs a) {
  extern __shared__ __align__(128) float smem[];
  constexpr int kChunks = RA / 8;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem);  // kBulk: kChunks mbarriers in the first 128 bytes
  float* xs = smem + (kBulk ? 32 : 0);                  // kBulk: (RA, W) x rows, then (RA, W) ε rows
  float* G = xs + (kBulk ? 2 * RA * W : 0);             // (KJ, oW): the measurement rows this strip touches
  float* dht = G + ((KJ * oW + 3) & ~3);                // (RA, KJ), 16-byte aligned
  float* wts = dht + RA * KJ;                           // (kt, W): column weights, tap-major
  const int strip = blockIdx.x % at.strips;
  const int c = blockIdx.x / at.strips;
  const int n = blockIdx.y;
  const int tid = threadIdx.x;
  const int jmin = at.rows.lo[strip], jcnt = at.rows.cnt[strip];
  constexpr int kB = 8;
  const bool masked = a.has_mask && a.mask_src.eps && a.mask_src.clip;
  const int64_t plane = (int64_t)c * H * W;
  const float* mx = masked ? a.mask_src.x + n * a.mask_src.x_stride + plane : nullptr;
  const float* me = masked ? a.mask_src.eps + n * a.mask_src.eps_stride + plane : nullptr;
  const float* ex = a.extra ? a.extra + n * a.extra_stride + plane : nullptr;
  float xv[kB], ev[kB], xt[kB];
  static_assert(kB == 8 && RA % 8 == 0 && kChunks <= 16, "one mbarrier per 8-row chunk");
  if (kBulk && tid == 0) {
#pragma unroll
    for (int q = 0; q < kChunks; ++q) mbar_init(&bars[q], 1);
    mbar_init_fence();
    if (masked) {
      for (int q = 0; q < kChunks; ++q) {
        const int row0 = strip * RA + q * 8;
        const int rows = min(8, H - row0);
        if (rows <= 0) break;
        const unsigned bytes = (unsigned)(rows * W * sizeof(float));
        mbar_expect_tx(&bars[q], 2 * bytes);
        bulk_load(xs + q * 8 * W, mx + (int64_t)row0 * W, bytes, &bars[q]);
        bulk_load(xs + (RA + q * 8) * W, me + (int64_t)row0 * W, bytes, &bars[q]);
      }
    }
  }
  auto load_batch = [&](int m, int i0) {
    if (kBulk) {
      if (masked && strip * RA + i0 < H) mbar_wait(&bars[i0 / 8], 0);
#pragma unroll
      for (int b = 0; b < kB; ++b) {
        xv[b] = masked ? xs[(i0 + b) * W + m] : 0.f;  // rows past H hold stale shared memory; they are never stored
        ev[b] = masked ? xs[(RA + i0 + b) * W + m] : 0.f;
        xt[b] = ex ? ldg_stream(ex + (int64_t)min(strip * RA + i0 + b, H - 1) * W + m) : 0.f;
      }
      return;
    }
#pragma unroll
    for (int b = 0; b < kB; ++b) {
      const int row = min(strip * RA + i0 + b, H - 1);
      const int64_t off = (int64_t)row * W + m;
      xv[b] = masked ? ldg_stream(mx + off) : 0.f;
      ev[b] = masked ? ldg_stream(me + off) : 0.f;
      xt[b] = ex ? ldg_stream(ex + off) : 0.f;
    }
  };
  if (!kBulk && tid < W) load_batch(tid, 0);  // in flight while the tiles are staged
  stage_async(G, a.r + ((int64_t)n * C + c) * oH * oW + (int64_t)jmin * oW, jcnt * oW, tid, kThreads);
  stage_async(dht, at.dht + (int64_t)strip * RA * KJ, RA * KJ, tid, kThreads);
  stage_async(wts, ac.wtt, ac.kt * W, tid, kThreads);
  const float coef = a.coef ? a.coef[n] : 1.0f;
  int js_next = tid < W ? ac.jstart[tid] : 0;
  stage_wait();
  __syncthreads();

  for (int m = tid; m < W; m += kThreads) {
    // E[jj] = Σ_k A_w[jstart+k][m]·G[jmin+jj][jstart+k]
    const int js = js_next;
    if (m + kThreads < W) js_next = ac.jstart[m + kThreads];
    float wt[KT];
#pragma unroll
    for (int k = 0; k < KT; ++k) wt[k] = k < ac.kt ? wts[k * W + m] : 0.f;
    float e[KJ];
#pragma unroll
    for (int jj = 0; jj < KJ; ++jj) {
      float s = 0.f;
      if (jj < jcnt) {
        const float* gr = G + jj * oW + js;
#pragma unroll
        for (int k = 0; k < KT; ++k)
          if (k < ac.kt && js + k < oW) s = fmaf(wt[k], gr[k], s);
      }
      e[jj] = s;
    }
    // out[i][m] = Σ_jj A_h[jmin+jj][i]·E[jj]; rows in batches of 8, the clamp-mask sources (x, ε) and `extra` of a
    // batch are loaded before any use
#pragma unroll 1
    for (int i0 = 0; i0 < RA; i0 += kB) {
      if (kBulk || i0 != 0 || m != tid) load_batch(m, i0);
#pragma unroll
      for (int b = 0; b < kB; ++b) {
        const int row = strip * RA + i0 + b;
        if (row < H) {
          const float4* dr = reinterpret_cast<const float4*>(dht + (i0 + b) * KJ);
          float s = 0.f;
#pragma unroll
          for (int q = 0; q < KJ / 4; ++q) {
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

// Lean short-strip adjoint — OPT-IN (DPSTTC_RESIZE_ADJ_LEAN=1), the adjoint's counterpart of resize_fwd_lean_kernel, prepared
// from the same N = 8 ncu capture (profiles/r1l_resize_n8_sass.md): resize_adj_kernel<8, 8, 4, true> executes ≈780 instructions
// per thread for 8 outputs, 12 % of them FFMA, at 60 % issue-slot utilisation — guards (`row < H`, `masked`, `extra`, `jj < jcnt`,
// `k < kt`), the generic column loop and runtime shapes make up the rest.  This variant is the guided step's case only — clamp
// mask on, no `extra`, W = 256 = one column per thread, H a multiple of 8, oW = 64, at most 4 column taps — with the guards
// turned into selects on the operand (fma(w, 0, s) = s exactly: w is finite and s is never −0) or removed where the shape
// makes them constant.  Same fma chains in the same order as resize_adj_kernel, hence the same bits; gate:
//   python tools/variant_check.py --op sr4 --n 8 --env DPSTTC_RESIZE_ADJ_LEAN=0 --env DPSTTC_RESIZE_ADJ_LEAN=1
__global__ void __launch_bounds__(kThreads, 4) resize_adj_lean_kernel(const AdjStrips at, const AdjCols ac, int C, int H, int oH,
                                                                      const AdjArgs a) {
  constexpr int RA = kRAs, KJ = kJs, KT = 4, W = 256, oW = 64;
  static_assert(kThreads == W && RA == 8 && KJ == 8, "one column per thread, one 8-row chunk per CTA");
  extern __shared__ __align__(128) float smem[];
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem);  // same carve-up as resize_adj_kernel<…, kBulk = true>
  float* xs = smem + 32;                               // (RA, W) x rows, then (RA, W) ε rows
  float* G = xs + 2 * RA * W;                          // (KJ, oW)
  float* dht = G + ((KJ * oW + 3) & ~3);               // (RA, KJ)
  float* wts = dht + RA * KJ;                          // (kt, W)
  const int strip = blockIdx.x % at.strips;
  const int c = blockIdx.x / at.strips;
  const int n = blockIdx.y;
  const int m = threadIdx.x;
  const int jmin = at.rows.lo[strip], jcnt = at.rows.cnt[strip];
  const int64_t plane = (int64_t)c * H * W;
  const int64_t row0 = (int64_t)strip * RA * W;
  if (m == 0) {
    mbar_init(bar, 1);
    mbar_init_fence();
    constexpr unsigned bytes = RA * W * sizeof(float);
    mbar_expect_tx(bar, 2 * bytes);
    bulk_load(xs, a.mask_src.x + n * a.mask_src.x_stride + plane + row0, bytes, bar);
    bulk_load(xs + RA * W, a.mask_src.eps + n * a.mask_src.eps_stride + plane + row0, bytes, bar);
  }
  stage_async(G, a.r + ((int64_t)n * C + c) * oH * oW + (int64_t)jmin * oW, jcnt * oW, m, kThreads);
  stage_async(dht, at.dht + (int64_t)strip * RA * KJ, RA * KJ, m, kThreads);
  stage_async(wts, ac.wtt, ac.kt * W, m, kThreads);
  const float coef = a.coef ? a.coef[n] : 1.0f;
  const int js = ac.jstart[m];
  stage_wait();
  __syncthreads();

  // E[jj] = Σ_k A_w[jstart+k][m]·G[jmin+jj][jstart+k]
  float wt[KT];
#pragma unroll
  for (int k = 0; k < KT; ++k) wt[k] = k < ac.kt ? wts[k * W + m] : 0.f;
  float e[KJ];
#pragma unroll
  for (int jj = 0; jj < KJ; ++jj) {
    const float* gr = G + jj * oW + js;
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < KT; ++k) {
      const bool on = jj < jcnt && k < ac.kt && js + k < oW;
      s = fmaf(wt[k], on ? gr[k] : 0.f, s);  // off: s + (±0) = s; the unstaged value never reaches the fma
    }
    e[jj] = s;
  }
  // out[i][m] = Σ_jj A_h[jmin+jj][i]·E[jj], clamp mask from the bulk-copied x / ε rows
  mbar_wait(bar, 0);
  const float c1 = a.mask_src.c1, c2 = a.mask_src.c2;
  float* gp = a.g + n * a.g_stride + plane + row0 + m;
#pragma unroll
  for (int b = 0; b < RA; ++b) {
    const float4* dr = reinterpret_cast<const float4*>(dht + b * KJ);
    float s = 0.f;
#pragma unroll
    for (int q = 0; q < KJ / 4; ++q) {
      const float4 w = dr[q];
      s = fmaf(w.x, e[4 * q + 0], s); s = fmaf(w.y, e[4 * q + 1], s);
      s = fmaf(w.z, e[4 * q + 2], s); s = fmaf(w.w, e[4 * q + 3], s);
    }
    float res = coef * s + 0.f;  // resize_adj_kernel: coef·s + extra with extra = 0
    res *= clamp_pass(x0_pre(xs[b * W + m], xs[(RA + b) * W + m], c1, c2));
    stg_stream(gp + b * W, res);
  }
}

// Streaming adjoint for full machines (W = 256, H a multiple of 32, clamp mask on): persistent CTAs, one producer warp
// and 256 consumer threads.  The producer walks the CTA's work list — units (particle, channel, 32-row strip), round
// robin over the grid — and keeps a ring of kSaStages 16-row chunks of the clamp-mask sources (x rows | ε rows, 32 KB per
// stage) plus the NEXT unit's header (its measurement rows and its A_hᵀ block) in flight through the TMA engine, gated by
// full/empty mbarriers, so loads never stop for a CTA's compute or store phase.  Consumers: 128 column pairs × 2 row
// halves; E = G·A_w in registers per unit, then 12-16 packed FFMA2 per output row pair element, the clamp mask from the
// ring, one 64-bit streaming store per row.  Arithmetic (order of every fma chain) is that of resize_adj_kernel:
// the result is bit-identical to the strip kernels', whichever variant a launch picks.
template <int KJ, int KT>
__global__ void __launch_bounds__(kSaThreads, 2) resize_adj_stream_kernel(const AdjStrips at, const AdjCols ac, int C, int H,
                                                                            int oH, int oW, int units, const AdjArgs a) {
  constexpr int W = 256, W2 = 128, RA = kRA, kChunks = RA / kSaRows, kHalfRows = kSaRows / 2;
  constexpr int kStageFloats = 2 * kSaRows * W;
  extern __shared__ __align__(128) float smem[];
  uint64_t* full = reinterpret_cast<uint64_t*>(smem);  // [kSaStages]
  uint64_t* empty = full + kSaStages;                  // [kSaStages]
  uint64_t* hfull = empty + kSaStages;                 // [2]
  uint64_t* hempty = hfull + 2;                        // [2]   → 2·kSaStages + 4 ≤ 16 barriers = 128 bytes
  static_assert(2 * kSaStages + 4 <= 16, "barriers must fit in the first 128 bytes");
  float* ring = smem + 32;
  const int hdr_floats = KJ * oW + RA * KJ;  // [G: (KJ, oW) | dht: (RA, KJ)], both multiples of 4 floats
  float* hdr = ring + kSaStages * kStageFloats;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < kSaStages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 8); }
    mbar_init(&hfull[0], 1); mbar_init(&hfull[1], 1);
    mbar_init(&hempty[0], 8); mbar_init(&hempty[1], 8);
    mbar_init_fence();
  }
  __syncthreads();
  const int strips = at.strips;

  if (warp == 8) {  // ---------------- producer ----------------
    if (lane != 0) return;
    int it = 0, hu = 0;
    for (int u = blockIdx.x; u < units; u += gridDim.x, ++hu) {
      const int strip = u % strips, c = (u / strips) % C, n = u / (strips * C);
      const int hs = hu & 1;
      const int jmin = at.rows.lo[strip], jcnt = at.rows.cnt[strip];
      mbar_wait_guarded(&hempty[hs], (unsigned)(((hu >> 1) & 1) ^ 1));
      float* hb = hdr + hs * hdr_floats;
      const unsigned bytes_g = (unsigned)(jcnt * oW * sizeof(float)), bytes_d = (unsigned)(RA * KJ * sizeof(float));
      mbar_expect_tx(&hfull[hs], bytes_g + bytes_d);
      bulk_load(hb, a.r + (((int64_t)n * C + c) * oH + jmin) * oW, bytes_g, &hfull[hs]);
      bulk_load(hb + KJ * oW, at.dht + (int64_t)strip * RA * KJ, bytes_d, &hfull[hs]);
      const int64_t base = (int64_t)c * H * W + (int64_t)strip * RA * W;
      const float* mx = a.mask_src.x + n * a.mask_src.x_stride + base;
      const float* me = a.mask_src.eps + n * a.mask_src.eps_stride + base;
#pragma unroll 1
      for (int q = 0; q < kChunks; ++q, ++it) {
        const int s = it % kSaStages;
        mbar_wait_guarded(&empty[s], (unsigned)(((it / kSaStages) & 1) ^ 1));
        float* dst = ring + s * kStageFloats;
        mbar_expect_tx(&full[s], (unsigned)(kStageFloats * sizeof(float)));
        bulk_load(dst, mx + q * kSaRows * W, kSaRows * W * sizeof(float), &full[s]);
        bulk_load(dst + kSaRows * W, me + q * kSaRows * W, kSaRows * W * sizeof(float), &full[s]);
      }
    }
    return;
  }

  // ---------------- consumers ----------------
  const int half = tid >> 7, cp = tid & 127;
  float w0[KT], w1[KT];
  int i0[KT], i1[KT];
  {
    const int js0 = ac.jstart[2 * cp], js1 = ac.jstart[2 * cp + 1];
#pragma unroll
    for (int k = 0; k < KT; ++k) {
      const bool v0 = k < ac.kt && js0 + k < oW, v1 = k < ac.kt && js1 + k < oW;
      w0[k] = v0 ? ac.wtt[k * W + 2 * cp] : 0.f;
      w1[k] = v1 ? ac.wtt[k * W + 2 * cp + 1] : 0.f;
      i0[k] = min(js0 + k, oW - 1);
      i1[k] = min(js1 + k, oW - 1);
    }
  }
  const float c1 = a.mask_src.c1, c2 = a.mask_src.c2;
  int it = 0, hu = 0;
  for (int u = blockIdx.x; u < units; u += gridDim.x, ++hu) {
    const int strip = u % strips, c = (u / strips) % C, n = u / (strips * C);
    const int hs = hu & 1;
    const int jcnt = at.rows.cnt[strip];
    const float coef = a.coef ? a.coef[n] : 1.0f;
    const int64_t plane = (int64_t)c * H * W;
    const float* ex = a.extra ? a.extra + n * a.extra_stride + plane : nullptr;
    float* gout = a.g + n * a.g_stride + plane;
    mbar_wait_guarded(&hfull[hs], (unsigned)((hu >> 1) & 1));
    const float* G = hdr + hs * hdr_floats;
    const float* D = G + KJ * oW;
    // E[jj] = Σ_k A_w[jstart+k][m]·G[jmin+jj][jstart+k] for the thread's two columns
    float2 e[KJ];
#pragma unroll
    for (int jj = 0; jj < KJ; ++jj) {
      float s0 = 0.f, s1 = 0.f;
      if (jj < jcnt) {
        const float* gr = G + jj * oW;
#pragma unroll
        for (int k = 0; k < KT; ++k) {
          s0 = fmaf(w0[k], gr[i0[k]], s0);
          s1 = fmaf(w1[k], gr[i1[k]], s1);
        }
      }
      e[jj] = make_float2(s0, s1);
    }
    // per-unit bases: every per-row offset below is a compile-time immediate
    const int64_t off0 = (int64_t)(strip * RA + half * kHalfRows) * W + 2 * cp;
    float* gp = gout + off0;
    const float* exp_ = ex ? ex + off0 : nullptr;
    const float4* dbase = reinterpret_cast<const float4*>(D + half * kHalfRows * KJ);
#pragma unroll
    for (int q = 0; q < kChunks; ++q, ++it) {
      const int s = it % kSaStages;
      mbar_wait_guarded(&full[s], (unsigned)((it / kSaStages) & 1));
      const float2* xs = reinterpret_cast<const float2*>(ring + s * kStageFloats) + half * kHalfRows * W2 + cp;
      const float2* es = xs + kSaRows * W2;
#pragma unroll
      for (int rr = 0; rr < kHalfRows; ++rr) {
        const int i = q * kSaRows + rr;  // row inside the strip, relative to the half's first row
        const float4* dr = dbase + i * (KJ / 4);
        float2 acc = make_float2(0.f, 0.f);
#pragma unroll
        for (int qq = 0; qq < KJ / 4; ++qq) {
          const float4 w = dr[qq];
          acc = __ffma2_rn(make_float2(w.x, w.x), e[4 * qq + 0], acc);
          acc = __ffma2_rn(make_float2(w.y, w.y), e[4 * qq + 1], acc);
          acc = __ffma2_rn(make_float2(w.z, w.z), e[4 * qq + 2], acc);
          acc = __ffma2_rn(make_float2(w.w, w.w), e[4 * qq + 3], acc);
        }
        float2 xt = make_float2(0.f, 0.f);
        if (exp_) xt = ldg_stream2(reinterpret_cast<const float2*>(exp_ + i * W));
        float2 res = __ffma2_rn(make_float2(coef, coef), acc, xt);
        const float2 pre = x0_pair_pre(xs[rr * W2], es[rr * W2], c1, c2);
        res.x *= clamp_pass(pre.x);
        res.y *= clamp_pass(pre.y);
        stg_stream2(gp + i * W, res);
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty[s]);
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(&hempty[hs]);
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

size_t fwd_smem(const FwdTables& t, int parts, int W, int oW) {
  return sizeof(float) * ((size_t)t.span * kRO + (size_t)parts * kRO * W + 64 + (size_t)((t.kw * oW + 3) & ~3) + oW);
}
size_t adj_smem(int ra, int kj, int oW, int kt, int W) {
  return sizeof(float) * ((size_t)((kj * oW + 3) & ~3) + (size_t)ra * kj + (size_t)kt * W);
}

// strip tables of A_hᵀ for strips of `ra` input rows; DPS_ERR_UNSUPPORTED when a strip touches more than `kj` rows
int build_adj_strips(const std::vector<double>& Ah, int H, int out_h, int ra, int kj, AdjStrips* out) {
  const int strips = (H + ra - 1) / ra;
  if (strips > kMaxStrips) return DPS_ERR_UNSUPPORTED;
  std::vector<int> jmin(strips), jcnt(strips);
  for (int s = 0; s < strips; ++s) {
    int lo = out_h, hi = -1;
    for (int i = s * ra; i < std::min(H, (s + 1) * ra); ++i)
      for (int j = 0; j < out_h; ++j)
        if (Ah[(size_t)j * H + i] != 0.0) { lo = std::min(lo, j); hi = std::max(hi, j); }
    if (hi < lo) { lo = 0; hi = 0; }
    jmin[s] = lo;
    jcnt[s] = hi - lo + 1;
    if (jcnt[s] > kj) return DPS_ERR_UNSUPPORTED;
  }
  std::vector<float> dht((size_t)strips * ra * kj, 0.f);
  for (int s = 0; s < strips; ++s)
    for (int ii = 0; ii < ra; ++ii)
      for (int jj = 0; jj < jcnt[s]; ++jj) {
        const int i = s * ra + ii;
        if (i < H) dht[((size_t)s * ra + ii) * kj + jj] = (float)Ah[(size_t)(jmin[s] + jj) * H + i];
      }
  for (int s = 0; s < strips; ++s) {
    out->rows.lo[s] = (short)jmin[s];
    out->rows.cnt[s] = (short)jcnt[s];
  }
  if (int rc = upload(dht, &out->dht)) return rc;
  out->strips = strips;
  return DPS_OK;
}

}  // namespace

int resize_create(dps_operator* op, const int32_t* fov_h, const float* w_h, int taps_h, int out_h,
                  const int32_t* fov_w, const float* w_w, int taps_w, int out_w) {
  const int H = op->H, W = op->W;
  DPS_REQUIRE(out_h > 0 && out_w > 0 && out_h <= H && out_w <= W, DPS_ERR_UNSUPPORTED,
              "resize: only down-scaling is supported (%dx%d -> %dx%d)", H, W, out_h, out_w);
  DPS_REQUIRE(H < 32768 && (out_h + kRO - 1) / kRO <= kMaxStrips, DPS_ERR_UNSUPPORTED,
              "resize: at most %d measurement rows are supported (got %d)", kMaxStrips * kRO, out_h);
  bool ok_h, ok_w;
  std::vector<double> Ah = dense_from_tables(fov_h, w_h, taps_h, out_h, H, &ok_h);
  std::vector<double> Aw = dense_from_tables(fov_w, w_w, taps_w, out_w, W, &ok_w);
  DPS_REQUIRE(ok_h && ok_w, DPS_ERR_INVALID, "resize: field-of-view index out of range");
  ResizeTables* t = new ResizeTables();
  op->resize = t;
  FwdTables& f = t->f;
  // ---- forward H-pass blocks ----
  f.fstrips = (out_h + kRO - 1) / kRO;
  int span = 1;
  for (int s = 0; s < f.fstrips; ++s) {
    int lo = H, hi = -1;
    for (int j = s * kRO; j < std::min(out_h, (s + 1) * kRO); ++j)
      for (int m = 0; m < H; ++m)
        if (Ah[(size_t)j * H + m] != 0.0) { lo = std::min(lo, m); hi = std::max(hi, m); }
    if (hi < lo) { lo = 0; hi = 0; }
    // windows start on absolute multiples of 8 rows: the bulk kernel's chunks (and its split of a chunk's rows into two
    // partial sums) then coincide with the streaming kernel's, which makes the two bit-identical
    if (H % 8 == 0) lo &= ~7;
    f.rows.lo[s] = (short)lo;
    f.rows.cnt[s] = (short)(hi - lo + 1);
    span = std::max(span, hi - lo + 1);
  }
  f.span = span;
  std::vector<float> dh((size_t)f.fstrips * span * kRO, 0.f);
  for (int s = 0; s < f.fstrips; ++s)
    for (int rr = 0; rr < f.rows.cnt[s]; ++rr)
      for (int j = 0; j < kRO; ++j) {
        const int jo = s * kRO + j;
        if (jo < out_h) dh[((size_t)s * span + rr) * kRO + j] = (float)Ah[(size_t)jo * H + f.rows.lo[s] + rr];
      }
  // ---- forward W-pass windows (tap-major so that a warp reads consecutive words) ----
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
  f.kw = kw;
  std::vector<float> wwt((size_t)kw * out_w, 0.f);
  for (int j = 0; j < out_w; ++j)
    for (int k = 0; k < kw; ++k)
      if (cstart[j] + k < W) wwt[(size_t)k * out_w + j] = (float)Aw[(size_t)j * W + cstart[j] + k];
  // ---- streaming forward: sliding-window tables (×4 / ×8 style operators on 256-wide images only) ----
  {
    FwdStream& fs = t->fs;
    const int nchunks = H / 8;
    const bool shape_ok = W == 256 && H % 8 == 0 && out_h % kRO == 0 && (H == 4 * out_h || H == 8 * out_h) &&
                          (f.fstrips + kStripsPerUnit - 1) / kStripsPerUnit <= 8;
    if (shape_ok) {
      const int D = 8 * out_h / H;
      int B = -(1 << 30);
      std::vector<int> first(out_h, nchunks), last(out_h, -1);
      for (int j = 0; j < out_h; ++j)
        for (int m = 0; m < H; ++m)
          if (Ah[(size_t)j * H + m] != 0.0) {
            const int Q = m / 8;
            first[j] = std::min(first[j], Q);
            last[j] = std::max(last[j], Q);
            B = std::max(B, D * Q - j);  // jb(Q) = D·Q − B ≤ j for every contribution
          }
      bool ok = B > -(1 << 30);
      for (int j = 0; ok && j < out_h; ++j) {
        if (last[j] < 0) continue;
        if (j >= D * first[j] - B + kWO) ok = false;  // inside the window of its first chunk (later chunks: larger base)
        if (last[j] > (j + B) / D) ok = false;        // complete when it is emitted
        if (j + B < 0) ok = false;
      }
      if (ok) {
        std::vector<float> wq((size_t)H * kWO, 0.f);
        for (int m = 0; m < H; ++m)
          for (int tt = 0; tt < kWO; ++tt) {
            const int j = D * (m / 8) - B + tt;
            if (j >= 0 && j < out_h) wq[(size_t)m * kWO + tt] = (float)Ah[(size_t)j * H + m];
          }
        fs.upp = (f.fstrips + kStripsPerUnit - 1) / kStripsPerUnit;
        for (int k = 0; k < fs.upp; ++k) {
          const int j_lo = k * kStripsPerUnit * kRO, j_hi = std::min(out_h, j_lo + kStripsPerUnit * kRO);
          int q_lo = nchunks;
          for (int j = j_lo; j < j_hi; ++j) q_lo = std::min(q_lo, first[j]);
          // the window base of the first chunk must not lie beyond the unit's first row
          q_lo = std::min(q_lo, (j_lo + B) / D);
          fs.q_lo[k] = (short)std::max(0, q_lo);
          fs.q_hi[k] = (short)((j_hi - 1 + B) / D);
        }
        if (int rc = upload(wq, &fs.wq)) return rc;
        fs.D = D;
        fs.B = B;
        fs.ok = 1;
      }
    }
  }
  // ---- adjoint: A_hᵀ blocks per input-row strip, two strip heights ----
  DPS_REQUIRE(build_adj_strips(Ah, H, out_h, kRA, kJMax, &t->big) == DPS_OK, DPS_ERR_UNSUPPORTED,
              "resize: a %d-row strip touches more than %d measurement rows", kRA, kJMax);
  if (build_adj_strips(Ah, H, out_h, kRAs, kJs, &t->small) != DPS_OK) t->small.strips = 0;  // optional variant
  // ---- adjoint: A_w column windows (tap-major) ----
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
  t->cols.kt = kt;
  std::vector<float> wtt((size_t)kt * W, 0.f);
  for (int m = 0; m < W; ++m)
    for (int k = 0; k < kt; ++k)
      if (jstart[m] + k < out_w) wtt[(size_t)k * W + m] = (float)Aw[(size_t)(jstart[m] + k) * W + m];

  DPS_REQUIRE(fwd_smem(f, 2, W, out_w) <= 227 * 1024 && adj_smem(kRA, kJMax, out_w, kt, W) <= 227 * 1024,
              DPS_ERR_UNSUPPORTED, "resize: tiles exceed shared memory");
  if (int rc = upload(dh, &f.dh)) return rc;
  if (int rc = upload(cstart, &f.cstart)) return rc;
  if (int rc = upload(wwt, &f.wwt)) return rc;
  if (int rc = upload(jstart, &t->cols.jstart)) return rc;
  if (int rc = upload(wtt, &t->cols.wtt)) return rc;
  op->oC = op->C;
  op->oH = out_h;
  op->oW = out_w;
  op->P = op->C * f.fstrips;
  op->taps = taps_h;
  return resize_fused_create(op, Ah, Aw, out_h, out_w);   // fused residual + cotangent kernel where the shape allows
}

void resize_destroy(dps_operator* op) {
  ResizeTables* t = op->resize;
  if (!t) return;
  cudaFree(t->f.dh); cudaFree(t->f.cstart); cudaFree(t->f.wwt); cudaFree(t->fs.wq);
  cudaFree(t->big.dht); cudaFree(t->small.dht);
  cudaFree(t->cols.jstart); cudaFree(t->cols.wtt);
  delete t;
  op->resize = nullptr;
}

// The short-strip adjoint (4x the CTAs, one 8-row chunk each) wins while the machine is not yet full of long strips —
// measured with the bulk-copy kernels, 256² → 64²: 7.7 vs 10.1 µs at N = 8, 11.9 vs 15.0 at N = 16, 21.4 vs 25.2 at
// N = 32, but 80.8 vs 75.5 µs at N = 128.
static int variant_override() {  // DPSTTC_RESIZE_VARIANT=big|small|stream pins the choice (profiling / test aid)
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("DPSTTC_RESIZE_VARIANT");
    v = !e ? 0 : (e[0] == 'b' ? 1 : (e[0] == 's' ? (e[1] == 't' ? 3 : 2) : 0));  // "stream" = 3
  }
  return v;
}
static bool small_grid(int64_t ctas) {
  const int v = variant_override();
  return v ? v == 2 : ctas <= 148 * 10;
}

// Deep ring (6 stages, 104 KB) only while the strip grid leaves a third of the SMs empty: measured ×4, deep vs 3 stages
// (profiles/r1k_ring.md): 6.40 vs 6.66 µs at N = 4 (96 CTAs), but 8.53 vs 8.31 µs at N = 8 (192) and 9.77 vs 9.14 µs at
// N = 12 — once every SM hosts a CTA the launch is paced by its 13-19 MB crossing HBM, not by a CTA's second round trip.
// DPSTTC_RESIZE_FWD_STAGES=3|6 pins the choice (A/B and tests/test_gpu_variants.py).
static bool fwd_deep_ring(int64_t ctas, int device) {
  static int pin = -1;
  if (pin < 0) {
    const char* e = getenv("DPSTTC_RESIZE_FWD_STAGES");
    pin = !e ? 0 : (e[0] == '6' ? 6 : 3);
  }
  if (pin) return pin == 6;
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
  return 3 * ctas <= 2 * (int64_t)sms;
}

// The lean strip forward / lean W pass are the default since round 2 (gate: bit-identical to the round-1 kernels at N = 3, 8,
// 12 — profiles/r2a_lean_gate_*.log — and 8.73 → 7.43 µs at N = 8, 19.9 → 17.9 at 32, 50.5 → 48.6 at 128).
// DPSTTC_RESIZE_FWD_LEAN=0 selects the round-1 kernels again (A/B and tests/test_gpu_variants.py).
static bool fwd_lean() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("DPSTTC_RESIZE_FWD_LEAN");
    v = (e && e[0] == '0') ? 0 : 1;
  }
  return v == 1;
}

int resize_forward(const dps_operator* op, const FwdArgs& a, cudaStream_t st) {
  const FwdTables& f = op->resize->f;
  DPS_SMEM_OPTIN((resize_fwd_kernel), 227 * 1024, op->device);
  DPS_SMEM_OPTIN((resize_fwd_bulk_kernel<kStagesStd>), 227 * 1024, op->device);
  DPS_SMEM_OPTIN((resize_fwd_bulk_kernel<kStagesDeep>), 227 * 1024, op->device);
  {  // streaming variant once the strip grid would fill the machine several times over
    const FwdStream& fs = op->resize->fs;
    const int64_t units = (int64_t)op->C * fs.upp * a.n;
    const int v = variant_override();
    const size_t smem = sizeof(float) * (32 + (size_t)kSfStages * 2 * 8 * 256 + 2 * kRO * 256 + (size_t)op->H * kWO + 128 +
                                         (size_t)((f.kw * op->oW + 3) & ~3) + op->oW);
    if (fs.ok && smem <= 113 * 1024 && units < (1 << 30) && (v == 3 || (v == 0 && (int64_t)op->C * f.fstrips * a.n >= 80 * 24))) {  // N ≥ 80: 41.1 vs 44.5 µs at 96, 50.7 vs 56.6 at 128
      DPS_SMEM_OPTIN((resize_fwd_stream_kernel<1>), 227 * 1024, op->device);
      DPS_SMEM_OPTIN((resize_fwd_stream_kernel<2>), 227 * 1024, op->device);
      int sms = 148;
      cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, op->device);
      const int rounds = (int)((units + 2 * sms - 1) / (2 * sms));
      const int grid = (int)((units + rounds - 1) / rounds);  // every CTA gets `rounds` units (±1): no ragged tail
      if (fwd_lean() && op->oW == 64 && f.kw == 16) {  // opt-in: lean W pass (wpass_one_lean)
        DPS_SMEM_OPTIN((resize_fwd_stream_kernel<1, true>), 227 * 1024, op->device);
        DPS_SMEM_OPTIN((resize_fwd_stream_kernel<2, true>), 227 * 1024, op->device);
        if (fs.D == 2)
          resize_fwd_stream_kernel<2, true><<<grid, kSaThreads, smem, st>>>(f, fs, op->C, op->H, op->oH, op->oW, (int)units, a);
        else
          resize_fwd_stream_kernel<1, true><<<grid, kSaThreads, smem, st>>>(f, fs, op->C, op->H, op->oH, op->oW, (int)units, a);
      } else if (fs.D == 2)
        resize_fwd_stream_kernel<2><<<grid, kSaThreads, smem, st>>>(f, fs, op->C, op->H, op->oH, op->oW, (int)units, a);
      else
        resize_fwd_stream_kernel<1><<<grid, kSaThreads, smem, st>>>(f, fs, op->C, op->H, op->oH, op->oW, (int)units, a);
      DPS_LAUNCH_CHECK("resize_forward");
      return DPS_OK;
    }
  }
  dim3 grid((unsigned)(op->C * f.fstrips), (unsigned)a.n);
  if (op->W == 256) {
    if (fwd_lean() && a.src.eps && a.src.clip && op->oW == 64 && f.kw == 16 && a.n > 0) {
      DPS_SMEM_OPTIN((resize_fwd_lean_kernel<kStagesStd, 16>), 227 * 1024, op->device);
      DPS_SMEM_OPTIN((resize_fwd_lean_kernel<kStagesDeep, 16>), 227 * 1024, op->device);
      if (fwd_deep_ring((int64_t)grid.x * grid.y, op->device))
        resize_fwd_lean_kernel<kStagesDeep, 16><<<grid, 256, fwd_smem(f, 0, 256, op->oW) + sizeof(float) * (32 + kStagesDeep * 2 * kCR * 256), st>>>(
            f, op->C, op->H, op->oH, a);
      else
        resize_fwd_lean_kernel<kStagesStd, 16><<<grid, 256, fwd_smem(f, 0, 256, op->oW) + sizeof(float) * (32 + kStagesStd * 2 * kCR * 256), st>>>(
            f, op->C, op->H, op->oH, a);
    } else if (fwd_deep_ring((int64_t)grid.x * grid.y, op->device))
      resize_fwd_bulk_kernel<kStagesDeep><<<grid, 256, fwd_smem(f, 0, 256, op->oW) + sizeof(float) * (32 + kStagesDeep * 2 * kCR * 256), st>>>(
          f, op->C, op->H, op->oH, op->oW, a);
    else
      resize_fwd_bulk_kernel<kStagesStd><<<grid, 256, fwd_smem(f, 0, 256, op->oW) + sizeof(float) * (32 + kStagesStd * 2 * kCR * 256), st>>>(
          f, op->C, op->H, op->oH, op->oW, a);
  } else {
    resize_fwd_kernel<<<grid, kThreads, fwd_smem(f, 1, op->W, op->oW), st>>>(f, op->C, op->H, op->W, op->oH, op->oW, a);
  }
  DPS_LAUNCH_CHECK("resize_forward");
  return DPS_OK;
}

template <int RA, int KJ, int KT>
static int launch_adj(const dps_operator* op, const AdjStrips& strips, const AdjArgs& a, cudaStream_t st) {
  const ResizeTables& t = *op->resize;
  DPS_SMEM_OPTIN((resize_adj_kernel<RA, KJ, KT, false>), 227 * 1024, op->device);
  DPS_SMEM_OPTIN((resize_adj_kernel<RA, KJ, KT, true>), 227 * 1024, op->device);
  dim3 grid((unsigned)(op->C * strips.strips), (unsigned)a.n);
  const size_t base = adj_smem(RA, KJ, op->oW, t.cols.kt, op->W);
  const size_t bulk = base + sizeof(float) * (32 + (size_t)2 * RA * op->W);
  // bulk copies need 16-byte rows and a tile that still leaves 3 CTAs per SM
  if (op->W % 4 == 0 && bulk <= 75 * 1024)
    resize_adj_kernel<RA, KJ, KT, true><<<grid, kThreads, bulk, st>>>(strips, t.cols, op->C, op->H, op->W, op->oH,
                                                                      op->oW, a);
  else
    resize_adj_kernel<RA, KJ, KT, false><<<grid, kThreads, base, st>>>(strips, t.cols, op->C, op->H, op->W, op->oH,
                                                                       op->oW, a);
  DPS_LAUNCH_CHECK("resize_adjoint");
  return DPS_OK;
}

// The streaming (persistent, TMA-pipelined) adjoint once the strip grid fills the machine several times over.
template <int KT>
static int launch_adj_stream(const dps_operator* op, const AdjArgs& a, cudaStream_t st) {
  const ResizeTables& t = *op->resize;
  DPS_SMEM_OPTIN((resize_adj_stream_kernel<kJMax, KT>), 227 * 1024, op->device);
  const int units = op->C * t.big.strips * a.n;
  const size_t smem = sizeof(float) * (32 + (size_t)kSaStages * 2 * kSaRows * 256 + 2 * ((size_t)kJMax * op->oW + kRA * kJMax));
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, op->device);
  const int rounds = (units + 2 * sms - 1) / (2 * sms);
  const int grid = (units + rounds - 1) / rounds;  // every CTA gets `rounds` units (±1): no ragged tail
  resize_adj_stream_kernel<kJMax, KT><<<grid, kSaThreads, smem, st>>>(t.big, t.cols, op->C, op->H, op->oH, op->oW, units, a);
  DPS_LAUNCH_CHECK("resize_adjoint");
  return DPS_OK;
}

int resize_adjoint(const dps_operator* op, const AdjArgs& a, cudaStream_t st) {
  const ResizeTables& t = *op->resize;
  const bool narrow = t.cols.kt <= 4;  // KT: compile-time bound of the column window (4 covers bicubic x4 and x8)
  {
    const bool masked = a.has_mask && a.mask_src.eps && a.mask_src.clip;
    const int64_t units = (int64_t)op->C * t.big.strips * a.n;
    const int v = variant_override();
    const bool eligible = masked && op->W == 256 && op->H % kRA == 0 && op->oW % 4 == 0 && units < (1 << 30) &&
                          sizeof(float) * (32 + (size_t)kSaStages * 2 * kSaRows * 256 + 2 * ((size_t)kJMax * op->oW + kRA * kJMax)) <= 113 * 1024;
    // measured (N: small / big / stream µs): 32: 28.3 / 28.4 / 27.1, 48: 34.4 / 36.4 / 27.2, 64: 44.4 / 45.0 / 34.3,
    // 96: 63.2 / 60.4 / 44.1, 128: 81.4 / 76.5 / 56.7
    if (eligible && (v == 3 || (v == 0 && units >= 40 * 24)))
      return narrow ? launch_adj_stream<4>(op, a, st) : launch_adj_stream<kKTMax>(op, a, st);
  }
  if (t.small.strips && small_grid((int64_t)op->C * t.big.strips * a.n)) {
    // default since round 2 (bit-identical gate passed; 8.38 → 6.78 µs at N = 8, 22.3 → 18.0 at N = 32); =0: round-1 kernel
    static const bool lean = !(getenv("DPSTTC_RESIZE_ADJ_LEAN") && getenv("DPSTTC_RESIZE_ADJ_LEAN")[0] == '0');
    const bool masked = a.has_mask && a.mask_src.eps && a.mask_src.clip;
    if (lean && narrow && masked && !a.extra && op->W == 256 && op->H % kRAs == 0 && op->oW == 64 && a.r) {
      DPS_SMEM_OPTIN((resize_adj_lean_kernel), 227 * 1024, op->device);
      dim3 grid((unsigned)(op->C * t.small.strips), (unsigned)a.n);
      const size_t smem = adj_smem(kRAs, kJs, op->oW, t.cols.kt, op->W) + sizeof(float) * (32 + (size_t)2 * kRAs * op->W);
      resize_adj_lean_kernel<<<grid, kThreads, smem, st>>>(t.small, t.cols, op->C, op->H, op->oH, a);
      DPS_LAUNCH_CHECK("resize_adjoint");
      return DPS_OK;
    }
    return narrow ? launch_adj<kRAs, kJs, 4>(op, t.small, a, st) : launch_adj<kRAs, kJs, kKTMax>(op, t.small, a, st);
  }
  return narrow ? launch_adj<kRA, kJMax, 4>(op, t.big, a, st) : launch_adj<kRA, kJMax, kKTMax>(op, t.big, a, st);
}
