// Separable (Gaussian) blur guidance in ONE kernel: residual r = y − A x̂₀, its partial sums and the UNSCALED masked
// cotangent g = 1[|pre| ≤ 1] ⊙ Aᵀ r for A = ReflectionPad2d(k/2) + depthwise correlation with a rank-1 kernel
// (measurements.py:129-149, util/img_utils.py:268-308; condition_methods.py:33-39 through autograd).
//
// Two kernels (blur_separable.cu) read x, ε twice and bounce r through HBM: 5T + 2M bytes with a halo re-read of the
// forward pass on top.  With the guidance coefficient deferred to the update kernel (dps_update_ext) nothing global stands
// between A and Aᵀ, so a thread-block CLUSTER of 8 CTAs keeps a whole (particle, channel) plane on chip: 3T + M bytes.
//
// CTA q of the cluster owns image rows [32q, 32q+32).  Passes (R = tap radius, TAPS = 2R+1):
//   0. x, ε rows → shared memory by TMA bulk copies (four 8-row chunks); x̂₀ in place; clamp mask of a column = one register.
//   1. vertical forward  (thread = column): 8 outputs per register block, halo rows from the neighbours' shared memory
//      (DSMEM) or, at the image border, own rows mirrored (reflect: −m ↦ m).  Result → column-padded tile T.
//   2. horizontal forward (thread = 4 columns × 8 rows, 128-bit shared loads): r = y − (·) stays in REGISTERS; Σr², Σ|r|.
//   3. horizontal adjoint: r → zero-padded tile (the T buffer), plain flipped-tap correlation + the folded border terms
//      (Aᵀ = Pᵀ Cᵀ: what the mirrored padding read twice comes back twice), result s → the x̂₀ buffer (dead by then).
//   4. vertical adjoint (thread = column) with halo rows of s over DSMEM, border folds, clamp mask, one store per row.
// Algebra of the adjoint in 1-D (forward out[i] = Σ_k w[k]·xp[i − R + k], xp[−j] = x[j], xp[n−1+j] = x[n−1−j]):
//   g[m] = Σ_i w[m − i + R]·u[i]  +  [1 ≤ m ≤ R] Σ_{i=0}^{R−m} w[R − m − i]·u[i]
//                                 +  [m = n−1−a, 1 ≤ a ≤ R] Σ_{b=0}^{R−a} w[R + a + b]·u[n−1−b].
#include <cooperative_groups.h>

#include <vector>

#include "operator.cuh"

namespace cg = cooperative_groups;

struct SepFused {
  int R = 0;           // template radius
  float wv[33] = {};   // vertical taps, zero-padded to 2R+1 (centre at R)
  float wh[33] = {};
};

namespace {
constexpr int kW = 256, kRI = 32, kCluster = 8, kT = 256, kChunks = 4, kChunkRows = kRI / kChunks, kG = 16, kNB = 18, kSL = 16;
constexpr int kRowB = kW * (int)sizeof(float);

struct SepFusedArgs {
  float wv[33];
  float wh[33];
  int C;
  dps_source src;
  const float* y;
  int64_t y_stride;
  float* r_out;
  float* g;
  int64_t g_stride;
  float* partials;
};

// smallest pitch ≥ need with pitch ≡ 2 (mod 4): lanes that walk ROWS of such a tile (one lane per row / row pair) hit
// distinct bank groups with 64-bit accesses of 4-byte elements and with 128-bit accesses of 8-byte elements
constexpr int pitch2mod4(int need) { return need + ((2 - need % 4) + 4) % 4; }
constexpr int cmax(int a, int b) { return a > b ? a : b; }

template <int R>
struct SepCfg {
  static constexpr int TAPS = 2 * R + 1;
  static constexpr int TP = R, TPITCH = pitch2mod4(kW + 2 * R);              // T2: (16 row pairs, TPITCH) float2, image column c at TP + c
  static constexpr int ZP = kSL + R, ZPITCH = pitch2mod4(ZP + kW + 2 * R);   // Z2: zero-padded residual, same interleaved layout
  static constexpr int SP = R, SPITCH = pitch2mod4(kW + 2 * R);              // s: two (16, SPITCH) float tiles (even rows, odd rows)
  static constexpr int STILE = (kRI / 2) * SPITCH;                           // floats per parity tile
  static constexpr int A_FLOATS = cmax(kRI * kW, 2 * STILE);
  static constexpr int B_FLOATS = cmax(cmax(kRI * kW, (kRI / 2) * TPITCH * 2), (kRI / 2) * ZPITCH * 2 + 16);
  static constexpr size_t SMEM = sizeof(float) * (size_t)(A_FLOATS + B_FLOATS + 64) + 8 * kChunks;
  static constexpr int CTAS = SMEM + 1024 <= 233472 / 3 ? 3 : 2;
};

#ifdef DPS_SEPF_TRACE  // experiment builds only (tools/build_variant.sh): per-CTA phase timestamps of the first 4096 CTAs
__device__ long long sepf_trace[4096 * 16];
#define SEPF_T(i) do { if (threadIdx.x == 0 && blockIdx.x < 4096) sepf_trace[blockIdx.x * 16 + (i)] = clock64(); } while (0)
#else
#define SEPF_T(i) do { } while (0)
#endif

// Cluster barrier halves (arrive early, wait late).  What an arrive has to guarantee, and how (A/B switches for experiment
// builds; see the stress test tools/fused_stress.py):
//   publish   — my CTA's shared-memory stores are visible to the neighbours that read them over DSMEM after their wait:
//               DPS_SEPF_PUB 0: barrier.cluster.arrive.release by every thread (ptxas: MEMBAR.ALL.GPU per warp, ≈2.5 k cycles);
//                            1: bar.sync, ONE warp executes fence.acq_rel.cluster (cumulative), every thread arrives relaxed.
//   done      — my loads from the neighbours' tiles have completed (they may overwrite the tile, or exit):
//               DPS_SEPF_DONE 0: arrive.release by every thread;
//                             1: relaxed arrive predicated on a value that depends on every halo load (acc[0] and acc[15] of the
//                                window cover all 2R halo rows): ptxas cannot schedule it above the FMAs that consume the loads.
#ifndef DPS_SEPF_PUB
#define DPS_SEPF_PUB 1
#endif
#ifndef DPS_SEPF_DONE
#define DPS_SEPF_DONE 1
#endif
DPS_DEV void cl_arrive_release() { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); }
DPS_DEV void cl_arrive_relaxed() { asm volatile("barrier.cluster.arrive.relaxed.aligned;" ::: "memory"); }
// `synced`: a bar.sync after the last store has already been executed by the caller
DPS_DEV void cl_arrive_publish(bool synced) {
#if DPS_SEPF_PUB == 0
  cl_arrive_release();
#else
  if (!synced) __syncthreads();
  if (threadIdx.x < 32) asm volatile("fence.acq_rel.cluster;" ::: "memory");
  cl_arrive_relaxed();
#endif
}
DPS_DEV void cl_arrive_done_reading(float dep0, float dep1) {
#if DPS_SEPF_DONE == 0
  cl_arrive_release();
#else
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .f32 t;\n\tadd.rn.f32 t, %0, %1;\n\tsetp.neu.f32 p, t, t;\n\t"
      "@p barrier.cluster.arrive.relaxed;\n\t@!p barrier.cluster.arrive.relaxed;\n\t}" ::"f"(dep0), "f"(dep1)
      : "memory");
#endif
}
DPS_DEV void cl_wait() { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }
DPS_DEV unsigned mapa_u32(unsigned addr, unsigned rank) {  // volatile: computed where it is written, not hoisted (and spilled)
  unsigned r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
// 64-bit load from the cluster's distributed shared memory (own or a neighbour CTA's tile); volatile: never moved across a barrier
DPS_DEV float2 ld_cluster2(unsigned addr) {
  float2 v;
  asm volatile("ld.shared::cluster.v2.f32 {%0,%1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(addr));
  return v;
}
DPS_DEV float2 ld_cluster2_or_zero(unsigned addr, int valid) {  // predicated: rows outside the image read as zero, no branch
  float2 v;
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.s32 p, %3, 0;\n\tmov.f32 %0, 0f00000000;\n\tmov.f32 %1, 0f00000000;\n\t"
      "@p ld.shared::cluster.v2.f32 {%0,%1}, [%2];\n\t}"
      : "=f"(v.x), "=f"(v.y)
      : "r"(addr), "r"(valid));
  return v;
}
// volatile read-only load: issued where it is written (ptxas otherwise sinks the y loads below the barrier that follows them)
DPS_DEV float4 ldg_ro4_pinned(const float* p) {
  float4 r;
  asm volatile("ld.global.nc.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
  return r;
}
DPS_DEV float2 fma2w(float w, float2 v, float2 acc) { return __ffma2_rn(make_float2(w, w), v, acc); }  // FFMA2 R, R, UR.F32, R

// Every FMA of the four passes is an FFMA2 (fma.rn.f32x2) whose weight is a uniform-register broadcast.  The vertical passes
// pair two adjacent COLUMNS (a thread owns columns 2p, 2p+1 and 16 rows: 16 + 2R 64-bit loads of a row-major tile, lanes along
// the row), the horizontal passes pair two adjacent ROWS (a thread owns a row pair and a run of 16 / 18 columns: 128-bit loads
// of a tile stored row-pair interleaved, T2[row pair][column] = (row a, row b), lanes along the row pairs).  All loops are
// fully unrolled over compile-time offsets and free of branches, so a pass is "load, 16 FFMA2, load, …" and nothing else.
// Both vertical passes start with the 16 window rows the CTA owns (61 % of the FMAs) BETWEEN the arrive and the wait of the
// cluster barrier that publishes the tile, and read the 2R halo rows afterwards: the skew between the 8 CTAs hides behind work.
template <int R>
__global__ void __cluster_dims__(kCluster, 1, 1) __launch_bounds__(kT, SepCfg<R>::CTAS) sep_guidance_kernel(const __grid_constant__ SepFusedArgs a) {
  using Cfg = SepCfg<R>;
  constexpr int TAPS = Cfg::TAPS, H = kRI * kCluster, TPITCH = Cfg::TPITCH, ZPITCH = Cfg::ZPITCH, SPITCH = Cfg::SPITCH;
  static_assert(R % 4 == 0 && R >= 4 && R <= 16, "radius");
  extern __shared__ __align__(16) float smem[];
  float* Sx = smem;                                    // region A: x → x̂₀ (32, 256) row-major → s = A_hᵀ r as two parity tiles
  float* TZ = smem + Cfg::A_FLOATS;                    // region B: ε → T2 (vertical pass, mirrored column pads) → Z2 (zero-padded r)
  float2* T2 = reinterpret_cast<float2*>(TZ);
  float* red = TZ + Cfg::B_FLOATS;                     // 64
  uint64_t* bar = reinterpret_cast<uint64_t*>(red + 64);

  cg::cluster_group cluster = cg::this_cluster();
  const int q = (int)cluster.block_rank();
  const int plane = blockIdx.x / kCluster, c = plane % a.C, n = plane / a.C;
  const int tid = threadIdx.x;
  const int p = tid & 127, h = tid >> 7;               // vertical passes: column pair, row half
  const int rp = tid & 15, cb = tid >> 4;              // horizontal passes: row pair, column block
  const int64_t poff = (int64_t)c * H * kW + (int64_t)q * kRI * kW;

  SEPF_T(0);
  if (tid == 0) {
#pragma unroll
    for (int ch = 0; ch < kChunks; ++ch) mbar_init(bar + ch, 1);
    mbar_init_fence();
  }
  __syncthreads();
  if (tid == 0) {
    const float* xg = a.src.x + n * a.src.x_stride + poff;
    const float* eg = a.src.eps + n * a.src.eps_stride + poff;
    constexpr unsigned bytes = kChunkRows * kW * sizeof(float);
#pragma unroll
    for (int ch = 0; ch < kChunks; ++ch) {
      mbar_expect_tx(bar + ch, 2u * bytes);
      bulk_load(Sx + ch * kChunkRows * kW, xg + ch * kChunkRows * kW, bytes, bar + ch);
      bulk_load(TZ + ch * kChunkRows * kW, eg + ch * kChunkRows * kW, bytes, bar + ch);
    }
  }
  // ---- 0. x̂₀ in place for my column pair and row half; clamp mask → one register (bit 2·rr + e: row 16h + rr, column 2p + e) ----
  unsigned pass_bits = 0;
  {
    const float lo = a.src.clip ? -1.0f : -INFINITY, hi = a.src.clip ? 1.0f : INFINITY;
#pragma unroll
    for (int chh = 0; chh < kChunks / 2; ++chh) {
      mbar_wait(bar + h * (kChunks / 2) + chh, 0);
#pragma unroll
      for (int r8 = 0; r8 < kChunkRows; ++r8) {
        const int rr = chh * kChunkRows + r8, r = h * kG + rr;
        float2* xs = reinterpret_cast<float2*>(Sx + r * kW) + p;
        const float2 pre = x0_pair_pre(*xs, reinterpret_cast<const float2*>(TZ + r * kW)[p], a.src.c1, a.src.c2);
        *xs = make_float2(fminf(fmaxf(pre.x, lo), hi), fminf(fmaxf(pre.y, lo), hi));
        pass_bits |= ((pre.x >= lo && pre.x <= hi) ? (1u << (2 * rr)) : 0u) | ((pre.y >= lo && pre.y <= hi) ? (2u << (2 * rr)) : 0u);
      }
    }
  }
  asm volatile("" : "+r"(pass_bits));  // one register, not sixteen scalar-replaced (and spilled) row masks
  SEPF_T(1);
  cl_arrive_publish(false);  // #1 (arrive): my x̂₀ values are in place

  // ---- 1. vertical forward for my column pair, rows 16h..16h+15 → T2 (row-pair interleaved, column-padded by mirroring) ----
  {
    float2 acc[kG];
#pragma unroll
    for (int j = 0; j < kG; ++j) acc[j] = make_float2(0.f, 0.f);
    {  // window rows R ≤ rr < 16 + R: rows 16h + (rr − R), the x̂₀ values this very thread has just written
      const float2* mid = reinterpret_cast<const float2*>(Sx + h * kG * kW) + p;
#pragma unroll
      for (int rr = R; rr < kG + R; ++rr) {
        const float2 v = mid[(rr - R) * (kW / 2)];
#pragma unroll
        for (int j = 0; j < kG; ++j) {
          const int k = rr - j;
          if (k >= 0 && k < TAPS) acc[j] = fma2w(a.wv[k], v, acc[j]);
        }
      }
    }
    SEPF_T(2);
    cl_wait();  // #1 (wait): every CTA's x̂₀ rows are in place; the ε rows are dead
    SEPF_T(3);
    {
      // The first R and the last R rows of the window are halo rows for one of the two halves: neighbour rows over DSMEM or, at
      // the image border, own rows mirrored without edge repeat (−m ↦ m, 31 + m ↦ 31 − m) — one base and one signed row stride.
      const unsigned sx = smem_u32(Sx);
      const unsigned own_a = mapa_u32(sx, (unsigned)q);
      unsigned a0, a1;
      int s0, s1;
      if (h == 0) {
        a0 = q > 0 ? mapa_u32(sx, (unsigned)(q - 1)) + (kRI - R) * kRowB : own_a + R * kRowB;
        s0 = q > 0 ? kRowB : -kRowB;
        a1 = own_a + kG * kRowB;
        s1 = kRowB;
      } else {
        a0 = own_a + (kG - R) * kRowB;
        s0 = kRowB;
        a1 = q < kCluster - 1 ? mapa_u32(sx, (unsigned)(q + 1)) : own_a + (kRI - 2) * kRowB;
        s1 = q < kCluster - 1 ? kRowB : -kRowB;
      }
      a0 += p * 8;
      a1 += p * 8;
#pragma unroll
      for (int rr = 0; rr < kG + 2 * R; ++rr) {
        if (rr >= R && rr < kG + R) continue;
        const float2 v = rr < R ? ld_cluster2(a0 + rr * s0) : ld_cluster2(a1 + (rr - kG - R) * s1);
#pragma unroll
        for (int j = 0; j < kG; ++j) {
          const int k = rr - j;
          if (k >= 0 && k < TAPS) acc[j] = fma2w(a.wv[k], v, acc[j]);
        }
      }
    }
    cl_arrive_done_reading(acc[0].x, acc[kG - 1].x);  // #2 (arrive): my reads of the neighbours' x̂₀ rows are done (their values have been consumed)
    float2* row0 = T2 + (size_t)(h * (kG / 2)) * TPITCH + Cfg::TP;
#pragma unroll
    for (int jp = 0; jp < kG / 2; ++jp)
      *reinterpret_cast<float4*>(row0 + jp * TPITCH + 2 * p) = make_float4(acc[2 * jp].x, acc[2 * jp + 1].x, acc[2 * jp].y, acc[2 * jp + 1].y);
    if (p <= R / 2 || p >= kW / 2 - 1 - R / 2) {  // the warps at the left / right edge also write the mirrored pad columns
      const int ca = 2 * p, cc = 2 * p + 1;
#pragma unroll
      for (int jp = 0; jp < kG / 2; ++jp) {
        float2* row = row0 + jp * TPITCH;
        const float2 c0 = make_float2(acc[2 * jp].x, acc[2 * jp + 1].x), c1 = make_float2(acc[2 * jp].y, acc[2 * jp + 1].y);
        if (ca >= 1 && ca <= R) row[-ca] = c0;                                   // column −m mirrors column m
        if (cc <= R) row[-cc] = c1;
        if (ca >= kW - 1 - R && ca <= kW - 2) row[2 * (kW - 1) - ca] = c0;       // column 255 + m mirrors 255 − m
        if (cc >= kW - 1 - R && cc <= kW - 2) row[2 * (kW - 1) - cc] = c1;
      }
    }
  }
  SEPF_T(4);
  __syncthreads();  // T2 complete
  SEPF_T(5);

  // ---- 2. horizontal forward + residual: thread = row pair rp, columns 16cb..16cb+15 ----
  float2 o[kNB];  // (row a, row b) per column; 16 used here, 18 by the adjoint
  {
    const float4* tp = reinterpret_cast<const float4*>(T2 + (size_t)rp * TPITCH + 16 * cb);  // padded column 16cb = image column 16cb − R
#pragma unroll
    for (int j = 0; j < 16; ++j) o[j] = make_float2(0.f, 0.f);
    // the measurement values of my 2 × 16 outputs are requested BEFORE the tap loop (L2-resident: one y for all particles; the
    // loads are pinned asm, so they issue here): their round trip hides behind the FMAs instead of standing in front of the
    // residual (162.9 → 156.9 µs at N = 128, 17.96 → 17.45 µs at N = 8; 80 registers either way)
    float4 ya[4], yb[4];
    if (a.y) {
      const float* yp = a.y + n * a.y_stride + poff + 2 * rp * kW + 16 * cb;
#pragma unroll
      for (int i = 0; i < 4; ++i) { ya[i] = ldg_ro4_pinned(yp + 4 * i); yb[i] = ldg_ro4_pinned(yp + kW + 4 * i); }
    }
#pragma unroll
    for (int m = 0; m < (16 + 2 * R) / 2; ++m) {
      const float4 t4 = tp[m];
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const float2 v = e ? make_float2(t4.z, t4.w) : make_float2(t4.x, t4.y);
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const int k = 2 * m + e - j;
          if (k >= 0 && k < TAPS) o[j] = fma2w(a.wh[k], v, o[j]);
        }
      }
    }
    SEPF_T(6);
    const int rowa = 2 * rp;
    if (a.y) {
      __syncthreads();  // everybody is done reading T2 → region B may take Z2
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        o[4 * i + 0] = make_float2(__fsub_rn(ya[i].x, o[4 * i + 0].x), __fsub_rn(yb[i].x, o[4 * i + 0].y));
        o[4 * i + 1] = make_float2(__fsub_rn(ya[i].y, o[4 * i + 1].x), __fsub_rn(yb[i].y, o[4 * i + 1].y));
        o[4 * i + 2] = make_float2(__fsub_rn(ya[i].z, o[4 * i + 2].x), __fsub_rn(yb[i].z, o[4 * i + 2].y));
        o[4 * i + 3] = make_float2(__fsub_rn(ya[i].w, o[4 * i + 3].x), __fsub_rn(yb[i].w, o[4 * i + 3].y));
      }
    } else {
      __syncthreads();
    }
    // ---- 3a. Z2 = zero-padded r, same interleaved layout ----
    float2* zrow = T2 + (size_t)rp * ZPITCH;
#pragma unroll
    for (int i = 0; i < 8; ++i)
      *reinterpret_cast<float4*>(zrow + Cfg::ZP + 16 * cb + 2 * i) = make_float4(o[2 * i].x, o[2 * i].y, o[2 * i + 1].x, o[2 * i + 1].y);
    const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int i = cb; i < Cfg::ZP / 2; i += 16) *reinterpret_cast<float4*>(zrow + 2 * i) = z4;
    for (int i = cb; i < (ZPITCH - Cfg::ZP - kW) / 2; i += 16) *reinterpret_cast<float4*>(zrow + Cfg::ZP + kW + 2 * i) = z4;
    if (a.r_out) {
      float* ro = a.r_out + ((int64_t)n * a.C + c) * H * kW + (int64_t)(q * kRI + rowa) * kW + 16 * cb;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        stg_stream4(ro + 4 * i, make_float4(o[4 * i].x, o[4 * i + 1].x, o[4 * i + 2].x, o[4 * i + 3].x));
        stg_stream4(ro + kW + 4 * i, make_float4(o[4 * i].y, o[4 * i + 1].y, o[4 * i + 2].y, o[4 * i + 3].y));
      }
    }
    if (a.partials) {  // per-CTA Σr², Σ|r|: warp sums now, the 8 per-warp values are added by warp 0 behind the next barrier
      float sq = 0.f, ab = 0.f;
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        sq = fmaf(o[j].x, o[j].x, fmaf(o[j].y, o[j].y, sq));
        ab += fabsf(o[j].x) + fabsf(o[j].y);
      }
      sq = warp_sum(sq);
      ab = warp_sum(ab);
      if ((tid & 31) == 0) {
        red[tid >> 5] = sq;
        red[32 + (tid >> 5)] = ab;
      }
    }
  }
  __syncthreads();  // Z2 and the per-warp sums are complete
  SEPF_T(7);
  if (a.partials && tid < 32) {
    float sq = tid < kT / 32 ? red[tid] : 0.0f, ab = tid < kT / 32 ? red[32 + tid] : 0.0f;
    sq = warp_sum(sq);
    ab = warp_sum(ab);
    if (tid == 0) {
      float* pp = a.partials + ((int64_t)n * (a.C * kCluster) + c * kCluster + q) * 2;
      pp[0] = sq;
      pp[1] = ab;
    }
  }
  // ---- 3b. horizontal adjoint on the PADDED domain: t[P] = Σ_d wh[R − d]·r[P + d] for P ∈ [−R, 256 + R), thread = row pair rp,
  //          18 columns from −16 + 18cb; the pad outputs are folded onto their mirror columns afterwards (Aᵀ = Pᵀ Cᵀ: what the
  //          mirrored padding read twice comes back twice). ----
  {
    const float4* tp = reinterpret_cast<const float4*>(T2 + (size_t)rp * ZPITCH + kNB * cb);  // Z2 index ZP + (−16 + 18cb) − R
#pragma unroll
    for (int j = 0; j < kNB; ++j) o[j] = make_float2(0.f, 0.f);
#pragma unroll
    for (int m = 0; m < (kNB + 2 * R) / 2; ++m) {
      const float4 t4 = tp[m];
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const float2 v = e ? make_float2(t4.z, t4.w) : make_float2(t4.x, t4.y);
#pragma unroll
        for (int j = 0; j < kNB; ++j) {
          const int k = 2 * m + e - j;
          if (k >= 0 && k < TAPS) o[j] = fma2w(a.wh[2 * R - k], v, o[j]);
        }
      }
    }
  }
  SEPF_T(8);
  cl_wait();  // #2 (wait): the neighbours are done reading my x̂₀ rows → region A may take s
  {
    float* se = Sx + (size_t)rp * SPITCH + Cfg::SP - kSL + kNB * cb;  // row 2rp (even tile); the odd tile follows at STILE
#pragma unroll
    for (int i = 0; i < kNB / 2; ++i) {
      const int P = -kSL + kNB * cb + 2 * i;
      if (P >= -R && P < kW + R) {
        *reinterpret_cast<float2*>(se + 2 * i) = make_float2(o[2 * i].x, o[2 * i + 1].x);
        *reinterpret_cast<float2*>(se + Cfg::STILE + 2 * i) = make_float2(o[2 * i].y, o[2 * i + 1].y);
      }
    }
  }
  __syncthreads();
  {  // fold the pad columns onto their mirrors: thread = row (tid & 31), m = 1 + (tid >> 5) + 8i, both sides
    const int row = tid & 31;
    float* srow = Sx + (row & 1) * Cfg::STILE + (size_t)(row >> 1) * SPITCH + Cfg::SP;
#pragma unroll
    for (int m = 1 + (tid >> 5); m <= R; m += kT / 32) {
      srow[m] += srow[-m];
      srow[kW - 1 - m] += srow[kW - 1 + m];
    }
  }
  __syncthreads();
  SEPF_T(9);
  cl_arrive_publish(true);  // #3 (arrive): my s rows are in place

  // ---- 4. vertical adjoint for my column pair, rows 16h..16h+15, + border folds + clamp mask ----
  {
    // window row lr = 16h − R + rr lives in parity tile (rr & 1) at tile row (lr >> 1): compile-time offsets from three bases
    constexpr int kRowPB = SPITCH * (int)sizeof(float), kTileB = Cfg::STILE * (int)sizeof(float);
    float2 acc[kG];
#pragma unroll
    for (int j = 0; j < kG; ++j) acc[j] = make_float2(0.f, 0.f);
    {
      const float* mid = Sx + (size_t)(h * (kG / 2)) * SPITCH + Cfg::SP + 2 * p;
#pragma unroll
      for (int rr = R; rr < kG + R; ++rr) {
        const float2 v = *reinterpret_cast<const float2*>(mid + (rr & 1) * Cfg::STILE + ((rr - R) / 2) * SPITCH);
#pragma unroll
        for (int j = 0; j < kG; ++j) {
          const int k = rr - j;
          if (k >= 0 && k < TAPS) acc[j] = fma2w(a.wv[2 * R - k], v, acc[j]);
        }
      }
    }
    if (q == 0 && h == 0) {  // image top: g[m] += Σ_{i=0}^{R−m} wv[R − m − i]·s[i],  1 ≤ m ≤ R
#pragma unroll
      for (int i = 0; i < R; ++i) {
        const float2 v = *reinterpret_cast<const float2*>(Sx + (i & 1) * Cfg::STILE + (size_t)(i >> 1) * SPITCH + Cfg::SP + 2 * p);
#pragma unroll
        for (int m = 1; m <= R && m < kG; ++m)
          if (i <= R - m) acc[m] = fma2w(a.wv[R - m - i], v, acc[m]);
      }
    }
    if (q == kCluster - 1 && h == 1) {  // image bottom: g[31 − a'] += Σ_{b=0}^{R−a'} wv[R + a' + b]·s[31 − b],  1 ≤ a' ≤ R
#pragma unroll
      for (int b = 0; b < R; ++b) {
        const int lr = kRI - 1 - b;
        const float2 v = *reinterpret_cast<const float2*>(Sx + (lr & 1) * Cfg::STILE + (size_t)(lr >> 1) * SPITCH + Cfg::SP + 2 * p);
#pragma unroll
        for (int ap = 1; ap <= R && ap < kG; ++ap)
          if (b <= R - ap) acc[kG - 1 - ap] = fma2w(a.wv[R + ap + b], v, acc[kG - 1 - ap]);
      }
    }
    SEPF_T(10);
    cl_wait();  // #3 (wait): every CTA's s rows are in place
    SEPF_T(11);
    {
      const unsigned sx = smem_u32(Sx);
      const unsigned colb = (Cfg::SP + 2 * p) * (unsigned)sizeof(float);
      // rr < R: tile row (h ? 8 : 16 of the upper CTA) + (rr − R)/2;  rr ≥ 16 + R: tile row (h ? 0 of the lower CTA : 8) + (rr − 16 − R)/2
      const unsigned b0 = (h == 0 ? mapa_u32(sx, (unsigned)(q > 0 ? q - 1 : q)) + (kRI / 2) * kRowPB : mapa_u32(sx, (unsigned)q) + (kG / 2) * kRowPB) + colb;
      const unsigned b1 = (h == 0 ? mapa_u32(sx, (unsigned)q) + (kG / 2) * kRowPB : mapa_u32(sx, (unsigned)(q < kCluster - 1 ? q + 1 : q))) + colb;
      const int v0 = (h == 1 || q > 0) ? 1 : 0, v1 = (h == 0 || q < kCluster - 1) ? 1 : 0;  // rows outside the image contribute nothing to the plain part
#pragma unroll
      for (int rr = 0; rr < kG + 2 * R; ++rr) {
        if (rr >= R && rr < kG + R) continue;
        const float2 v = rr < R ? ld_cluster2_or_zero(b0 + (rr & 1) * kTileB + ((rr - R - (rr & 1)) / 2) * kRowPB, v0)
                                : ld_cluster2_or_zero(b1 + (rr & 1) * kTileB + ((rr - kG - R) / 2) * kRowPB, v1);
#pragma unroll
        for (int j = 0; j < kG; ++j) {
          const int k = rr - j;
          if (k >= 0 && k < TAPS) acc[j] = fma2w(a.wv[2 * R - k], v, acc[j]);
        }
      }
    }
    cl_arrive_done_reading(acc[0].x, acc[kG - 1].x);  // #4 (arrive): my reads of the neighbours' s rows are done
    float* gp = a.g + n * a.g_stride + poff + (size_t)(h * kG) * kW + 2 * p;
#pragma unroll
    for (int j = 0; j < kG; ++j) {
      const unsigned bts = pass_bits >> (2 * j);
      stg_stream2(gp + j * kW, make_float2((bts & 1u) ? acc[j].x : 0.f, (bts & 2u) ? acc[j].y : 0.f));
    }
  }
  cl_wait();  // #4 (wait): no CTA leaves while a neighbour may still read its s rows
  SEPF_T(12);
#ifdef DPS_SEPF_TRACE
  if (threadIdx.x == 0 && blockIdx.x < 4096) { unsigned sm; asm("mov.u32 %0, %smid;" : "=r"(sm)); sepf_trace[blockIdx.x * 16 + 15] = sm; }
#endif
}

template <int R>
int launch_sepf(const dps_operator* op, const SepFusedArgs& a, int n, cudaStream_t st) {
  DPS_SMEM_OPTIN((sep_guidance_kernel<R>), SepCfg<R>::SMEM, op->device);
  dim3 grid((unsigned)((int64_t)op->C * n * kCluster));
  sep_guidance_kernel<R><<<grid, kT, SepCfg<R>::SMEM, st>>>(a);
  DPS_LAUNCH_CHECK("sep_guidance");
  return DPS_OK;
}
}  // namespace

#ifdef DPS_SEPF_TRACE
extern "C" int dps_debug_sepf_trace(long long* host_out) {
  return (int)cudaMemcpyFromSymbol(host_out, sepf_trace, sizeof(long long) * 4096 * 16);
}
#endif

// Called by sep_create with the raw 1-D taps (radius rv / rh around the centre).  Leaves op->sepfused null when not covered.
int sep_fused_create(dps_operator* op, const float* taps_v, int rv, const float* taps_h, int rh) {
  const int r = rv > rh ? rv : rh;
  if (op->H != kRI * kCluster || op->W != kW || r > 16 || r < 1) return DPS_OK;
  SepFused* t = new SepFused();
  t->R = (r + 3) / 4 * 4;
  for (int k = 0; k <= 2 * rv; ++k) t->wv[t->R - rv + k] = taps_v[k];
  for (int k = 0; k <= 2 * rh; ++k) t->wh[t->R - rh + k] = taps_h[k];
  op->sepfused = t;
  op->guidance_P = op->C * kCluster;
  return DPS_OK;
}

void sep_fused_destroy(dps_operator* op) {
  delete op->sepfused;
  op->sepfused = nullptr;
}

int sep_fused_guidance(const dps_operator* op, const dps_source& src, const float* y, int64_t y_stride, float* r_out, float* g,
                       int64_t g_stride, float* partials, int n, cudaStream_t st) {
  const SepFused& t = *op->sepfused;
  DPS_REQUIRE(src.eps, DPS_ERR_INVALID, "blur guidance: the fused kernel forms x̂₀ from x and ε (eps is required)");
  SepFusedArgs a;
  for (int k = 0; k < 33; ++k) { a.wv[k] = t.wv[k]; a.wh[k] = t.wh[k]; }
  a.C = op->C;
  a.src = src;
  a.y = y;
  a.y_stride = y_stride;
  a.r_out = r_out;
  a.g = g;
  a.g_stride = g_stride;
  a.partials = partials;
  switch (t.R) {
    case 4: return launch_sepf<4>(op, a, n, st);
    case 8: return launch_sepf<8>(op, a, n, st);
    case 12: return launch_sepf<12>(op, a, n, st);
    case 16: return launch_sepf<16>(op, a, n, st);
  }
  dps_set_error("blur guidance: unsupported radius %d", t.R);
  return DPS_ERR_UNSUPPORTED;
}
