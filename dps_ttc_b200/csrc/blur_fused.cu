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
constexpr int kW = 256, kRI = 32, kCluster = 8, kT = 256, kChunks = 4, kChunkRows = kRI / kChunks, kG = 8;

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

template <int R>
size_t sepf_smem() {
  return sizeof(float) * ((size_t)kRI * kW + (size_t)kRI * (kW + 2 * R) + (size_t)kRI * 2 * R + 64) + 8 * kChunks;
}

DPS_DEV void cl_arrive() { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); }
DPS_DEV void cl_wait() { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }

// Packed arithmetic: every FMA of the four passes is an FFMA2 (fma.rn.f32x2).  The vertical passes pair two adjacent
// COLUMNS (a thread owns columns 2p, 2p+1 and half of the CTA's rows; 64-bit shared loads of the natural row-major
// layout), the horizontal passes pair two adjacent ROWS (a thread owns 4 columns of 4 row pairs) — for that the tiles
// between a vertical and a horizontal pass are stored row-pair interleaved: T2[row pair][column] = (row a, row b).
template <int R>
__global__ void __cluster_dims__(kCluster, 1, 1) __launch_bounds__(kT, 3) sep_guidance_kernel(const __grid_constant__ SepFusedArgs a) {
  constexpr int TAPS = 2 * R + 1, PADW = kW + 2 * R, H = kRI * kCluster, NV = 4 + 2 * R, RIH = kRI / 2;
  static_assert(R % 4 == 0 && R >= 4 && R <= 16, "radius");
  extern __shared__ __align__(16) float smem[];
  float* Sx = smem;                                   // (32, 256)  x → x̂₀ → (pass 3) s = A_hᵀ r, row-major
  float* TZ = Sx + kRI * kW;                          // ε (first 32·256 floats) → T2 (pass 1) → zero-padded r (pass 3)
  float2* T2 = reinterpret_cast<float2*>(TZ);         // (16 row pairs, PADW) of (row a, row b)
  float* E = TZ + kRI * PADW;                         // (32, 2R)   folded border terms of the horizontal adjoint
  float* red = E + kRI * 2 * R;                       // 64
  uint64_t* bar = reinterpret_cast<uint64_t*>(red + 64);

  cg::cluster_group cluster = cg::this_cluster();
  const int q = (int)cluster.block_rank();
  const int plane = blockIdx.x / kCluster, c = plane % a.C, n = plane / a.C;
  const int tid = threadIdx.x;
  const int p = tid & 127, h = tid >> 7;              // vertical passes: column pair, row half (warp-uniform)
  const int cgi = tid & 63, rq = tid >> 6;            // horizontal passes: 4 columns, 4 row pairs
  const int64_t poff = (int64_t)c * H * kW + (int64_t)q * kRI * kW;
  const float* xg = a.src.x + n * a.src.x_stride + poff;
  const float* eg = a.src.eps + n * a.src.eps_stride + poff;

  if (tid == 0) {
#pragma unroll
    for (int ch = 0; ch < kChunks; ++ch) mbar_init(bar + ch, 1);
    mbar_init_fence();
  }
  __syncthreads();
  if (tid == 0) {
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
        const int rr = chh * kChunkRows + r8, r = h * RIH + rr;
        float2* xs = reinterpret_cast<float2*>(Sx + r * kW) + p;
        const float2 pre = x0_pair_pre(*xs, reinterpret_cast<const float2*>(TZ + r * kW)[p], a.src.c1, a.src.c2);
        *xs = make_float2(fminf(fmaxf(pre.x, lo), hi), fminf(fmaxf(pre.y, lo), hi));
        pass_bits |= ((pre.x >= lo && pre.x <= hi) ? (1u << (2 * rr)) : 0u) | ((pre.y >= lo && pre.y <= hi) ? (2u << (2 * rr)) : 0u);
      }
    }
  }
  cluster.sync();  // #1: every CTA's x̂₀ rows are in place; the ε rows are dead

  // ---- 1. vertical forward for my column pair, rows 16h..16h+15 → T2 (row-pair interleaved, column-padded by mirroring) ----
  {
    const float* up = q > 0 ? cluster.map_shared_rank(Sx, q - 1) : Sx;
    const float* dn = q < kCluster - 1 ? cluster.map_shared_rank(Sx, q + 1) : Sx;
#pragma unroll
    for (int gI = 0; gI < RIH / kG; ++gI) {
      float2 acc[kG];
#pragma unroll
      for (int j = 0; j < kG; ++j) acc[j] = make_float2(0.f, 0.f);
#pragma unroll
      for (int rr = 0; rr < kG + 2 * R; ++rr) {
        const int lr = h * RIH + gI * kG - R + rr;  // warp-uniform
        const float* rowp;
        if (lr < 0)  // above my rows: neighbour rows, or (image top) my own rows mirrored without edge repeat: −m ↦ m
          rowp = q > 0 ? up + (kRI + lr) * kW : Sx + (-lr) * kW;
        else if (lr >= kRI)  // below: neighbour rows, or (image bottom) 31 + m ↦ 31 − m
          rowp = q < kCluster - 1 ? dn + (lr - kRI) * kW : Sx + (2 * (kRI - 1) - lr) * kW;
        else
          rowp = Sx + lr * kW;
        const float2 v = reinterpret_cast<const float2*>(rowp)[p];
#pragma unroll
        for (int j = 0; j < kG; ++j) {
          const int k = rr - j;
          if (k >= 0 && k < TAPS) acc[j] = __ffma2_rn(make_float2(a.wv[k], a.wv[k]), v, acc[j]);
        }
      }
#pragma unroll
      for (int jp = 0; jp < kG / 2; ++jp) {
        float2* row = T2 + (size_t)((h * RIH + gI * kG) / 2 + jp) * PADW + R;  // row pair (a, b) = rows 2jp, 2jp+1 of the group
        const float2 c0 = make_float2(acc[2 * jp].x, acc[2 * jp + 1].x), c1 = make_float2(acc[2 * jp].y, acc[2 * jp + 1].y);
        *reinterpret_cast<float4*>(row + 2 * p) = make_float4(c0.x, c0.y, c1.x, c1.y);
        const int ca = 2 * p, cb = 2 * p + 1;
        if (ca >= 1 && ca <= R) row[-ca] = c0;                                   // column −m mirrors column m
        if (cb <= R) row[-cb] = c1;
        if (ca >= kW - 1 - R && ca <= kW - 2) row[2 * (kW - 1) - ca] = c0;       // column 255 + m mirrors 255 − m
        if (cb >= kW - 1 - R && cb <= kW - 2) row[2 * (kW - 1) - cb] = c1;
      }
    }
  }
  cl_arrive();      // #2 (arrive): my reads of the neighbours' x̂₀ rows are done
  __syncthreads();  // T2 complete

  // ---- 2. horizontal forward + residual: thread = columns 4cg..4cg+3 of row pairs 4rq..4rq+3 ----
  float2 rres[4][4];  // [row pair][column] = (row a, row b)
  float sq = 0.f, ab = 0.f;
  {
    const float* yp = a.y ? a.y + n * a.y_stride + poff : nullptr;
#pragma unroll
    for (int rp = 0; rp < 4; ++rp) {
      const int rowa = (4 * rq + rp) * 2;
      const float4 ya = yp ? ldg_ro4(yp + rowa * kW + 4 * cgi) : make_float4(0.f, 0.f, 0.f, 0.f);
      const float4 yb = yp ? ldg_ro4(yp + (rowa + 1) * kW + 4 * cgi) : make_float4(0.f, 0.f, 0.f, 0.f);
      const float4* tp = reinterpret_cast<const float4*>(T2 + (size_t)(4 * rq + rp) * PADW + 4 * cgi);  // padded column 4cg = image column 4cg − R
      float2 v[NV];
#pragma unroll
      for (int m = 0; m < NV / 2; ++m) {
        const float4 t4 = tp[m];
        v[2 * m] = make_float2(t4.x, t4.y);
        v[2 * m + 1] = make_float2(t4.z, t4.w);
      }
      float2 o[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) o[j] = make_float2(0.f, 0.f);
#pragma unroll
      for (int k = 0; k < TAPS; ++k) {
        const float2 w2 = make_float2(a.wh[k], a.wh[k]);
#pragma unroll
        for (int j = 0; j < 4; ++j) o[j] = __ffma2_rn(w2, v[j + k], o[j]);
      }
      const float yav[4] = {ya.x, ya.y, ya.z, ya.w}, ybv[4] = {yb.x, yb.y, yb.z, yb.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float ra = yp ? __fsub_rn(yav[j], o[j].x) : o[j].x, rb = yp ? __fsub_rn(ybv[j], o[j].y) : o[j].y;
        rres[rp][j] = make_float2(ra, rb);
        sq = fmaf(ra, ra, fmaf(rb, rb, sq));
        ab += fabsf(ra) + fabsf(rb);
      }
      if (a.r_out) {
        float* ro = a.r_out + ((int64_t)n * a.C + c) * H * kW + (int64_t)(q * kRI + rowa) * kW + 4 * cgi;
        stg_stream4(ro, make_float4(rres[rp][0].x, rres[rp][1].x, rres[rp][2].x, rres[rp][3].x));
        stg_stream4(ro + kW, make_float4(rres[rp][0].y, rres[rp][1].y, rres[rp][2].y, rres[rp][3].y));
      }
    }
  }
  if (a.partials) {
    block_sum2(sq, ab, red);
    if (tid == 0) {
      float* pp = a.partials + ((int64_t)n * (a.C * kCluster) + c * kCluster + q) * 2;
      pp[0] = sq;
      pp[1] = ab;
    }
  }
  __syncthreads();  // everybody is done reading T2
  // ---- 3. horizontal adjoint: zero-padded r (same interleaved layout) → plain flipped-tap correlation + folded border terms ----
#pragma unroll
  for (int rp = 0; rp < 4; ++rp) {
    float2* row = T2 + (size_t)(4 * rq + rp) * PADW;
    *reinterpret_cast<float4*>(row + R + 4 * cgi) = make_float4(rres[rp][0].x, rres[rp][0].y, rres[rp][1].x, rres[rp][1].y);
    *reinterpret_cast<float4*>(row + R + 4 * cgi + 2) = make_float4(rres[rp][2].x, rres[rp][2].y, rres[rp][3].x, rres[rp][3].y);
    const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
    if (cgi < R / 4) { *reinterpret_cast<float4*>(row + 4 * cgi) = z4; *reinterpret_cast<float4*>(row + 4 * cgi + 2) = z4; }
    if (cgi >= 64 - R / 4) { *reinterpret_cast<float4*>(row + 2 * R + 4 * cgi) = z4; *reinterpret_cast<float4*>(row + 2 * R + 4 * cgi + 2) = z4; }
  }
  __syncthreads();
  for (int id = tid; id < kRI * 2 * R; id += kT) {  // folded terms: (row, side, m)
    const int row = id / (2 * R), rem = id - row * (2 * R), side = rem / R, m = rem - side * R + 1;
    const float* rrow = reinterpret_cast<const float*>(T2 + (size_t)(row >> 1) * PADW + R) + (row & 1);  // element i at rrow[2i]
    float s = 0.f;
    if (side == 0) {
      for (int i = 0; i <= R - m; ++i) s = fmaf(a.wh[R - m - i], rrow[2 * i], s);
    } else {
      for (int b = 0; b <= R - m; ++b) s = fmaf(a.wh[R + m + b], rrow[2 * (kW - 1 - b)], s);
    }
    E[row * 2 * R + side * R + (m - 1)] = s;
  }
#pragma unroll
  for (int rp = 0; rp < 4; ++rp) {  // plain part (rres is re-used for s)
    const float4* tp = reinterpret_cast<const float4*>(T2 + (size_t)(4 * rq + rp) * PADW + 4 * cgi);
    float2 v[NV];
#pragma unroll
    for (int m = 0; m < NV / 2; ++m) {
      const float4 t4 = tp[m];
      v[2 * m] = make_float2(t4.x, t4.y);
      v[2 * m + 1] = make_float2(t4.z, t4.w);
    }
    float2 o[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) o[j] = make_float2(0.f, 0.f);
#pragma unroll
    for (int k = 0; k < TAPS; ++k) {
      const float2 w2 = make_float2(a.wh[2 * R - k], a.wh[2 * R - k]);
#pragma unroll
      for (int j = 0; j < 4; ++j) o[j] = __ffma2_rn(w2, v[j + k], o[j]);
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) rres[rp][j] = o[j];
  }
  __syncthreads();  // E complete
  cl_wait();        // #2 (wait): the neighbours are done reading my x̂₀ rows → the buffer may take s
#pragma unroll
  for (int rp = 0; rp < 4; ++rp) {
    const int rowa = (4 * rq + rp) * 2;
    float sa[4], sb[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) { sa[j] = rres[rp][j].x; sb[j] = rres[rp][j].y; }
    if (cgi <= R / 4 || cgi >= 63 - R / 4) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int m = 4 * cgi + j;
        if (m >= 1 && m <= R) { sa[j] += E[rowa * 2 * R + (m - 1)]; sb[j] += E[(rowa + 1) * 2 * R + (m - 1)]; }
        if (m >= kW - 1 - R && m <= kW - 2) {
          sa[j] += E[rowa * 2 * R + R + (kW - 1 - m) - 1];
          sb[j] += E[(rowa + 1) * 2 * R + R + (kW - 1 - m) - 1];
        }
      }
    }
    *reinterpret_cast<float4*>(Sx + rowa * kW + 4 * cgi) = make_float4(sa[0], sa[1], sa[2], sa[3]);
    *reinterpret_cast<float4*>(Sx + (rowa + 1) * kW + 4 * cgi) = make_float4(sb[0], sb[1], sb[2], sb[3]);
  }
  cluster.sync();  // #3: every CTA's s rows are in place

  // ---- 4. vertical adjoint for my column pair, rows 16h..16h+15, + border folds + clamp mask ----
  {
    const float* up = q > 0 ? cluster.map_shared_rank(Sx, q - 1) : Sx;
    const float* dn = q < kCluster - 1 ? cluster.map_shared_rank(Sx, q + 1) : Sx;
    float* gp = a.g + n * a.g_stride + poff;
    const float2* sx2 = reinterpret_cast<const float2*>(Sx);
#pragma unroll
    for (int gI = 0; gI < RIH / kG; ++gI) {
      float2 acc[kG];
#pragma unroll
      for (int j = 0; j < kG; ++j) acc[j] = make_float2(0.f, 0.f);
#pragma unroll
      for (int rr = 0; rr < kG + 2 * R; ++rr) {
        const int lr = h * RIH + gI * kG - R + rr;  // warp-uniform
        float2 v = make_float2(0.f, 0.f);            // rows outside the image contribute nothing to the plain part
        if (lr < 0) {
          if (q > 0) v = reinterpret_cast<const float2*>(up + (kRI + lr) * kW)[p];
        } else if (lr >= kRI) {
          if (q < kCluster - 1) v = reinterpret_cast<const float2*>(dn + (lr - kRI) * kW)[p];
        } else {
          v = sx2[lr * (kW / 2) + p];
        }
#pragma unroll
        for (int j = 0; j < kG; ++j) {
          const int k = rr - j;
          if (k >= 0 && k < TAPS) acc[j] = __ffma2_rn(make_float2(a.wv[2 * R - k], a.wv[2 * R - k]), v, acc[j]);
        }
      }
      if (q == 0) {  // image top: g[m] += Σ_{i=0}^{R−m} wv[R − m − i]·s[i],  1 ≤ m ≤ R   (m is warp-uniform)
#pragma unroll
        for (int j = 0; j < kG; ++j) {
          const int m = h * RIH + gI * kG + j;
          if (m >= 1 && m <= R) {
#pragma unroll
            for (int i = 0; i < R; ++i)
              if (i <= R - m) acc[j] = __ffma2_rn(make_float2(a.wv[R - m - i], a.wv[R - m - i]), sx2[i * (kW / 2) + p], acc[j]);
          }
        }
      }
      if (q == kCluster - 1) {  // image bottom: g[31 − a'] += Σ_{b=0}^{R−a'} wv[R + a' + b]·s[31 − b],  1 ≤ a' ≤ R
#pragma unroll
        for (int j = 0; j < kG; ++j) {
          const int ap = kRI - 1 - (h * RIH + gI * kG + j);
          if (ap >= 1 && ap <= R) {
#pragma unroll
            for (int b = 0; b < R; ++b)
              if (b <= R - ap)
                acc[j] = __ffma2_rn(make_float2(a.wv[R + ap + b], a.wv[R + ap + b]), sx2[(kRI - 1 - b) * (kW / 2) + p], acc[j]);
          }
        }
      }
#pragma unroll
      for (int j = 0; j < kG; ++j) {
        const int il = gI * kG + j;  // row within my half
        const unsigned bts = pass_bits >> (2 * il);
        stg_stream2(gp + (h * RIH + il) * kW + 2 * p, make_float2((bts & 1u) ? acc[j].x : 0.f, (bts & 2u) ? acc[j].y : 0.f));
      }
    }
  }
  cluster.sync();  // #4: the neighbours may still be reading my s rows
}

template <int R>
int launch_sepf(const dps_operator* op, const SepFusedArgs& a, int n, cudaStream_t st) {
  DPS_SMEM_OPTIN((sep_guidance_kernel<R>), sepf_smem<R>(), op->device);
  dim3 grid((unsigned)((int64_t)op->C * n * kCluster));
  sep_guidance_kernel<R><<<grid, kT, sepf_smem<R>(), st>>>(a);
  DPS_LAUNCH_CHECK("sep_guidance");
  return DPS_OK;
}
}  // namespace

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
