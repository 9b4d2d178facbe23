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

template <int R>
__global__ void __cluster_dims__(kCluster, 1, 1) __launch_bounds__(kT, 3) sep_guidance_kernel(const __grid_constant__ SepFusedArgs a) {
  constexpr int TAPS = 2 * R + 1, PADW = kW + 2 * R, H = kRI * kCluster, NV = (4 + 2 * R) / 4;
  static_assert(R % 4 == 0 && R >= 4 && R <= 16, "radius");
  extern __shared__ __align__(16) float smem[];
  float* Sx = smem;                    // (32, 256)  x → x̂₀ → (pass 3) s = A_hᵀ r
  float* TZ = Sx + kRI * kW;           // (32, PADW) ε (first 32·256 floats) → T (pass 1) → zero-padded r (pass 3)
  float* E = TZ + kRI * PADW;          // (32, 2R)   folded border terms of the horizontal adjoint
  float* red = E + kRI * 2 * R;        // 64
  uint64_t* bar = reinterpret_cast<uint64_t*>(red + 64);

  cg::cluster_group cluster = cg::this_cluster();
  const int q = (int)cluster.block_rank();
  const int plane = blockIdx.x / kCluster, c = plane % a.C, n = plane / a.C;
  const int tid = threadIdx.x;
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
  // ---- 0. x̂₀ in place (thread = column); clamp mask → one register ----
  unsigned pass_bits = 0;
  {
    const float lo = a.src.clip ? -1.0f : -INFINITY, hi = a.src.clip ? 1.0f : INFINITY;
#pragma unroll
    for (int ch = 0; ch < kChunks; ++ch) {
      mbar_wait(bar + ch, 0);
#pragma unroll
      for (int rr = 0; rr < kChunkRows; ++rr) {
        const int r = ch * kChunkRows + rr;
        const float pre = x0_pre(Sx[r * kW + tid], TZ[r * kW + tid], a.src.c1, a.src.c2);
        Sx[r * kW + tid] = fminf(fmaxf(pre, lo), hi);
        pass_bits |= (pre >= lo && pre <= hi) ? (1u << r) : 0u;
      }
    }
  }
  cluster.sync();  // #1: every CTA's x̂₀ rows are in place; the ε rows are dead

  // ---- 1. vertical forward: T[i][col] = Σ_k wv[k] · x̂₀[ρ(32q + i − R + k)][col] ----
  {
    const float* up = q > 0 ? cluster.map_shared_rank(Sx, q - 1) : Sx;
    const float* dn = q < kCluster - 1 ? cluster.map_shared_rank(Sx, q + 1) : Sx;
#pragma unroll
    for (int gI = 0; gI < kRI / kG; ++gI) {
      float acc[kG];
#pragma unroll
      for (int j = 0; j < kG; ++j) acc[j] = 0.f;
#pragma unroll
      for (int rr = 0; rr < kG + 2 * R; ++rr) {
        const int lr = gI * kG - R + rr;  // compile-time
        float v;
        if (lr < 0)  // above my rows: neighbour rows, or (image top) my own rows mirrored without edge repeat: −m ↦ m
          v = q > 0 ? up[(kRI + lr) * kW + tid] : Sx[(-lr) * kW + tid];
        else if (lr >= kRI)  // below: neighbour rows, or (image bottom) 31 + m ↦ 31 − m
          v = q < kCluster - 1 ? dn[(lr - kRI) * kW + tid] : Sx[(2 * (kRI - 1) - lr) * kW + tid];
        else
          v = Sx[lr * kW + tid];
#pragma unroll
        for (int j = 0; j < kG; ++j) {
          const int k = rr - j;
          if (k >= 0 && k < TAPS) acc[j] = fmaf(a.wv[k], v, acc[j]);
        }
      }
#pragma unroll
      for (int j = 0; j < kG; ++j) {
        float* row = TZ + (gI * kG + j) * PADW + R;
        row[tid] = acc[j];
        if (tid >= 1 && tid <= R) row[-tid] = acc[j];                                  // column −m mirrors column m
        if (tid >= kW - 1 - R && tid <= kW - 2) row[2 * (kW - 1) - tid] = acc[j];      // column 255 + m mirrors 255 − m
      }
    }
  }
  cl_arrive();      // #2 (arrive): my reads of the neighbours' x̂₀ rows are done
  __syncthreads();  // T complete

  // ---- 2. horizontal forward + residual: thread = columns 4cg..4cg+3 of rows 8rg..8rg+7 ----
  const int cgi = tid & 63, rg = tid >> 6;
  float4 rres[kG];
  float sq = 0.f, ab = 0.f;
  {
    const float* yp = a.y ? a.y + n * a.y_stride + poff : nullptr;
#pragma unroll
    for (int rr = 0; rr < kG; ++rr) {
      const int row = rg * kG + rr;
      const float4 yv = yp ? ldg_ro4(yp + row * kW + 4 * cgi) : make_float4(0.f, 0.f, 0.f, 0.f);
      const float4* tp = reinterpret_cast<const float4*>(TZ + row * PADW + 4 * cgi);  // padded column 4cg = image column 4cg − R
      float v[4 * NV];
#pragma unroll
      for (int m = 0; m < NV; ++m) {
        const float4 t4 = tp[m];
        v[4 * m] = t4.x; v[4 * m + 1] = t4.y; v[4 * m + 2] = t4.z; v[4 * m + 3] = t4.w;
      }
      float o[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int k = 0; k < TAPS; ++k) {
#pragma unroll
        for (int j = 0; j < 4; ++j) o[j] = fmaf(a.wh[k], v[j + k], o[j]);
      }
      float4 res = make_float4(o[0], o[1], o[2], o[3]);
      if (yp) res = make_float4(__fsub_rn(yv.x, o[0]), __fsub_rn(yv.y, o[1]), __fsub_rn(yv.z, o[2]), __fsub_rn(yv.w, o[3]));
      rres[rr] = res;
      if (a.r_out) stg_stream4(a.r_out + ((int64_t)n * a.C + c) * H * kW + (int64_t)(q * kRI + row) * kW + 4 * cgi, res);
      sq += res.x * res.x + res.y * res.y + res.z * res.z + res.w * res.w;
      ab += fabsf(res.x) + fabsf(res.y) + fabsf(res.z) + fabsf(res.w);
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
  __syncthreads();  // everybody is done reading T
  // ---- 3. horizontal adjoint: zero-padded r → plain flipped-tap correlation + folded border terms ----
#pragma unroll
  for (int rr = 0; rr < kG; ++rr) {
    float* row = TZ + (rg * kG + rr) * PADW;
    *reinterpret_cast<float4*>(row + R + 4 * cgi) = rres[rr];
    if (cgi < R / 4) *reinterpret_cast<float4*>(row + 4 * cgi) = make_float4(0.f, 0.f, 0.f, 0.f);
    if (cgi >= 64 - R / 4) *reinterpret_cast<float4*>(row + 2 * R + 4 * cgi) = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  __syncthreads();
  for (int id = tid; id < kRI * 2 * R; id += kT) {  // folded terms: (row, side, m)
    const int row = id / (2 * R), rem = id - row * (2 * R), side = rem / R, m = rem - side * R + 1;
    const float* rrow = TZ + row * PADW + R;
    float s = 0.f;
    if (side == 0) {
      for (int i = 0; i <= R - m; ++i) s = fmaf(a.wh[R - m - i], rrow[i], s);
    } else {
      for (int b = 0; b <= R - m; ++b) s = fmaf(a.wh[R + m + b], rrow[kW - 1 - b], s);
    }
    E[row * 2 * R + side * R + (m - 1)] = s;
  }
#pragma unroll
  for (int rr = 0; rr < kG; ++rr) {  // plain part (rres is re-used for s)
    const int row = rg * kG + rr;
    const float4* tp = reinterpret_cast<const float4*>(TZ + row * PADW + 4 * cgi);
    float v[4 * NV];
#pragma unroll
    for (int m = 0; m < NV; ++m) {
      const float4 t4 = tp[m];
      v[4 * m] = t4.x; v[4 * m + 1] = t4.y; v[4 * m + 2] = t4.z; v[4 * m + 3] = t4.w;
    }
    float o[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int k = 0; k < TAPS; ++k) {
#pragma unroll
      for (int j = 0; j < 4; ++j) o[j] = fmaf(a.wh[2 * R - k], v[j + k], o[j]);
    }
    rres[rr] = make_float4(o[0], o[1], o[2], o[3]);
  }
  __syncthreads();  // E complete
  cl_wait();        // #2 (wait): the neighbours are done reading my x̂₀ rows → the buffer may take s
#pragma unroll
  for (int rr = 0; rr < kG; ++rr) {
    const int row = rg * kG + rr;
    float sv[4] = {rres[rr].x, rres[rr].y, rres[rr].z, rres[rr].w};
    if (cgi <= R / 4 || cgi >= 63 - R / 4) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int m = 4 * cgi + j;
        if (m >= 1 && m <= R) sv[j] += E[row * 2 * R + (m - 1)];
        if (m >= kW - 1 - R && m <= kW - 2) sv[j] += E[row * 2 * R + R + (kW - 1 - m) - 1];
      }
    }
    *reinterpret_cast<float4*>(Sx + row * kW + 4 * cgi) = make_float4(sv[0], sv[1], sv[2], sv[3]);
  }
  cluster.sync();  // #3: every CTA's s rows are in place

  // ---- 4. vertical adjoint (thread = column) + border folds + clamp mask ----
  {
    const float* up = q > 0 ? cluster.map_shared_rank(Sx, q - 1) : Sx;
    const float* dn = q < kCluster - 1 ? cluster.map_shared_rank(Sx, q + 1) : Sx;
    float* gp = a.g + n * a.g_stride + poff;
#pragma unroll
    for (int gI = 0; gI < kRI / kG; ++gI) {
      float acc[kG];
#pragma unroll
      for (int j = 0; j < kG; ++j) acc[j] = 0.f;
#pragma unroll
      for (int rr = 0; rr < kG + 2 * R; ++rr) {
        const int lr = gI * kG - R + rr;  // compile-time
        float v;
        if (lr < 0)  // rows outside the image contribute nothing to the plain part
          v = q > 0 ? up[(kRI + lr) * kW + tid] : 0.f;
        else if (lr >= kRI)
          v = q < kCluster - 1 ? dn[(lr - kRI) * kW + tid] : 0.f;
        else
          v = Sx[lr * kW + tid];
#pragma unroll
        for (int j = 0; j < kG; ++j) {
          const int k = rr - j;
          if (k >= 0 && k < TAPS) acc[j] = fmaf(a.wv[2 * R - k], v, acc[j]);
        }
      }
      if (q == 0) {  // image top: g[m] += Σ_{i=0}^{R−m} wv[R − m − i]·s[i],  1 ≤ m ≤ R
#pragma unroll
        for (int j = 0; j < kG; ++j) {
          const int m = gI * kG + j;
          if (m >= 1 && m <= R) {
#pragma unroll
            for (int i = 0; i <= R - m; ++i) acc[j] = fmaf(a.wv[R - m - i], Sx[i * kW + tid], acc[j]);
          }
        }
      }
      if (q == kCluster - 1) {  // image bottom: g[31 − a'] += Σ_{b=0}^{R−a'} wv[R + a' + b]·s[31 − b],  1 ≤ a' ≤ R
#pragma unroll
        for (int j = 0; j < kG; ++j) {
          const int ap = kRI - 1 - (gI * kG + j);
          if (ap >= 1 && ap <= R) {
#pragma unroll
            for (int b = 0; b <= R - ap; ++b) acc[j] = fmaf(a.wv[R + ap + b], Sx[(kRI - 1 - b) * kW + tid], acc[j]);
          }
        }
      }
#pragma unroll
      for (int j = 0; j < kG; ++j) {
        const int i = gI * kG + j;
        stg_stream(gp + i * kW + tid, ((pass_bits >> i) & 1u) ? acc[j] : 0.f);
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
