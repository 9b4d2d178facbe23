// Super-resolution guidance in ONE kernel: residual r = y − A x̂₀, its partial sums and the UNSCALED masked cotangent
// g = 1[|pre| ≤ 1] ⊙ Aᵀ r  (Resizer bicubic ↓F, util/resizer.py:55-74; condition_methods.py:33-39 through autograd).
//
// Why one kernel: the forward operator shrinks a 256×256 plane to a (256/F)² residual (4 KB at F = 4) that the adjoint
// immediately expands again.  As two kernels the step reads x, ε twice (forward: x̂₀; adjoint: the clamp mask) and bounces r
// through HBM: 5T + 2M bytes, two launches and a third one for the per-particle 1/‖r‖.  That factor commutes with Aᵀ, the
// mask and the UNet VJP (all linear in the cotangent), so it is applied by the posterior-update kernel instead
// (dps_update_ext), and nothing global stands between A and Aᵀ any more: 3T + M bytes, one launch.
//
// Mapping: a thread-block CLUSTER of 8 CTAs owns one (particle, channel) plane; CTA q owns image rows [32q, 32q+32) and
// residual rows [RJ·q, RJ·(q+1)), RJ = 32/F.
//   0. x, ε rows → shared memory by two 32 KB TMA bulk copies; x̂₀ (clamped) and the pre-clamp value replace them in place.
//   1. cluster barrier; H pass: a thread owns a column and walks the 32 + 2·HALO rows its RJ residual rows need — the halo
//      rows are read from the neighbour CTAs' shared memory (DSMEM), not from HBM.
//   2. W pass from the RJ×256 tile; r = y − (·), Σr², Σ|r| → one partial-sum pair per CTA; r stays in shared memory.
//   3. cluster barrier; Aᵀ: u = r·A_w for the RJ+4 residual rows that touch this CTA's image rows (2+2 of them from the
//      neighbours' shared memory), then g = A_hᵀ u, masked with the pre-clamp value still in shared memory; one coalesced
//      store per image row.
// HBM traffic = the algorithmic minimum: x, ε read once, g written once, y read once.
#include <cooperative_groups.h>

#include <vector>

#include "operator.cuh"

namespace cg = cooperative_groups;

struct ResizeFused {
  int F = 0;            // 4 or 8
  float* wf_h = nullptr;  // (oH, TAPS) folded forward weights along H: row j uses image rows F·j − HALO + k
  float* wf_w = nullptr;  // (oW, TAPS) the same along W
  float* at_h = nullptr;  // (H, 4) transposed band along H: image row i receives from residual rows j0(i) + d, d < 4
  float* at_w = nullptr;  // (W, 4) the same along W
};

namespace {
constexpr int kW = 256, kRI = 32, kCluster = 8, kT = 256;

template <int F>
struct Geo {
  static constexpr int TAPS = 4 * F, HALO = (TAPS - F) / 2, RJ = kRI / F, OW = kW / F, RU = RJ + 4;
  // first residual row/col (relative, may be negative) that touches image row/col p:  ceil((p + HALO − TAPS + 1) / F)
  static constexpr int j0(int p) { return (p + HALO - TAPS + 1 + 1024 * F + F - 1) / F - 1024; }
};

template <int F>
size_t fused_smem() {
  using G = Geo<F>;
  return sizeof(float) * ((size_t)2 * kRI * kW + (size_t)G::RJ * kW + (size_t)G::RJ * G::OW + (size_t)G::RJ * G::TAPS +
                          (size_t)G::OW * G::TAPS + (size_t)kRI * 4 + 64) + 16;
}

template <int F>
__global__ void __cluster_dims__(kCluster, 1, 1) __launch_bounds__(kT, 2)
    resize_guidance_kernel(const ResizeFused t, int C, const dps_source src, const float* __restrict__ y, int64_t y_stride,
                           float* __restrict__ r_out, float* __restrict__ g, int64_t g_stride, float* __restrict__ partials) {
  using G = Geo<F>;
  constexpr int TAPS = G::TAPS, HALO = G::HALO, RJ = G::RJ, OW = G::OW, RU = G::RU, H = kRI * kCluster;
  extern __shared__ __align__(16) float smem[];
  float* Sx = smem;                       // (32, 256)  x → x̂₀ (clamped)
  float* Se = Sx + kRI * kW;              // (32, 256)  ε → pre-clamp value
  float* St = Se + kRI * kW;              // (RJ, 256)  H-pass result
  float* Sr = St + RJ * kW;               // (RJ, OW)   residual rows of this CTA
  float* Wh = Sr + RJ * OW;               // (RJ, TAPS) folded H weights of my residual rows
  float* Ww = Wh + RJ * TAPS;             // (OW, TAPS) folded W weights
  float* Ah = Ww + OW * TAPS;             // (32, 4)    transposed H band of my image rows
  float* red = Ah + kRI * 4;              // 64
  uint64_t* bar = reinterpret_cast<uint64_t*>(red + 64);

  cg::cluster_group cluster = cg::this_cluster();
  const int q = (int)cluster.block_rank();
  const int plane = blockIdx.x / kCluster, c = plane % C, n = plane / C;
  const int tid = threadIdx.x;
  const int64_t poff = (int64_t)c * H * kW + (int64_t)q * kRI * kW;
  const float* xg = src.x + n * src.x_stride + poff;
  const float* eg = src.eps + n * src.eps_stride + poff;

  if (tid == 0) {
    mbar_init(bar, 1);
    mbar_init_fence();
  }
  __syncthreads();
  if (tid == 0) {
    mbar_expect_tx(bar, 2u * kRI * kW * sizeof(float));
    bulk_load(Sx, xg, kRI * kW * sizeof(float), bar);
    bulk_load(Se, eg, kRI * kW * sizeof(float), bar);
  }
  // tables and measurement values while the rows are in flight
  stage_async(Wh, t.wf_h + (size_t)q * RJ * TAPS, RJ * TAPS, tid, kT);
  stage_async(Ww, t.wf_w, OW * TAPS, tid, kT);
  stage_async(Ah, t.at_h + (size_t)q * kRI * 4, kRI * 4, tid, kT);
  const float4 aw = __ldg(reinterpret_cast<const float4*>(t.at_w) + tid);  // my column's transposed W band
  constexpr int kRPerThread = RJ * OW / kT;  // residual values per thread in the W pass: 2 (F=4) or … ≥ 1
  static_assert(RJ * OW % kT == 0 || RJ * OW < kT, "W-pass mapping");
  float yv[kRPerThread > 0 ? kRPerThread : 1];
  const float* yp = y ? y + n * y_stride + (int64_t)c * (H / F) * OW + (int64_t)q * RJ * OW : nullptr;
#pragma unroll
  for (int u = 0; u < (kRPerThread > 0 ? kRPerThread : 1); ++u) {
    const int o = tid + u * kT;
    yv[u] = (yp && o < RJ * OW) ? ldg_ro(yp + o) : 0.f;
  }
  stage_wait();
  mbar_wait(bar, 0);
  // x̂₀ and the pre-clamp value in place (column = thread: conflict-free)
  {
    const float lo = src.clip ? -1.0f : -INFINITY, hi = src.clip ? 1.0f : INFINITY;
#pragma unroll 8
    for (int r = 0; r < kRI; ++r) {
      const float pre = x0_pre(Sx[r * kW + tid], Se[r * kW + tid], src.c1, src.c2);
      Sx[r * kW + tid] = fminf(fmaxf(pre, lo), hi);
      Se[r * kW + tid] = pre;
    }
  }
  cluster.sync();  // every CTA's x̂₀ rows are in place

  // ---- 1. H pass: t[jj][col] = Σ_k Wh[jj][k] · x̂₀[F·(RJ·q + jj) − HALO + k][col] ----
  {
    const float* up = q > 0 ? cluster.map_shared_rank(Sx, q - 1) : Sx;
    const float* dn = q < kCluster - 1 ? cluster.map_shared_rank(Sx, q + 1) : Sx;
    float acc[RJ];
#pragma unroll
    for (int jj = 0; jj < RJ; ++jj) acc[jj] = 0.f;
#pragma unroll
    for (int w = 0; w < kRI + 2 * HALO; ++w) {
      const int lr = w - HALO;  // row relative to my first image row
      float v;
      if (lr < 0) {
        if (q == 0) continue;  // above the image: its folded weights are zero
        v = up[(kRI + lr) * kW + tid];
      } else if (lr >= kRI) {
        if (q == kCluster - 1) continue;
        v = dn[(lr - kRI) * kW + tid];
      } else {
        v = Sx[lr * kW + tid];
      }
#pragma unroll
      for (int jj = 0; jj < RJ; ++jj) {
        const int k = w - F * jj;  // compile-time after unrolling
        if (k >= 0 && k < TAPS) acc[jj] = fmaf(Wh[jj * TAPS + k], v, acc[jj]);
      }
    }
#pragma unroll
    for (int jj = 0; jj < RJ; ++jj) St[jj * kW + tid] = acc[jj];
  }
  __syncthreads();

  // ---- 2. W pass, residual, partial sums ----
  float sq = 0.f, ab = 0.f;
#pragma unroll
  for (int u = 0; u < (kRPerThread > 0 ? kRPerThread : 1); ++u) {
    const int o = tid + u * kT;
    if (o < RJ * OW) {
      const int jj = o / OW, l = o - jj * OW;
      const float* tr = St + jj * kW;
      const float* wl = Ww + l * TAPS;
      float a = 0.f;
#pragma unroll
      for (int k = 0; k < TAPS; ++k) {
        const int col = F * l - HALO + k;
        if (col >= 0 && col < kW) a = fmaf(wl[k], tr[col], a);
      }
      const float res = yp ? yv[u] - a : a;
      Sr[o] = res;
      if (r_out) r_out[((int64_t)n * C + c) * (H / F) * OW + (int64_t)q * RJ * OW + o] = res;
      sq = fmaf(res, res, sq);
      ab += fabsf(res);
    }
  }
  if (partials) {
    block_sum2(sq, ab, red);
    if (tid == 0) {
      float* pp = partials + ((int64_t)n * (C * kCluster) + c * kCluster + q) * 2;
      pp[0] = sq;
      pp[1] = ab;
    }
  }
  cluster.sync();  // every CTA's residual rows are in place

  // ---- 3. Aᵀ: u[m][col] = Σ_d aw[d] · r[RJ·q − 2 + m][l0(col) + d],  then g[ii][col] = mask · Σ_d Ah[ii][d] · u[m0(ii) + d] ----
  {
    const float* rup = q > 0 ? cluster.map_shared_rank(Sr, q - 1) : Sr;
    const float* rdn = q < kCluster - 1 ? cluster.map_shared_rank(Sr, q + 1) : Sr;
    // first residual column that touches image column `tid`:  ceil((tid + HALO − TAPS + 1)/F)
    const int num = tid + HALO - TAPS + 1;
    const int lfirst = num >= 0 ? (num + F - 1) / F : -((-num) / F);
    const float awv[4] = {aw.x, aw.y, aw.z, aw.w};
    float uu[RU];
#pragma unroll
    for (int m = 0; m < RU; ++m) {
      const int jr = m - 2;  // residual row relative to my first one
      const float* rr;
      if (jr < 0) {
        if (q == 0) { uu[m] = 0.f; continue; }
        rr = rup + (RJ + jr) * OW;
      } else if (jr >= RJ) {
        if (q == kCluster - 1) { uu[m] = 0.f; continue; }
        rr = rdn + (jr - RJ) * OW;
      } else {
        rr = Sr + jr * OW;
      }
      float a = 0.f;
#pragma unroll
      for (int d = 0; d < 4; ++d) {
        const int l = lfirst + d;
        if (l >= 0 && l < OW) a = fmaf(awv[d], rr[l], a);
      }
      uu[m] = a;
    }
    float* gp = g + n * g_stride + poff;
#pragma unroll
    for (int ii = 0; ii < kRI; ++ii) {
      // first residual row (relative to RJ·q − 2) that touches image row ii:  ceil((ii + HALO − TAPS + 1)/F) + 2
      const int m0 = G::j0(ii) + 2;  // compile-time after unrolling
      float a = 0.f;
#pragma unroll
      for (int d = 0; d < 4; ++d)
        if (m0 + d >= 0 && m0 + d < RU) a = fmaf(Ah[ii * 4 + d], uu[m0 + d], a);
      const float pre = Se[ii * kW + tid];
      const float pass = (!src.clip || (pre >= -1.0f && pre <= 1.0f)) ? 1.0f : 0.f;
      stg_stream(gp + ii * kW + tid, a * pass);
    }
  }
  cluster.sync();  // neighbours may still be reading my shared memory
}

// folded forward band: wf[j][k] = Σ A[j][F·j − HALO + k]  (dense A already carries the reflected taps merged)
std::vector<float> folded(const std::vector<double>& A, int out_len, int in_len, int F, int TAPS, int HALO, bool* ok) {
  std::vector<float> wf((size_t)out_len * TAPS, 0.f);
  for (int j = 0; j < out_len; ++j)
    for (int m = 0; m < in_len; ++m) {
      const double v = A[(size_t)j * in_len + m];
      if (v == 0.0) continue;
      const int k = m - (F * j - HALO);
      if (k < 0 || k >= TAPS) { *ok = false; continue; }
      wf[(size_t)j * TAPS + k] = (float)v;
    }
  return wf;
}
// transposed band: at[p][d] = A[j0(p) + d][p]
std::vector<float> transposed(const std::vector<double>& A, int out_len, int in_len, int F, int TAPS, int HALO, bool* ok) {
  std::vector<float> at((size_t)in_len * 4, 0.f);
  for (int p = 0; p < in_len; ++p) {
    const int num = p + HALO - TAPS + 1;
    const int j0 = num >= 0 ? (num + F - 1) / F : -((-num) / F);
    for (int j = 0; j < out_len; ++j) {
      const double v = A[(size_t)j * in_len + p];
      if (v == 0.0) continue;
      const int d = j - j0;
      if (d < 0 || d >= 4) { *ok = false; continue; }
      at[(size_t)p * 4 + d] = (float)v;
    }
  }
  return at;
}

int upload(float** dst, const std::vector<float>& v) {
  DPS_CUDA(cudaMalloc(dst, v.size() * sizeof(float)));
  DPS_CUDA(cudaMemcpy(*dst, v.data(), v.size() * sizeof(float), cudaMemcpyHostToDevice));
  return DPS_OK;
}
}  // namespace

// Called by resize_create with the dense operator matrices.  Leaves op->rfused null when the shape is not covered.
int resize_fused_create(dps_operator* op, const std::vector<double>& Ah, const std::vector<double>& Aw, int out_h, int out_w) {
  if (op->H != kRI * kCluster || op->W != kW || out_h != out_w || (op->H != 4 * out_h && op->H != 8 * out_h)) return DPS_OK;
  const int F = op->H / out_h, TAPS = 4 * F, HALO = (TAPS - F) / 2;
  bool ok = true;
  std::vector<float> wfh = folded(Ah, out_h, op->H, F, TAPS, HALO, &ok), wfw = folded(Aw, out_w, op->W, F, TAPS, HALO, &ok);
  std::vector<float> ath = transposed(Ah, out_h, op->H, F, TAPS, HALO, &ok), atw = transposed(Aw, out_w, op->W, F, TAPS, HALO, &ok);
  if (!ok) return DPS_OK;  // a band wider than the bicubic ×F one (other kernels / antialiasing off): not covered
  ResizeFused* t = new ResizeFused();
  t->F = F;
  op->rfused = t;
  if (int rc = upload(&t->wf_h, wfh)) return rc;
  if (int rc = upload(&t->wf_w, wfw)) return rc;
  if (int rc = upload(&t->at_h, ath)) return rc;
  if (int rc = upload(&t->at_w, atw)) return rc;
  if (F == 4)
    DPS_CUDA(cudaFuncSetAttribute(resize_guidance_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fused_smem<4>()));
  else
    DPS_CUDA(cudaFuncSetAttribute(resize_guidance_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fused_smem<8>()));
  op->guidance_P = op->C * kCluster;
  return DPS_OK;
}

void resize_fused_destroy(dps_operator* op) {
  ResizeFused* t = op->rfused;
  if (!t) return;
  cudaFree(t->wf_h); cudaFree(t->wf_w); cudaFree(t->at_h); cudaFree(t->at_w);
  delete t;
  op->rfused = nullptr;
}

int resize_fused_guidance(const dps_operator* op, const dps_source& src, const float* y, int64_t y_stride, float* r_out,
                          float* g, int64_t g_stride, float* partials, int n, cudaStream_t st) {
  const ResizeFused& t = *op->rfused;
  DPS_REQUIRE(src.eps, DPS_ERR_INVALID, "resize guidance: the fused kernel forms x̂₀ from x and ε (eps is required)");
  dim3 grid((unsigned)((int64_t)op->C * n * kCluster));
  if (t.F == 4)
    resize_guidance_kernel<4><<<grid, kT, fused_smem<4>(), st>>>(t, op->C, src, y, y_stride, r_out, g, g_stride, partials);
  else
    resize_guidance_kernel<8><<<grid, kT, fused_smem<8>(), st>>>(t, op->C, src, y, y_stride, r_out, g, g_stride, partials);
  DPS_LAUNCH_CHECK("resize_guidance");
  return DPS_OK;
}
