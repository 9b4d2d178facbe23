// Super-resolution guidance in ONE kernel: residual r = y − A x̂₀, its partial sums and the UNSCALED masked cotangent
// g = 1[|pre| ≤ 1] ⊙ Aᵀ r  (Resizer bicubic ↓F, util/resizer.py:55-74; condition_methods.py:33-39 through autograd).
//
// Why one kernel: the forward operator shrinks a 256×256 plane to a (256/F)² residual (4 KB at F = 4) that the adjoint
// immediately expands again.  As two kernels the step reads x, ε twice (forward: x̂₀; adjoint: the clamp mask) and bounces r
// through HBM: 5T + 2M bytes, two launches and a third one for the per-particle 1/‖r‖.  That factor commutes with Aᵀ, the
// mask and the UNet VJP (all linear in the cotangent), so it is applied by the posterior-update kernel instead
// (dps_update_ext), and nothing global stands between A and Aᵀ any more: 3T + M bytes, one launch.
//
// The Resizer with an integer factor is a strided convolution with ONE weight vector w[4F] over the symmetrically padded
// image (field_of_view = mirror[...], weights independent of the output position; checked at plan creation, else no fused
// kernel).  So the weights travel as kernel parameters (constant bank: no shared-memory table, no bank conflicts) and the
// borders are handled by mirrored reads of rows / columns the CTA already holds.
//
// Mapping: a thread-block CLUSTER of 8 CTAs owns one (particle, channel) plane; CTA q owns image rows [32q, 32q+32) and
// residual rows [RJ·q, RJ·(q+1)), RJ = 32/F.
//   0. x, ε rows → shared memory by TMA bulk copies in four 8-row chunks; each chunk becomes x̂₀ (clamped) in place as it
//      lands; the clamp mask of a thread's column is one 32-bit register.  The ε buffer is dead afterwards and is reused
//      for the H-pass tile and the residual rows (3 CTAs per SM).
//   1. cluster barrier; H pass: a thread owns a column and walks the 32 + 2·HALO rows its RJ residual rows need — halo
//      rows come from the neighbour CTAs' shared memory (DSMEM), mirrored own rows at the image border.
//   2. W pass from the column-padded RJ×(256+2·HALO) tile with 128-bit shared loads; r = y − (·), Σr², Σ|r| → one
//      partial-sum pair per CTA; r stays in shared memory.
//   3. cluster barrier (arrive early, wait late); Aᵀ: u = r·A_w for the RJ+4 residual rows that touch this CTA's image
//      rows (2+2 of them from the neighbours), then g = A_hᵀ u, masked, one coalesced store per image row.
// HBM traffic = the algorithmic minimum: x, ε read once, g written once, y read once.
#include <cooperative_groups.h>

#include <vector>

#include "operator.cuh"

namespace cg = cooperative_groups;

struct ResizeFused {
  int F = 0;              // 4 or 8
  float w[32] = {};       // the convolution weights (4F of them)
  float* at_h = nullptr;  // (H, 4) transposed band along H: image row i receives from residual rows j0(i) + d, d < 4
  float* at_w = nullptr;  // (W, 4) the same along W
};

namespace {
constexpr int kW = 256, kRI = 32, kCluster = 8, kT = 256, kChunks = 4, kChunkRows = kRI / kChunks;

struct FusedArgs {
  float w[32];
  const float* at_h;
  const float* at_w;
  int C;
  dps_source src;
  const float* y;
  int64_t y_stride;
  float* r_out;
  float* g;
  int64_t g_stride;
  float* partials;
};

template <int F>
struct Geo {
  static constexpr int TAPS = 4 * F, HALO = (TAPS - F) / 2, RJ = kRI / F, OW = kW / F, RU = RJ + 4, PADW = kW + 2 * HALO;
  // first residual row/col (relative, may be negative) that touches image row/col p:  ceil((p + HALO − TAPS + 1) / F)
  static constexpr int j0(int p) { return (p + HALO - TAPS + 1 + 1024 * F + F - 1) / F - 1024; }
};

template <int F>
size_t fused_smem() {
  return sizeof(float) * ((size_t)2 * kRI * kW + (size_t)2 * kRI * 4 + 64) + 8 * kChunks;  // x, ε | Ah2 (float2) | red | barriers
}

DPS_DEV void cluster_arrive() { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); }
DPS_DEV void cluster_wait() { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }

template <int F>
__global__ void __cluster_dims__(kCluster, 1, 1) __launch_bounds__(kT, 3) resize_guidance_kernel(const __grid_constant__ FusedArgs a) {
  using G = Geo<F>;
  constexpr int TAPS = G::TAPS, HALO = G::HALO, RJ = G::RJ, OW = G::OW, PADW = G::PADW, H = kRI * kCluster;
  constexpr int RJH = RJ / 2, RIH = kRI / 2, RUH = RJH + 4;  // per row-half: residual rows, image rows, u rows
  static_assert(RJ % 2 == 0 && HALO % 2 == 0, "column pairs / row halves");
  static_assert((size_t)RJ * PADW + (size_t)2 * RJ * OW <= (size_t)kRI * kW, "tile + residual rows fit into the dead ε buffer");
  extern __shared__ __align__(16) float smem[];
  float* Sx = smem;                       // (32, 256)  x → x̂₀ (clamped)
  float* Se = Sx + kRI * kW;              // (32, 256)  ε; dead after step 0 → St, Sr2
  float* St = Se;                         // (RJ, PADW) H-pass result, column-padded by mirroring
  float2* Sr2 = reinterpret_cast<float2*>(St + RJ * PADW);  // (RJ, OW) residual rows of this CTA, each value duplicated (r, r)
  float2* Ah2 = reinterpret_cast<float2*>(Se + kRI * kW);   // (32, 4)  transposed H band of my image rows, duplicated (w, w)
  float* red = reinterpret_cast<float*>(Ah2 + kRI * 4);     // 64
  uint64_t* bar = reinterpret_cast<uint64_t*>(red + 64);    // kChunks barriers

  cg::cluster_group cluster = cg::this_cluster();
  const int q = (int)cluster.block_rank();
  const int plane = blockIdx.x / kCluster, c = plane % a.C, n = plane / a.C;
  const int tid = threadIdx.x;
  const int p = tid & 127, h = tid >> 7;  // column pair (columns 2p, 2p+1) and row half; a warp has one h
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
      bulk_load(Se + ch * kChunkRows * kW, eg + ch * kChunkRows * kW, bytes, bar + ch);
    }
  }
  // tables and measurement values while the rows are in flight
  if (tid < kRI * 4) {
    const float v = __ldg(a.at_h + (size_t)q * kRI * 4 + tid);
    Ah2[tid] = make_float2(v, v);
  }
  // transposed W band of my two columns as 5 weight pairs over the residual columns lmin … lmin+4
  int lmin;
  float2 aw2[5];
  {
    const float4 w0 = __ldg(reinterpret_cast<const float4*>(a.at_w) + 2 * p), w1 = __ldg(reinterpret_cast<const float4*>(a.at_w) + 2 * p + 1);
    const int n0 = 2 * p + HALO - TAPS + 1, n1 = n0 + 1;
    const int l0 = n0 >= 0 ? (n0 + F - 1) / F : -((-n0) / F), l1 = n1 >= 0 ? (n1 + F - 1) / F : -((-n1) / F);
    lmin = l0;
    const bool sh = l1 != l0;  // the odd column starts one residual column later
    aw2[0] = make_float2(w0.x, sh ? 0.f : w1.x);
    aw2[1] = make_float2(w0.y, sh ? w1.x : w1.y);
    aw2[2] = make_float2(w0.z, sh ? w1.y : w1.z);
    aw2[3] = make_float2(w0.w, sh ? w1.z : w1.w);
    aw2[4] = make_float2(0.f, sh ? w1.w : 0.f);
  }
  constexpr int kRP = (RJ * OW + kT - 1) / kT;  // residual values per thread in the W pass
  float yv[kRP];
  const float* yp = a.y ? a.y + n * a.y_stride + (int64_t)c * (H / F) * OW + (int64_t)q * RJ * OW : nullptr;
#pragma unroll
  for (int u = 0; u < kRP; ++u) {
    const int o = tid + u * kT;
    yv[u] = (yp && o < RJ * OW) ? ldg_ro(yp + o) : 0.f;
  }
  // ---- 0. x̂₀ in place for my column pair and row half, chunk by chunk as the copies land; clamp mask → one register ----
  unsigned pass_bits = 0;  // bit 2·rr + e: row 16h + rr, column 2p + e
  {
    const float lo = a.src.clip ? -1.0f : -INFINITY, hi = a.src.clip ? 1.0f : INFINITY;
    constexpr int kChunksPerHalf = kChunks / 2;
#pragma unroll
    for (int chh = 0; chh < kChunksPerHalf; ++chh) {
      mbar_wait(bar + h * kChunksPerHalf + chh, 0);
#pragma unroll
      for (int r8 = 0; r8 < kChunkRows; ++r8) {
        const int rr = chh * kChunkRows + r8, r = h * RIH + rr;
        float2* xs = reinterpret_cast<float2*>(Sx + r * kW) + p;
        const float2 pre = x0_pair_pre(*xs, reinterpret_cast<const float2*>(Se + r * kW)[p], a.src.c1, a.src.c2);
        *xs = make_float2(fminf(fmaxf(pre.x, lo), hi), fminf(fmaxf(pre.y, lo), hi));
        pass_bits |= ((pre.x >= lo && pre.x <= hi) ? (1u << (2 * rr)) : 0u) | ((pre.y >= lo && pre.y <= hi) ? (2u << (2 * rr)) : 0u);
      }
    }
  }
  cluster.sync();  // every CTA's x̂₀ rows are in place (and nobody reads the ε buffer any more)

  // ---- 1. H pass for my RJH residual rows: t[jj][cols] = Σ_k w[k] · x̂₀[sym(F·(RJ·q + jj) − HALO + k)][cols] ----
  {
    const float* up = q > 0 ? cluster.map_shared_rank(Sx, q - 1) : Sx;
    const float* dn = q < kCluster - 1 ? cluster.map_shared_rank(Sx, q + 1) : Sx;
    float2 acc[RJH];
#pragma unroll
    for (int jl = 0; jl < RJH; ++jl) acc[jl] = make_float2(0.f, 0.f);
#pragma unroll
    for (int ww = 0; ww < F * (RJH - 1) + TAPS; ++ww) {
      const int lr = F * RJH * h + ww - HALO;  // image row relative to my CTA's first row (warp-uniform)
      const float* rowp;
      if (lr < 0)  // above my rows: the neighbour's last rows, or (top of the image) my own rows mirrored: −1 ↦ 0, −2 ↦ 1 …
        rowp = q > 0 ? up + (kRI + lr) * kW : Sx + (-lr - 1) * kW;
      else if (lr >= kRI)
        rowp = q < kCluster - 1 ? dn + (lr - kRI) * kW : Sx + (2 * kRI - 1 - lr) * kW;
      else
        rowp = Sx + lr * kW;
      const float2 v = reinterpret_cast<const float2*>(rowp)[p];
#pragma unroll
      for (int jl = 0; jl < RJH; ++jl) {
        const int k = ww - F * jl;  // compile-time after unrolling
        if (k >= 0 && k < TAPS) acc[jl] = __ffma2_rn(make_float2(a.w[k], a.w[k]), v, acc[jl]);
      }
    }
#pragma unroll
    for (int jl = 0; jl < RJH; ++jl) {
      float* row = St + (RJH * h + jl) * PADW + HALO;
      reinterpret_cast<float2*>(row)[p] = acc[jl];
      if (2 * p < HALO) { row[-1 - 2 * p] = acc[jl].x; row[-2 - 2 * p] = acc[jl].y; }             // columns −1, −2, … mirror 0, 1, …
      if (2 * p >= kW - HALO) { row[2 * kW - 1 - 2 * p] = acc[jl].x; row[2 * kW - 2 - 2 * p] = acc[jl].y; }  // 256, 257, … mirror 255, 254, …
    }
  }
  __syncthreads();

  // ---- 2. W pass, residual, partial sums ----
  float sq = 0.f, ab = 0.f;
#pragma unroll
  for (int u = 0; u < kRP; ++u) {
    const int o = tid + u * kT;
    if (o < RJ * OW) {
      const int jj = o / OW, l = o - jj * OW;
      const float4* tr = reinterpret_cast<const float4*>(St + jj * PADW + F * l);  // padded column F·l = image column F·l − HALO
      float acc = 0.f;
#pragma unroll
      for (int m = 0; m < TAPS / 4; ++m) {
        const float4 t4 = tr[m];
        acc = fmaf(a.w[4 * m + 0], t4.x, acc);
        acc = fmaf(a.w[4 * m + 1], t4.y, acc);
        acc = fmaf(a.w[4 * m + 2], t4.z, acc);
        acc = fmaf(a.w[4 * m + 3], t4.w, acc);
      }
      const float res = yp ? yv[u] - acc : acc;
      Sr2[o] = make_float2(res, res);
      if (a.r_out) a.r_out[((int64_t)n * a.C + c) * (H / F) * OW + (int64_t)q * RJ * OW + o] = res;
      sq = fmaf(res, res, sq);
      ab += fabsf(res);
    }
  }
  __syncthreads();   // my residual rows are complete …
  cluster_arrive();  // … and announced; the partial sums below overlap the other CTAs' arrival
  if (a.partials) {
    block_sum2(sq, ab, red);
    if (tid == 0) {
      float* pp = a.partials + ((int64_t)n * (a.C * kCluster) + c * kCluster + q) * 2;
      pp[0] = sq;
      pp[1] = ab;
    }
  }
  // ---- 3. Aᵀ for my column pair and row half: u[m] = Σ_d aw2[d] · r[RJ·q − 2 + RJH·h + m][lmin + d],  m < RUH,
  //         then g[16h + il] = mask · Σ_d Ah[16h + il][d] · u[m0(il) + d] ----
  {
    int li[5];
#pragma unroll
    for (int d = 0; d < 5; ++d) li[d] = min(max(lmin + d, 0), OW - 1);  // out-of-range columns carry zero weights
    float2 uu[RUH];
    auto urow = [&](const float2* rr) {
      float2 s2 = make_float2(0.f, 0.f);
#pragma unroll
      for (int d = 0; d < 5; ++d) s2 = __ffma2_rn(aw2[d], rr[li[d]], s2);
      return s2;
    };
    // residual row of u-slot m, relative to my CTA's first residual row: jr = RJH·h + m − 2 ∈ [−2, RJ + 2)
    bool remote[RUH];
#pragma unroll
    for (int m = 0; m < RUH; ++m) {
      const int jr = RJH * h + m - 2;
      remote[m] = jr < 0 || jr >= RJ;
      uu[m] = remote[m] ? make_float2(0.f, 0.f) : urow(Sr2 + jr * OW);  // my own rows first
    }
    cluster_wait();  // every CTA's residual rows are in place
    const float2* rup = q > 0 ? cluster.map_shared_rank(Sr2, q - 1) : Sr2;
    const float2* rdn = q < kCluster - 1 ? cluster.map_shared_rank(Sr2, q + 1) : Sr2;
#pragma unroll
    for (int m = 0; m < RUH; ++m) {
      const int jr = RJH * h + m - 2;
      if (jr < 0 && q > 0) uu[m] = urow(rup + (RJ + jr) * OW);              // residual rows outside the image do not exist
      if (jr >= RJ && q < kCluster - 1) uu[m] = urow(rdn + (jr - RJ) * OW);
    }
    cluster_arrive();  // my remote reads are done: the neighbours may exit once everybody has said so
    float* gp = a.g + n * a.g_stride + poff + (int64_t)h * RIH * kW;
#pragma unroll
    for (int il = 0; il < RIH; ++il) {
      const int m0 = G::j0(il) + 2;  // first u-slot that touches image row 16h + il (compile-time: F·RJH = 16 rows per half)
      float2 s2 = make_float2(0.f, 0.f);
#pragma unroll
      for (int d = 0; d < 4; ++d)
        if (m0 + d >= 0 && m0 + d < RUH) s2 = __ffma2_rn(Ah2[(h * RIH + il) * 4 + d], uu[m0 + d], s2);
      const unsigned b = pass_bits >> (2 * il);
      stg_stream2(gp + il * kW + 2 * p, make_float2((b & 1u) ? s2.x : 0.f, (b & 2u) ? s2.y : 0.f));
    }
    cluster_wait();
  }
}

int sym_idx(int p, int n) { return p < 0 ? -p - 1 : (p >= n ? 2 * n - 1 - p : p); }

// Is A (out_len × in_len, dense) the stride-F convolution with weights w[k] = A[mid][F·mid − HALO + k] over the symmetrically
// padded signal?  (True for the Resizer with an integer factor: util/resizer.py:104-167 builds every row from the same kernel
// samples and mirrors the field of view.)
bool is_symmetric_conv(const std::vector<double>& A, int out_len, int in_len, int F, int TAPS, int HALO, float* w) {
  const int mid = out_len / 2;
  double wmax = 0.0;
  for (int k = 0; k < TAPS; ++k) {
    const int p = F * mid - HALO + k;
    if (p < 0 || p >= in_len) return false;
    w[k] = (float)A[(size_t)mid * in_len + p];
    wmax = std::max(wmax, fabs((double)w[k]));
  }
  std::vector<double> row(in_len);
  for (int j = 0; j < out_len; ++j) {
    std::fill(row.begin(), row.end(), 0.0);
    for (int k = 0; k < TAPS; ++k) row[sym_idx(F * j - HALO + k, in_len)] += (double)w[k];
    for (int m = 0; m < in_len; ++m)
      if (fabs(row[m] - A[(size_t)j * in_len + m]) > 1e-6 * wmax) return false;
  }
  return true;
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
  float wh[32] = {}, ww[32] = {};
  if (!is_symmetric_conv(Ah, out_h, op->H, F, TAPS, HALO, wh) || !is_symmetric_conv(Aw, out_w, op->W, F, TAPS, HALO, ww)) return DPS_OK;
  for (int k = 0; k < TAPS; ++k)
    if (wh[k] != ww[k]) return DPS_OK;  // one weight vector serves both axes (square images, same factor)
  bool ok = true;
  std::vector<float> ath = transposed(Ah, out_h, op->H, F, TAPS, HALO, &ok), atw = transposed(Aw, out_w, op->W, F, TAPS, HALO, &ok);
  if (!ok) return DPS_OK;
  ResizeFused* t = new ResizeFused();
  t->F = F;
  for (int k = 0; k < TAPS; ++k) t->w[k] = wh[k];
  op->rfused = t;
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
  cudaFree(t->at_h); cudaFree(t->at_w);
  delete t;
  op->rfused = nullptr;
}

int resize_fused_guidance(const dps_operator* op, const dps_source& src, const float* y, int64_t y_stride, float* r_out,
                          float* g, int64_t g_stride, float* partials, int n, cudaStream_t st) {
  const ResizeFused& t = *op->rfused;
  DPS_REQUIRE(src.eps, DPS_ERR_INVALID, "resize guidance: the fused kernel forms x̂₀ from x and ε (eps is required)");
  FusedArgs a;
  for (int k = 0; k < 32; ++k) a.w[k] = t.w[k];
  a.at_h = t.at_h;
  a.at_w = t.at_w;
  a.C = op->C;
  a.src = src;
  a.y = y;
  a.y_stride = y_stride;
  a.r_out = r_out;
  a.g = g;
  a.g_stride = g_stride;
  a.partials = partials;
  dim3 grid((unsigned)((int64_t)op->C * n * kCluster));
  if (t.F == 4)
    resize_guidance_kernel<4><<<grid, kT, fused_smem<4>(), st>>>(a);
  else
    resize_guidance_kernel<8><<<grid, kT, fused_smem<8>(), st>>>(a);
  DPS_LAUNCH_CHECK("resize_guidance");
  return DPS_OK;
}
