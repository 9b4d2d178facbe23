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
//   0. x rows → shared memory by TMA bulk copies in four 8-row chunks, ε rows → registers (a thread owns a column); each
//      chunk becomes x̂₀ (clamped) in place as it lands; the clamp mask of a thread's column is one 32-bit register; the
//      first / last HALO rows are PUSHED into the neighbour CTAs' halo buffers (st.async over DSMEM, completing bytes on
//      the neighbour's mbarrier) as they are produced.
//   1. H pass: a thread owns a column and walks the 32 + 2·HALO rows its RJ residual rows need — own rows first, then
//      (after the wait on its own halo barrier) the pushed rows, mirrored own rows at the image border.
//   2. W pass from the column-padded RJ×(256+2·HALO) tile with 128-bit shared loads; r = y − (·), Σr², Σ|r| → one
//      partial-sum pair per CTA; r stays in shared memory; the first / last two residual rows are pushed to the neighbours.
//   3. Aᵀ: u = r·A_w for the RJ+4 residual rows that touch this CTA's image rows (own rows first, then the 2+2 pushed
//      ones), then g = A_hᵀ u, masked, one coalesced store per image row.
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
#ifndef DPS_RSF_CHUNKS
#define DPS_RSF_CHUNKS 4
#endif
constexpr int kW = 256, kRI = 32, kCluster = 8, kT = 256, kChunks = DPS_RSF_CHUNKS, kChunkRows = kRI / kChunks;

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
constexpr size_t fused_smem() {  // x rows + halo rows (the H-pass tile aliases them) + residual rows + neighbour rows + A_hᵀ band + scratch + mbarriers
  return sizeof(float) * ((size_t)kRI * kW + (size_t)2 * Geo<F>::HALO * kW + (size_t)Geo<F>::RJ * Geo<F>::OW + (size_t)4 * Geo<F>::OW + (size_t)kRI * 4 + 64) +
         8 * (kChunks + 2);
}
template <int F>
constexpr size_t project_smem() {  // the same plus the measurement rows (kept beside A·data for the up-sampling epilogue)
  return fused_smem<F>() + sizeof(float) * (size_t)Geo<F>::RJ * Geo<F>::OW + 16;
}

#ifdef DPS_RSF_TRACE  // experiment builds only (tools/build_variant.sh): per-CTA phase timestamps of the first 4096 CTAs
__device__ long long rsf_trace[4096 * 16];
#define RSF_T(i) do { if (threadIdx.x == 0 && blockIdx.x < 4096) rsf_trace[blockIdx.x * 16 + (i)] = clock64(); } while (0)
#else
#define RSF_T(i) do { } while (0)
#endif
DPS_DEV unsigned mapa_u32(unsigned addr, unsigned rank) {
  unsigned r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
// PROJ = false: the guidance kernel described above.
// PROJ = true : SuperResolutionOperator.project / ortho_project (measurements.py:48-50, :90-91) in one launch:
//               out = (data − up(A·data)) + up(y), up = the reference's transpose = nearest-neighbour ×F (F.interpolate).
//               Passes 0-2 are shared (no ε, no clamp: src.x is the data itself); A·data and y stay in shared memory and the
//               epilogue writes the image rows.  Same A·data bits as the guidance kernel, same two roundings as torch's
//               `data - up(..) + up(..)`.  y = null gives ortho_project (… + 0).
//
// Neighbour exchange = PUSH (st.async, SASS `STAS`): a CTA writes the x̂₀ rows its neighbours' windows need straight into THEIR
// halo buffers as it produces them (pass 0), and later its first / last two residual rows into their Sn buffers (W pass).  Every
// such store completes bytes on an mbarrier in the DESTINATION CTA, which waits on its own barrier like for a TMA copy and
// then reads LOCAL shared memory.  Compared with the pull form of the first versions (cluster barrier, then
// ld.shared::cluster from the neighbour):
//   * no cluster-scope release: `fence.acq_rel.cluster` is MEMBAR.ALL.GPU + CCTL.IVALL in SASS, twice per CTA, on the
//     critical path of an 8-CTA hardware barrier that waits for the slowest of the eight;
//   * a CTA depends on its two neighbours only; nothing waits for "everybody has finished reading" before it may exit;
//   * DSMEM moves ≈17–21 B per cycle and SM: pulled rows were latency AND bandwidth exposed in front of the H pass
//     (≈1 500 cycles of a 13 000-cycle CTA at N = 8, tools/rsf_trace.py); pushed rows travel while pass 0 is still converting.
// One cluster barrier remains, at the very top (arrive after the mbarrier init, wait before the first push): a neighbour must
// not complete bytes on a barrier that is not initialised yet.  A CTA leaves only after both of its barriers have
// completed, i.e. after every store aimed at it has landed.
DPS_DEV void st_async_f32(unsigned remote_addr, float v, unsigned remote_bar) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.f32 [%0], %1, [%2];" ::"r"(remote_addr), "f"(v), "r"(remote_bar)
               : "memory");
}

template <int F, bool PROJ>
__global__ void __cluster_dims__(kCluster, 1, 1) __launch_bounds__(kT, F == 4 ? 4 : 3) resize_guidance_kernel(const __grid_constant__ FusedArgs a) {
  using G = Geo<F>;
  constexpr int TAPS = G::TAPS, HALO = G::HALO, RJ = G::RJ, OW = G::OW, RU = G::RU, PADW = G::PADW, H = kRI * kCluster;
  static_assert(2 * HALO * kW >= RJ * PADW, "the H-pass tile aliases the halo buffer");
  static_assert(HALO <= kRI && 2 <= RJ, "a window reaches one neighbour only");
  extern __shared__ __align__(16) float smem[];
  float* Sx = smem;                       // (32, 256)  x → x̂₀ (clamped)
  float* Sh = Sx + kRI * kW;              // (2·HALO, 256) x̂₀ of the HALO rows above and the HALO rows below mine, PUSHED by the neighbours
  float* St = Sh;                         // (RJ, PADW) H-pass result, column-padded by mirroring — aliases Sh (dead by then)
  float* Sr = Sh + 2 * HALO * kW;         // (RJ, OW)   residual rows of this CTA
  float* Sn = Sr + RJ * OW;               // (4, OW)    residual rows RJ·q − 2, RJ·q − 1, RJ·(q+1), RJ·(q+1) + 1, pushed by the neighbours
  float* Ah = Sn + 4 * OW;                // (32, 4)    transposed H band of my image rows
  float* red = Ah + kRI * 4;              // 64
  uint64_t* bar = reinterpret_cast<uint64_t*>(red + 64);  // kChunks TMA barriers, then hbar (halo rows), nbar (neighbour residual rows)
  uint64_t* hbar = bar + kChunks;
  uint64_t* nbar = hbar + 1;
  float* Sy = reinterpret_cast<float*>(bar + kChunks + 2);  // PROJ only: (RJ, OW) measurement rows

  cg::cluster_group cluster = cg::this_cluster();
  const int q = (int)cluster.block_rank();
  const int plane = blockIdx.x / kCluster, c = plane % a.C, n = plane / a.C;
  const int tid = threadIdx.x;
  const bool has_up = q > 0, has_dn = q < kCluster - 1;  // image borders: the halo is my own rows, mirrored
  const int64_t poff = (int64_t)c * H * kW + (int64_t)q * kRI * kW;
  const float* xg = a.src.x + n * a.src.x_stride + poff;
  const float* eg = PROJ ? nullptr : a.src.eps + n * a.src.eps_stride + poff;

  RSF_T(0);
  if (tid == 0) {
#pragma unroll
    for (int ch = 0; ch < kChunks; ++ch) mbar_init(bar + ch, 1);
    mbar_init(hbar, 1);
    mbar_init(nbar, 1);
    mbar_init_fence();
    // my single arrival on the two neighbour barriers, with the bytes the neighbours will complete
    mbar_expect_tx(hbar, ((has_up ? 1u : 0u) + (has_dn ? 1u : 0u)) * HALO * kW * (unsigned)sizeof(float));
    if (!PROJ) mbar_expect_tx(nbar, ((has_up ? 1u : 0u) + (has_dn ? 1u : 0u)) * 2 * OW * (unsigned)sizeof(float));
  }
  asm volatile("barrier.cluster.arrive.relaxed.aligned;" ::: "memory");  // (thread 0: after its init fence)
  // ε never passes through shared memory: a thread owns a column, so a row of ε is one coalesced 1 KB request per CTA; all 32
  // are in flight before the first x chunk has landed (registers are cheap here: the accumulators are not live yet).
  float ev[kRI];
  if (!PROJ) {
#pragma unroll
    for (int r = 0; r < kRI; ++r) ev[r] = ldg_stream(eg + r * kW + tid);
  }
  __syncthreads();
  if (tid == 0) {
    constexpr unsigned bytes = kChunkRows * kW * sizeof(float);
#pragma unroll
    for (int ch = 0; ch < kChunks; ++ch) {
      mbar_expect_tx(bar + ch, bytes);
      bulk_load(Sx + ch * kChunkRows * kW, xg + ch * kChunkRows * kW, bytes, bar + ch);
    }
  }
  // tables and measurement values while the rows are in flight
  float4 aw = make_float4(0.f, 0.f, 0.f, 0.f);
  if (!PROJ) {
    stage_async(Ah, a.at_h + (size_t)q * kRI * 4, kRI * 4, tid, kT);
    aw = __ldg(reinterpret_cast<const float4*>(a.at_w) + tid);  // my column's transposed W band
    if (tid < 4 * OW && (tid < 2 * OW ? !has_up : !has_dn)) Sn[tid] = 0.f;  // residual rows beyond the image do not exist
  }
  constexpr int kRP = (RJ * OW + kT - 1) / kT;  // residual values per thread in the W pass
  float yv[kRP];
  const float* yp = a.y ? a.y + n * a.y_stride + (int64_t)c * (H / F) * OW + (int64_t)q * RJ * OW : nullptr;
#pragma unroll
  for (int u = 0; u < kRP; ++u) {
    const int o = tid + u * kT;
    yv[u] = (yp && o < RJ * OW) ? ldg_ro(yp + o) : 0.f;
  }
  stage_wait();
  // where my rows go: rows 0 … HALO−1 are the DOWN halo of CTA q−1, rows 32−HALO … 31 the UP halo of CTA q+1
  const unsigned sh_mine = smem_u32(Sh) + tid * 4, hb_mine = smem_u32(hbar);
  const unsigned up_dst = has_up ? mapa_u32(sh_mine, (unsigned)(q - 1)) + HALO * kW * 4 : 0u, up_bar = has_up ? mapa_u32(hb_mine, (unsigned)(q - 1)) : 0u;
  const unsigned dn_dst = has_dn ? mapa_u32(sh_mine, (unsigned)(q + 1)) : 0u, dn_bar = has_dn ? mapa_u32(hb_mine, (unsigned)(q + 1)) : 0u;
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");  // every CTA of the cluster has initialised its barriers
  RSF_T(1);
  // ---- 0. x̂₀ in place, chunk by chunk as the copies land; clamp mask of my column → one register; boundary rows → neighbours ----
  unsigned pass_bits = 0;
  {
    const float lo = a.src.clip ? -1.0f : -INFINITY, hi = a.src.clip ? 1.0f : INFINITY;
#pragma unroll
    for (int ch = 0; ch < kChunks; ++ch) {
      mbar_wait(bar + ch, 0);
#pragma unroll
      for (int rr = 0; rr < kChunkRows; ++rr) {
        const int r = ch * kChunkRows + rr;
        float v = Sx[r * kW + tid];
        if (!PROJ) {
          const float pre = x0_pre(v, ev[r], a.src.c1, a.src.c2);
          v = fminf(fmaxf(pre, lo), hi);
          Sx[r * kW + tid] = v;
          pass_bits |= (pre >= lo && pre <= hi) ? (1u << r) : 0u;
        }
        if (r < HALO && has_up) st_async_f32(up_dst + r * kW * 4, v, up_bar);
        if (r >= kRI - HALO && has_dn) st_async_f32(dn_dst + (r - (kRI - HALO)) * kW * 4, v, dn_bar);
      }
    }
  }
  RSF_T(2);

  // ---- 1. H pass: t[jj][col] = Σ_k w[k] · x̂₀[sym(F·(RJ·q + jj) − HALO + k)][col] ----
  {
    float acc[RJ];
#pragma unroll
    for (int jj = 0; jj < RJ; ++jj) acc[jj] = 0.f;
    // Before the halo rows are needed: every term that uses only my own rows and comes FIRST in its accumulator (the order of
    // the additions stays k = 0, 1, …: bit-identical to the two-kernel path) — residual rows whose window starts inside my
    // rows, up to my last row; 69 % of the FMAs at F = 4.  A thread reads only the column it wrote: no barrier.
#pragma unroll
    for (int lr = 0; lr < kRI; ++lr) {
      const float v = Sx[lr * kW + tid];
#pragma unroll
      for (int jj = 0; jj < RJ; ++jj) {
        const int k = lr + HALO - F * jj;  // compile-time after unrolling
        if (F * jj - HALO >= 0 && k >= 0 && k < TAPS) acc[jj] = fmaf(a.w[k], v, acc[jj]);
      }
    }
    RSF_T(3);
    mbar_wait_guarded(hbar, 0);  // the neighbours' rows have landed in Sh
    RSF_T(4);
    // halo rows: pushed by the neighbour or, at the image border, my own rows mirrored WITH edge repeat
    // (−1 ↦ 0, −2 ↦ 1 …; 32 ↦ 31, 33 ↦ 30 …) — one base and one signed row stride each
    const float* pu = has_up ? Sh + tid : Sx + (HALO - 1) * kW + tid;
    const int su = has_up ? kW : -kW;
    const float* pd = has_dn ? Sh + HALO * kW + tid : Sx + (kRI - 1) * kW + tid;
    const int sd = has_dn ? kW : -kW;
#pragma unroll
    for (int wdx = 0; wdx < kRI + 2 * HALO; ++wdx) {
      const int lr = wdx - HALO;  // row relative to my first image row
      bool used = false;
#pragma unroll
      for (int jj = 0; jj < RJ; ++jj) {
        const int k = wdx - F * jj;
        used = used || (k >= 0 && k < TAPS && !(F * jj - HALO >= 0 && lr < kRI));
      }
      if (!used) continue;
      const float v = lr < 0 ? pu[wdx * su] : (lr >= kRI ? pd[(lr - kRI) * sd] : Sx[lr * kW + tid]);
#pragma unroll
      for (int jj = 0; jj < RJ; ++jj) {
        const int k = wdx - F * jj;
        if (k >= 0 && k < TAPS && !(F * jj - HALO >= 0 && lr < kRI)) acc[jj] = fmaf(a.w[k], v, acc[jj]);
      }
    }
    __syncthreads();  // St aliases Sh: every thread has read its halo column
#pragma unroll
    for (int jj = 0; jj < RJ; ++jj) {
      float* row = St + jj * PADW + HALO;
      row[tid] = acc[jj];
      if (tid < HALO) row[-1 - tid] = acc[jj];                   // columns −1, −2, … mirror columns 0, 1, …
      if (tid >= kW - HALO) row[2 * kW - 1 - tid] = acc[jj];     // columns 256, 257, … mirror 255, 254, …
    }
  }
  __syncthreads();

  RSF_T(5);
  // ---- 2. W pass, residual, partial sums; my first / last two residual rows → the neighbours' Sn ----
  float sq = 0.f, ab = 0.f;
  {
    const unsigned sn_mine = smem_u32(Sn), nb_mine = smem_u32(nbar);
    // my rows 0, 1 are rows RJ, RJ+1 of CTA q−1 (its Sn rows 2, 3); my rows RJ−2, RJ−1 are rows −2, −1 of CTA q+1 (its Sn rows 0, 1)
    const unsigned nup_dst = (!PROJ && has_up) ? mapa_u32(sn_mine, (unsigned)(q - 1)) + 2 * OW * 4 : 0u;
    const unsigned nup_bar = (!PROJ && has_up) ? mapa_u32(nb_mine, (unsigned)(q - 1)) : 0u;
    const unsigned ndn_dst = (!PROJ && has_dn) ? mapa_u32(sn_mine, (unsigned)(q + 1)) : 0u;
    const unsigned ndn_bar = (!PROJ && has_dn) ? mapa_u32(nb_mine, (unsigned)(q + 1)) : 0u;
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
        const float res = (yp && !PROJ) ? yv[u] - acc : acc;
        Sr[o] = res;
        if (PROJ) Sy[o] = yv[u];
        if (!PROJ) {
          if (jj < 2 && has_up) st_async_f32(nup_dst + o * 4, res, nup_bar);
          if (jj >= RJ - 2 && has_dn) st_async_f32(ndn_dst + (o - (RJ - 2) * OW) * 4, res, ndn_bar);
          if (a.r_out) a.r_out[((int64_t)n * a.C + c) * (H / F) * OW + (int64_t)q * RJ * OW + o] = res;
        }
        sq = fmaf(res, res, sq);
        ab += fabsf(res);
      }
    }
  }
  RSF_T(6);
  if (!PROJ && a.partials) {  // per-warp Σr², Σ|r| now (the values die here); warp 0 adds the eight behind the barrier below
    sq = warp_sum(sq);
    ab = warp_sum(ab);
    if ((tid & 31) == 0) {
      red[tid >> 5] = sq;
      red[32 + (tid >> 5)] = ab;
    }
  }
  __syncthreads();  // my residual rows (and the per-warp sums) are complete
  if (PROJ) {
    float* op_ = a.g + n * a.g_stride + poff;
    const int l = tid / F;
#pragma unroll
    for (int ii = 0; ii < kRI; ++ii) {
      const int o = (ii / F) * OW + l;
      stg_stream(op_ + ii * kW + tid, __fadd_rn(__fsub_rn(Sx[ii * kW + tid], Sr[o]), Sy[o]));
    }
    return;
  }
  if (a.partials && tid < 32) {  // the second half of block_sum2 (same order, same bits) by warp 0 alone: no CTA barrier, the
    float s2 = tid < kT / 32 ? red[tid] : 0.0f, a2 = tid < kT / 32 ? red[32 + tid] : 0.0f;  // other warps go straight on to Aᵀ
    s2 = warp_sum(s2);
    a2 = warp_sum(a2);
    if (tid == 0) {
      float* pp = a.partials + ((int64_t)n * (a.C * kCluster) + c * kCluster + q) * 2;
      pp[0] = s2;
      pp[1] = a2;
    }
  }
  // ---- 3. Aᵀ: u[m][col] = Σ_d aw[d] · r[RJ·q − 2 + m][l0(col) + d],  then g[ii][col] = mask · Σ_d Ah[ii][d] · u[m0(ii) + d] ----
  {
    // first residual column that touches image column `tid`:  ceil((tid + HALO − TAPS + 1)/F)
    const int num = tid + HALO - TAPS + 1;
    const int lfirst = num >= 0 ? (num + F - 1) / F : -((-num) / F);
    const float awv[4] = {aw.x, aw.y, aw.z, aw.w};
    float uu[RU];
    auto urow = [&](const float* rr) {
      float s = 0.f;
#pragma unroll
      for (int d = 0; d < 4; ++d) {
        const int l = lfirst + d;
        if (l >= 0 && l < OW) s = fmaf(awv[d], rr[l], s);
      }
      return s;
    };
#pragma unroll
    for (int m = 2; m < RJ + 2; ++m) uu[m] = urow(Sr + (m - 2) * OW);  // my own rows first
    RSF_T(7);
    mbar_wait_guarded(nbar, 0);  // the neighbours' two rows each have landed in Sn
    RSF_T(8);
#pragma unroll
    for (int m = 0; m < 2; ++m) {
      uu[m] = urow(Sn + m * OW);
      uu[RJ + 2 + m] = urow(Sn + (2 + m) * OW);
    }
    RSF_T(9);
    float* gp = a.g + n * a.g_stride + poff;
#pragma unroll
    for (int ii = 0; ii < kRI; ++ii) {
      const int m0 = G::j0(ii) + 2;  // first residual row, relative to RJ·q − 2, that touches image row ii (compile-time)
      float s = 0.f;
#pragma unroll
      for (int d = 0; d < 4; ++d)
        if (m0 + d >= 0 && m0 + d < RU) s = fmaf(Ah[ii * 4 + d], uu[m0 + d], s);
      stg_stream(gp + ii * kW + tid, ((pass_bits >> ii) & 1u) ? s : 0.f);
    }
    RSF_T(10);
    RSF_T(11);
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

#ifdef DPS_RSF_TRACE
extern "C" int dps_debug_rsf_trace(long long* host_out) {
  return (int)cudaMemcpyFromSymbol(host_out, rsf_trace, sizeof(long long) * 4096 * 16);
}
#endif

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
  if (F == 4) {
    DPS_CUDA(cudaFuncSetAttribute(resize_guidance_kernel<4, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fused_smem<4>()));
    DPS_CUDA(cudaFuncSetAttribute(resize_guidance_kernel<4, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)project_smem<4>()));
  } else {
    DPS_CUDA(cudaFuncSetAttribute(resize_guidance_kernel<8, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fused_smem<8>()));
    DPS_CUDA(cudaFuncSetAttribute(resize_guidance_kernel<8, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)project_smem<8>()));
  }
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
    resize_guidance_kernel<4, false><<<grid, kT, fused_smem<4>(), st>>>(a);
  else
    resize_guidance_kernel<8, false><<<grid, kT, fused_smem<8>(), st>>>(a);
  DPS_LAUNCH_CHECK("resize_guidance");
  return DPS_OK;
}

// out = (data − up(A·data)) + up(y)   (y null: ortho_project).  One launch of the cluster kernel in PROJ mode.
int resize_fused_project(const dps_operator* op, const float* data, int64_t data_stride, const float* y, int64_t y_stride, float* out,
                         int64_t out_stride, int n, cudaStream_t st) {
  const ResizeFused& t = *op->rfused;
  FusedArgs a;
  for (int k = 0; k < 32; ++k) a.w[k] = t.w[k];
  a.at_h = t.at_h;
  a.at_w = t.at_w;
  a.C = op->C;
  a.src = dps_source{};
  a.src.x = data;
  a.src.x_stride = data_stride;
  a.y = y;
  a.y_stride = y_stride;
  a.r_out = nullptr;
  a.g = out;
  a.g_stride = out_stride;
  a.partials = nullptr;
  dim3 grid((unsigned)((int64_t)op->C * n * kCluster));
  if (t.F == 4)
    resize_guidance_kernel<4, true><<<grid, kT, project_smem<4>(), st>>>(a);
  else
    resize_guidance_kernel<8, true><<<grid, kT, project_smem<8>(), st>>>(a);
  DPS_LAUNCH_CHECK("resize_project");
  return DPS_OK;
}
