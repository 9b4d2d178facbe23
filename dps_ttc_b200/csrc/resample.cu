// Per-particle reductions, guidance coefficients, particle reweighting and resampling
// (SURVEY.md §8 rows A9-A11, B1 and the "north-star items with no reference counterpart").
//
// Everything here is O(N) or a particle-sized copy:
//   * norms / coefficients : one warp per particle reduces the P per-CTA partial sums in fp64 with a
//     fixed order → the value of a particle's norm never depends on N or on how particles are sharded;
//   * weights + CDF        : one CTA; max by warp shuffles (log-sum-exp shift), then ONE thread accumulates
//     the CDF sequentially in fp32 in index order and divides by the fp32 total — the arithmetic of
//     torch.multinomial's CPU kernel (scalar_t accumulators; pinned in tests/test_oracle_pins.py), so the
//     ancestor indices are bit-identical to the reference's given the same weights and uniforms;
//   * ancestors            : inverse-CDF binary search ("first j with cdf[j] >= u");
//   * gather / broadcast   : 128-bit streaming copies, 2T bytes per particle, HBM-bound.
#include <float.h>

#include "common.cuh"

namespace {

__global__ void particle_norms_kernel(const float* __restrict__ partials, int P, int n, int mode, float scale,
                                      float* __restrict__ l2, float* __restrict__ l1, float* __restrict__ coef) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= n) return;
  const float* p = partials + (int64_t)warp * P * 2;
  double sq = 0.0, ab = 0.0;
  for (int i = lane; i < P; i += 32) {
    sq += (double)p[2 * i];
    ab += (double)p[2 * i + 1];
  }
  sq = warp_sum(sq);
  ab = warp_sum(ab);
  if (lane == 0) {
    const float nrm = (float)sqrt(sq);
    if (l2) l2[warp] = nrm;
    if (l1) l1[warp] = (float)ab;
    if (coef) {
      // ∇‖r‖ = −Aᵀr/‖r‖ (mode 1);  ∇‖r‖² = −2Aᵀr (mode 2).  0/0 → 0 like torch's norm backward at 0.
      float c = mode == DPS_COEF_NORM ? (nrm > 0.f ? -scale / nrm : 0.f) : -2.0f * scale;
      coef[warp] = c;
    }
  }
}

// Poisson-likelihood guidance (condition_methods.py:50-55): the loss is the Frobenius norm over ALL particles times
// mean(1/|y|), so every particle gets the same coefficient −scale/‖r‖_all (scale = ζ·mean(1/|y|), folded by the host).
// One CTA: per-particle norms like particle_norms_kernel, then the global sum in a fixed order.
__global__ void __launch_bounds__(256) global_norm_coef_kernel(const float* __restrict__ partials, int P, int n, float scale,
                                                               float* __restrict__ l2, float* __restrict__ coef) {
  __shared__ double red[8];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  double tot = 0.0;
  for (int i = warp; i < n; i += 8) {  // a warp per particle
    const float* p = partials + (int64_t)i * P * 2;
    double sq = 0.0;
    for (int j = lane; j < P; j += 32) sq += (double)p[2 * j];
    sq = warp_sum(sq);
    if (lane == 0 && l2) l2[i] = (float)sqrt(sq);
    tot += sq;
  }
  if (lane == 0) red[warp] = tot;
  __syncthreads();
  double all = 0.0;
#pragma unroll
  for (int w = 0; w < 8; ++w) all += red[w];
  const float nrm = (float)sqrt(all);
  const float c = nrm > 0.f ? -scale / nrm : 0.f;
  for (int i = threadIdx.x; i < n; i += 256) coef[i] = c;
}

__global__ void logweights_kernel(const float* __restrict__ meas, const float* __restrict__ sem, int n, float tau,
                                  float meas_scale, int meas_pow, float sem_scale, int sem_pow,
                                  float* __restrict__ logw) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float m = meas[i];
  if (meas_pow == 2) m = m * m;
  float cost = meas_scale * m;
  if (sem) {
    float s = sem[i];
    if (sem_pow == 2) s = s * s;
    cost += sem_scale * s;
  }
  logw[i] = -tau * cost;
}

// one CTA of 1024 threads
__global__ void __launch_bounds__(1024) weights_cdf_kernel(const float* __restrict__ logw, int n, int linear_mode,
                                                           float* __restrict__ weights_out, float* __restrict__ cdf,
                                                           float* __restrict__ lse_out, int32_t* __restrict__ degenerate) {
  __shared__ float s_max[32], s_min[32];
  __shared__ float s_m, s_mn;
  __shared__ float s_total;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  float mx = -FLT_MAX, mn = FLT_MAX;
  for (int i = tid; i < n; i += blockDim.x) {
    const float v = logw[i];
    mx = fmaxf(mx, v);
    mn = fminf(mn, v);
  }
  for (int o = 16; o > 0; o >>= 1) {
    mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, o));
  }
  if (lane == 0) { s_max[warp] = mx; s_min[warp] = mn; }
  __syncthreads();
  if (warp == 0) {
    mx = s_max[lane];  // blockDim = 1024 → 32 warps
    mn = s_min[lane];
    for (int o = 16; o > 0; o >>= 1) {
      mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
      mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, o));
    }
    if (lane == 0) { s_m = mx; s_mn = mn; }
  }
  __syncthreads();
  const float shift = linear_mode ? 0.0f : s_m;
  // unnormalised fp32 weights, exactly what the reference hands to torch.multinomial
  for (int i = tid; i < n; i += blockDim.x) weights_out[i] = expf(logw[i] - shift);
  __syncthreads();
  if (tid == 0) {
    // sequential fp32 cumulative sum in index order (torch CPU multinomial: sum += val; cum_dist[j] = sum)
    float sum = 0.0f;
    float wmax = -FLT_MAX, wmin = FLT_MAX;
    for (int i = 0; i < n; ++i) {
      const float w = weights_out[i];
      wmax = fmaxf(wmax, w);
      wmin = fminf(wmin, w);
      sum = __fadd_rn(sum, w);
      cdf[i] = sum;
    }
    s_total = sum;
    const bool bad = !(sum > 0.0f) || !isfinite(sum) || (wmax == wmin);
    *degenerate = bad ? 1 : 0;
    if (lse_out) *lse_out = shift + logf(sum);
  }
  __syncthreads();
  const float total = s_total;
  const bool ok = total > 0.0f && isfinite(total);
  for (int i = tid; i < n; i += blockDim.x) {
    if (ok) {
      cdf[i] = (i == n - 1) ? 1.0f : __fdiv_rn(cdf[i], total);  // cum_dist[j] /= sum
      weights_out[i] = __fdiv_rn(weights_out[i], total);
    } else {
      cdf[i] = (float)(i + 1) / (float)n;
      weights_out[i] = 1.0f / (float)n;
    }
  }
}

DPS_DEV int64_t search_cdf(const float* cdf, int n, double u) {
  // first j with cdf[j] >= u  (torch: while: if cum_dist[mid] < u → left = mid+1 else right = mid)
  int left = 0, right = n;
  while (right - left > 0) {
    const int mid = left + (right - left) / 2;
    if ((double)cdf[mid] < u) left = mid + 1; else right = mid;
  }
  return left < n ? left : n - 1;
}

__global__ void ancestors_kernel(const float* __restrict__ cdf, int n, const double* __restrict__ u, int n_draws,
                                 int systematic, const int32_t* __restrict__ degenerate,
                                 int64_t* __restrict__ ancestors) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_draws) return;
  if (degenerate && *degenerate) {
    ancestors[i] = i < n ? i : n - 1;
    return;
  }
  const double ui = systematic ? ((double)i + u[0]) / (double)n_draws : u[i];
  ancestors[i] = search_cdf(cdf, n, ui);
}

constexpr int kCopyThreads = 256;
constexpr int kCopyVec = 2;  // float4 per thread; measured against 1 and 4: 29.5 vs 31.6 / 31.8 µs at N = 128, equal below

// dst[i] = src[idx(i)], elems4 float4 per particle; grid (chunks, n_dst)
__global__ void __launch_bounds__(kCopyThreads) gather_kernel(const float* __restrict__ src,
                                                              const int64_t* __restrict__ ancestors, int broadcast,
                                                              float* __restrict__ dst, int64_t elems4) {
  const int i = blockIdx.y;
  const int64_t a = broadcast ? ancestors[0] : ancestors[i];
  const float* s = src + a * elems4 * 4;
  float* d = dst + (int64_t)i * elems4 * 4;
  const int64_t base = (int64_t)blockIdx.x * (kCopyThreads * kCopyVec) + threadIdx.x;
  float4 v[kCopyVec];
#pragma unroll
  for (int u = 0; u < kCopyVec; ++u) {
    const int64_t j = base + (int64_t)u * kCopyThreads;
    if (j < elems4) v[u] = ldg_stream4(s + j * 4);
  }
#pragma unroll
  for (int u = 0; u < kCopyVec; ++u) {
    const int64_t j = base + (int64_t)u * kCopyThreads;
    if (j < elems4) stg_stream4(d + j * 4, v[u]);
  }
}

// dst[i] = particle `ancestors[i]` read straight from its owner's memory: rank = a / n_per_rank, slot = a % n_per_rank.
// peer_bases[] are device pointers that are valid in THIS process for every rank's particle buffer (CUDA IPC /
// symmetric memory over NVLink); loads to a peer go over NVSwitch, loads to the own rank stay in local HBM.
__global__ void __launch_bounds__(kCopyThreads) gather_p2p_kernel(const float* const* __restrict__ peer_bases,
                                                                  int n_per_rank,
                                                                  const int64_t* __restrict__ ancestors,
                                                                  float* __restrict__ dst, int64_t elems4) {
  const int i = blockIdx.y;
  const int64_t a = ancestors[i];
  const int owner = (int)(a / n_per_rank);
  const float4* s = reinterpret_cast<const float4*>(peer_bases[owner]) + (a - (int64_t)owner * n_per_rank) * elems4;
  float* d = dst + (int64_t)i * elems4 * 4;
  const int64_t base = (int64_t)blockIdx.x * (kCopyThreads * kCopyVec) + threadIdx.x;
  float4 v[kCopyVec];
#pragma unroll
  for (int u = 0; u < kCopyVec; ++u) {
    const int64_t j = base + (int64_t)u * kCopyThreads;
    if (j < elems4) v[u] = s[j];  // plain ld.global: peer memory is not cached in the local L2
  }
#pragma unroll
  for (int u = 0; u < kCopyVec; ++u) {
    const int64_t j = base + (int64_t)u * kCopyThreads;
    if (j < elems4) stg_stream4(d + j * 4, v[u]);
  }
}

// ---- the same with the inter-GPU rendezvous INSIDE the kernel (no barrier launches around it) -------------------------
// signal_pads[r] points (in this process) at rank r's pad of `world` uint32 words; rank q announces "my particle buffer of
// exchange `epoch` is complete" by storing `epoch` into word q of EVERY rank's pad (release, system scope: the stores of the
// kernel that produced the particles — an earlier kernel of the same stream — are ordered before it).  A CTA then only
// waits for the OWNER of the particle it copies, so copies from ranks that are ready start while others are still arriving.
// CTA (0,0) additionally waits for every peer: when this kernel has finished on rank A, all ranks have started exchange
// `epoch`, hence finished exchange `epoch − 1` (stream order).  With the particle buffers double-buffered by epoch parity
// (slot_elems selects the half) a buffer is therefore never overwritten while a peer still reads it, and no trailing
// barrier is needed.  Epochs only grow (wrap-safe signed comparison); a peer that never arrives traps instead of hanging.
DPS_DEV unsigned ld_acquire_sys(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
DPS_DEV void st_release_sys(unsigned* p, unsigned v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
DPS_DEV float4 ld_relaxed_sys4(const float4* p) {
  float4 r;
  asm volatile("ld.relaxed.sys.global.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
  return r;
}
DPS_DEV void wait_epoch(const unsigned* word, unsigned epoch) {
  unsigned spins = 0;
  while ((int)(ld_acquire_sys(word) - epoch) < 0) {
    if (++spins > (1u << 26)) __trap();  // a peer never announced this exchange: fail the launch instead of hanging the GPU
    __nanosleep(64);
  }
}

__global__ void __launch_bounds__(kCopyThreads) gather_p2p_sync_kernel(const float* const* __restrict__ peer_bases,
                                                                       unsigned* const* __restrict__ signal_pads, int rank,
                                                                       int world, unsigned epoch, int64_t slot_elems,
                                                                       int n_per_rank,
                                                                       const int64_t* __restrict__ ancestors,
                                                                       float* __restrict__ dst, int64_t elems4) {
  const int i = blockIdx.y;
  const bool lead = blockIdx.x == 0 && blockIdx.y == 0;
  if (lead && threadIdx.x < world) {
    __threadfence_system();
    st_release_sys(signal_pads[threadIdx.x] + rank, epoch);
  }
  const int64_t a = ancestors[i];
  const int owner = (int)(a / n_per_rank);
  const unsigned* mine = signal_pads[rank];
  if (threadIdx.x == 0) wait_epoch(mine + owner, epoch);
  __syncthreads();
  const float4* s = reinterpret_cast<const float4*>(peer_bases[owner] + slot_elems) + (a - (int64_t)owner * n_per_rank) * elems4;
  float* d = dst + (int64_t)i * elems4 * 4;
  const int64_t base = (int64_t)blockIdx.x * (kCopyThreads * kCopyVec) + threadIdx.x;
  float4 v[kCopyVec];
#pragma unroll
  for (int u = 0; u < kCopyVec; ++u) {
    const int64_t j = base + (int64_t)u * kCopyThreads;
    if (j < elems4) v[u] = s[j];  // plain ld.global after the acquire above: this kernel has not touched these lines before
                                  // (L1 is empty of them) and peer memory is not cached in the local L2
  }
#pragma unroll
  for (int u = 0; u < kCopyVec; ++u) {
    const int64_t j = base + (int64_t)u * kCopyThreads;
    if (j < elems4) stg_stream4(d + j * 4, v[u]);
  }
  if (lead && threadIdx.x < world) wait_epoch(mine + threadIdx.x, epoch);
}

__global__ void __launch_bounds__(1024) argmin_kernel(const float* __restrict__ costs, int n, int64_t* __restrict__ best,
                                                      float* __restrict__ best_cost) {
  __shared__ float s_v[32];
  __shared__ int s_i[32];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  float bv = FLT_MAX;
  int bi = 0x7fffffff;
  for (int i = tid; i < n; i += blockDim.x) {
    const float v = costs[i];
    // first minimum wins (torch.argmin); NaN is treated as smaller than everything, like torch
    if (v < bv || (v == bv && i < bi) || (v != v && !(bv != bv))) { bv = v; bi = i; }
  }
  auto better = [](float v, int i, float bv, int bi) {
    const bool vn = v != v, bn = bv != bv;
    if (vn != bn) return vn;
    if (vn && bn) return i < bi;
    return v < bv || (v == bv && i < bi);
  };
  for (int o = 16; o > 0; o >>= 1) {
    const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
    const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
    if (better(ov, oi, bv, bi)) { bv = ov; bi = oi; }
  }
  if (lane == 0) { s_v[warp] = bv; s_i[warp] = bi; }
  __syncthreads();
  if (warp == 0) {
    bv = s_v[lane];
    bi = s_i[lane];
    for (int o = 16; o > 0; o >>= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (better(ov, oi, bv, bi)) { bv = ov; bi = oi; }
    }
    if (lane == 0) {
      best[0] = bi;
      if (best_cost) best_cost[0] = bv;
    }
  }
}

int launch_gather(const float* src, const int64_t* idx, int broadcast, float* dst, int n_dst, int64_t elems,
                  dps_stream_t stream, const char* who) {
  DPS_REQUIRE(src && idx && dst && n_dst > 0 && n_dst <= 65535 && elems > 0, DPS_ERR_INVALID, "%s: bad arguments", who);
  DPS_REQUIRE(elems % 4 == 0, DPS_ERR_UNSUPPORTED, "%s: elems must be a multiple of 4", who);
  DPS_REQUIRE(dps_aligned16(src) && dps_aligned16(dst), DPS_ERR_ALIGN, "%s: tensors must be 16-byte aligned", who);
  DPS_REQUIRE(src != dst, DPS_ERR_INVALID, "%s: src and dst must not alias", who);
  const int64_t e4 = elems / 4;
  const int per = kCopyThreads * kCopyVec;
  dim3 grid((unsigned)((e4 + per - 1) / per), (unsigned)n_dst);
  gather_kernel<<<grid, kCopyThreads, 0, (cudaStream_t)stream>>>(src, idx, broadcast, dst, e4);
  DPS_LAUNCH_CHECK(who);
  return DPS_OK;
}

}  // namespace

extern "C" {

int dps_particle_norms(const float* partials, int P, int n, float* l2, float* l1, dps_stream_t stream) {
  DPS_REQUIRE(partials && P > 0 && n > 0 && (l2 || l1), DPS_ERR_INVALID, "dps_particle_norms: bad arguments");
  const int threads = 128, warps_per_block = threads / 32;
  particle_norms_kernel<<<(n + warps_per_block - 1) / warps_per_block, threads, 0, (cudaStream_t)stream>>>(
      partials, P, n, 0, 0.f, l2, l1, nullptr);
  DPS_LAUNCH_CHECK("dps_particle_norms");
  return DPS_OK;
}

int dps_guidance_coef(const float* partials, int P, int n, int mode, float scale, float* l2, float* coef,
                      dps_stream_t stream) {
  DPS_REQUIRE(partials && P > 0 && n > 0 && coef, DPS_ERR_INVALID, "dps_guidance_coef: bad arguments");
  DPS_REQUIRE(mode == DPS_COEF_NORM || mode == DPS_COEF_NORM_SQ || mode == DPS_COEF_GLOBAL_NORM, DPS_ERR_INVALID,
              "dps_guidance_coef: bad mode %d", mode);
  if (mode == DPS_COEF_GLOBAL_NORM) {
    global_norm_coef_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(partials, P, n, scale, l2, coef);
    DPS_LAUNCH_CHECK("dps_guidance_coef");
    return DPS_OK;
  }
  const int threads = 128, warps_per_block = threads / 32;
  particle_norms_kernel<<<(n + warps_per_block - 1) / warps_per_block, threads, 0, (cudaStream_t)stream>>>(
      partials, P, n, mode, scale, l2, nullptr, coef);
  DPS_LAUNCH_CHECK("dps_guidance_coef");
  return DPS_OK;
}

int dps_particle_logweights(const float* meas, const float* sem, int n, float tau, float meas_scale, int meas_pow,
                            float sem_scale, int sem_pow, float* logw, dps_stream_t stream) {
  DPS_REQUIRE(meas && logw && n > 0, DPS_ERR_INVALID, "dps_particle_logweights: bad arguments");
  DPS_REQUIRE((meas_pow == 1 || meas_pow == 2) && (sem_pow == 1 || sem_pow == 2), DPS_ERR_INVALID,
              "dps_particle_logweights: powers must be 1 or 2");
  logweights_kernel<<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(meas, sem, n, tau, meas_scale, meas_pow,
                                                                       sem_scale, sem_pow, logw);
  DPS_LAUNCH_CHECK("dps_particle_logweights");
  return DPS_OK;
}

int dps_weights_cdf(const float* logw, int n, int linear_mode, float* weights_out, float* cdf, float* lse_out,
                    int32_t* degenerate_out, dps_stream_t stream) {
  DPS_REQUIRE(logw && weights_out && cdf && degenerate_out && n > 0, DPS_ERR_INVALID, "dps_weights_cdf: bad arguments");
  weights_cdf_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(logw, n, linear_mode, weights_out, cdf, lse_out,
                                                           degenerate_out);
  DPS_LAUNCH_CHECK("dps_weights_cdf");
  return DPS_OK;
}

int dps_ancestors_multinomial(const float* cdf, int n, const double* uniforms, int n_draws, const int32_t* degenerate,
                              int64_t* ancestors, dps_stream_t stream) {
  DPS_REQUIRE(cdf && uniforms && ancestors && n > 0 && n_draws > 0, DPS_ERR_INVALID,
              "dps_ancestors_multinomial: bad arguments");
  ancestors_kernel<<<(n_draws + 255) / 256, 256, 0, (cudaStream_t)stream>>>(cdf, n, uniforms, n_draws, 0, degenerate,
                                                                            ancestors);
  DPS_LAUNCH_CHECK("dps_ancestors_multinomial");
  return DPS_OK;
}

int dps_ancestors_systematic(const float* cdf, int n, const double* u0, int n_draws, const int32_t* degenerate,
                             int64_t* ancestors, dps_stream_t stream) {
  DPS_REQUIRE(cdf && u0 && ancestors && n > 0 && n_draws > 0, DPS_ERR_INVALID,
              "dps_ancestors_systematic: bad arguments");
  ancestors_kernel<<<(n_draws + 255) / 256, 256, 0, (cudaStream_t)stream>>>(cdf, n, u0, n_draws, 1, degenerate,
                                                                            ancestors);
  DPS_LAUNCH_CHECK("dps_ancestors_systematic");
  return DPS_OK;
}

int dps_gather_particles(const float* src, const int64_t* ancestors, float* dst, int n_dst, int64_t elems,
                         dps_stream_t stream) {
  return launch_gather(src, ancestors, 0, dst, n_dst, elems, stream, "dps_gather_particles");
}

int dps_gather_particles_p2p(const float* const* peer_bases_dev, int n_per_rank, const int64_t* ancestors, float* dst,
                             int n_dst, int64_t elems, dps_stream_t stream) {
  DPS_REQUIRE(peer_bases_dev && ancestors && dst && n_per_rank > 0 && n_dst > 0 && n_dst <= 65535 && elems > 0,
              DPS_ERR_INVALID, "dps_gather_particles_p2p: bad arguments");
  DPS_REQUIRE(elems % 4 == 0, DPS_ERR_UNSUPPORTED, "dps_gather_particles_p2p: elems must be a multiple of 4");
  DPS_REQUIRE(dps_aligned16(dst), DPS_ERR_ALIGN, "dps_gather_particles_p2p: dst must be 16-byte aligned");
  const int64_t e4 = elems / 4;
  const int per = kCopyThreads * kCopyVec;
  dim3 grid((unsigned)((e4 + per - 1) / per), (unsigned)n_dst);
  gather_p2p_kernel<<<grid, kCopyThreads, 0, (cudaStream_t)stream>>>(peer_bases_dev, n_per_rank, ancestors, dst, e4);
  DPS_LAUNCH_CHECK("dps_gather_particles_p2p");
  return DPS_OK;
}

int dps_exchange_particles_p2p(const float* const* peer_bases_dev, uint32_t* const* signal_pads_dev, int rank, int world,
                               uint32_t epoch, int64_t slot_elems, int n_per_rank, const int64_t* ancestors, float* dst,
                               int n_dst, int64_t elems, dps_stream_t stream) {
  DPS_REQUIRE(peer_bases_dev && signal_pads_dev && ancestors && dst && n_per_rank > 0 && n_dst > 0 && n_dst <= 65535 &&
                  elems > 0,
              DPS_ERR_INVALID, "dps_exchange_particles_p2p: bad arguments");
  DPS_REQUIRE(world > 0 && world <= kCopyThreads && rank >= 0 && rank < world, DPS_ERR_INVALID,
              "dps_exchange_particles_p2p: bad rank %d of %d", rank, world);
  DPS_REQUIRE(elems % 4 == 0 && slot_elems % 4 == 0 && slot_elems >= 0, DPS_ERR_UNSUPPORTED,
              "dps_exchange_particles_p2p: elems and slot_elems must be multiples of 4");
  DPS_REQUIRE(dps_aligned16(dst), DPS_ERR_ALIGN, "dps_exchange_particles_p2p: dst must be 16-byte aligned");
  const int64_t e4 = elems / 4;
  const int per = kCopyThreads * kCopyVec;
  dim3 grid((unsigned)((e4 + per - 1) / per), (unsigned)n_dst);
  gather_p2p_sync_kernel<<<grid, kCopyThreads, 0, (cudaStream_t)stream>>>(peer_bases_dev, signal_pads_dev, rank, world, epoch,
                                                                          slot_elems, n_per_rank, ancestors, dst, e4);
  DPS_LAUNCH_CHECK("dps_exchange_particles_p2p");
  return DPS_OK;
}

int dps_argmin(const float* costs, int n, int64_t* best, float* best_cost, dps_stream_t stream) {
  DPS_REQUIRE(costs && best && n > 0, DPS_ERR_INVALID, "dps_argmin: bad arguments");
  argmin_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(costs, n, best, best_cost);
  DPS_LAUNCH_CHECK("dps_argmin");
  return DPS_OK;
}

int dps_broadcast_particle(const float* src, const int64_t* index, float* dst, int n_dst, int64_t elems,
                           dps_stream_t stream) {
  return launch_gather(src, index, 1, dst, n_dst, elems, stream, "dps_broadcast_particle");
}

}  // extern "C"
