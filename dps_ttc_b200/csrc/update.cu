// Fused posterior update: x̂₀ from ε, posterior mean, learned-range
// log-variance, σ·z noise and the ζ·∇ guidance step in ONE pass over the particle tensors.
//
// Roofline: pure streaming, HBM-bound.  Algorithmic bytes per particle (T = C·H·W·4 B):
//   DDPM: read x, ε, v, z, g, vjp; write x'  = 7T      DDIM: read x, ε, z?, g, vjp; write x' = 5-6T
// Every tensor is read once with 128-bit L1-bypassing loads; arithmetic order follows the
// reference line by line (mul, mul, add — no FMA contraction) so x̂₀, μ and the DDIM sample are
// bit-identical to the ATen path and only exp() may differ by an ulp.
#include "common.cuh"

namespace {

#ifndef UPD_THREADS
#define UPD_THREADS 256
#endif
#ifndef UPD_VEC
#define UPD_VEC 1
#endif
constexpr int kThreads = UPD_THREADS;
// float4 per thread and stream.  Measured (tools/build_variant.sh + kernel_bench, DDPM update): one vector per thread —
// 6-7 independent 16 B loads in flight per thread, twice the CTAs — beats two at every particle count: 8.98 → 8.32 µs at
// N = 8 (75 → 81 % of the HBM peak), 15.0 → 14.4 at 16, 27.2 → 26.7 at 32, 104.1 → 103.3 at 128; three or four lose.
constexpr int kVecPerThread = UPD_VEC;

struct UpdateArgs {
  const float* x;
  const float* eps;
  const float* v;
  const float* z;
  const float* g;
  const float* vjp;
  float* x_next;
  float* sample_out;
  float* x0_out;
  int64_t x_stride, eps_stride, v_stride, g_stride;
  int64_t chw4;  // float4 per particle
  float c1, c2;
  int clip;
  dps_step_consts k;
  dps_update_ext ext;  // used by the kExt instantiations only
};

// ---- Philox4x32-10 (Salmon et al., "Parallel random numbers: as easy as 1, 2, 3"; the Random123 constants) + Box–Muller ----
DPS_DEV uint4 philox4x32_10(uint4 c, uint2 key) {
  constexpr unsigned M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const unsigned hi0 = __umulhi(M0, c.x), lo0 = M0 * c.x;
    const unsigned hi1 = __umulhi(M1, c.z), lo1 = M1 * c.z;
    c = make_uint4(hi1 ^ c.y ^ key.x, lo1, hi0 ^ c.w ^ key.y, lo0);
    key.x += W0;
    key.y += W1;
  }
  return c;
}
// 24 random bits → u = (k + ½)·2⁻²⁴ ∈ [2⁻²⁵, 1 − 2⁻²⁵]: exact in fp32, never 0 or 1 (ln u finite, −2 ln u > 0)
DPS_DEV float u01(unsigned r) { return ((float)(r >> 8) + 0.5f) * 5.9604644775390625e-08f; }
// Box–Muller with the SFU forms (lg2 / rsq / sin / cos): the generator must not cost more than the HBM read of a z tensor
// it replaces.  Angle 2π(u − ½) ∈ (−π, π), the range where __sincosf is accurate to 2⁻²¹.
DPS_DEV float4 philox_normal4(int64_t i4, int64_t particle, const dps_update_ext& e) {
  const uint4 r = philox4x32_10(make_uint4((unsigned)i4, (unsigned)((uint64_t)i4 >> 32), (unsigned)particle, (unsigned)e.philox_step),
                                make_uint2((unsigned)e.philox_seed, (unsigned)(e.philox_seed >> 32)));
  float4 z;
  float s, c;
  const float t0 = -2.0f * __logf(u01(r.x));
  const float m0 = t0 * rsqrtf(t0);
  __sincosf(6.283185307179586f * (u01(r.y) - 0.5f), &s, &c);
  z.x = m0 * c;
  z.y = m0 * s;
  const float t1 = -2.0f * __logf(u01(r.z));
  const float m1 = t1 * rsqrtf(t1);
  __sincosf(6.283185307179586f * (u01(r.w) - 0.5f), &s, &c);
  z.z = m1 * c;
  z.w = m1 * s;
  return z;
}

DPS_DEV float ddpm_sample(float x, float e, float x0, float v, float z, const dps_step_consts& k) {
  // μ = p1·x̂₀ + p2·x                                       posterior_mean_variance.py:110-118
  // μ = model output (previous_x processor)                 :62-65
  float mean = k.mean_mode ? e : __fadd_rn(__fmul_rn(k.p1, x0), __fmul_rn(k.p2, x));
  if (!k.noise_on) return mean;  // gaussian_diffusion.py:473
  float logvar;
  if (k.var_mode == 0) {
    // frac = (v+1)/2 ; logσ² = frac·max_log + (1−frac)·min_log             :239-240
    float frac = __fmul_rn(__fadd_rn(v, 1.0f), 0.5f);
    logvar = __fadd_rn(__fmul_rn(frac, k.max_log), __fmul_rn(__fsub_rn(1.0f, frac), k.min_log));
  } else if (k.var_mode == 1) {
    logvar = k.max_log;
  } else {
    logvar = v;
  }
  // sample = μ + exp(½·logσ²)·z                                           gaussian_diffusion.py:474
  return __fadd_rn(mean, __fmul_rn(expf(__fmul_rn(0.5f, logvar)), z));
}

DPS_DEV float ddim_sample(float x, float x0, float z, float c1, float c2, const dps_step_consts& k) {
  // ε' = (c1·x − x̂₀)/c2                                                   gaussian_diffusion.py:506-509
  float eps2 = __fdiv_rn(__fsub_rn(__fmul_rn(c1, x), x0), c2);
  // x̂₀·sqrt(ᾱ_prev) + sqrt(1−ᾱ_prev−σ²)·ε'                                :495-498
  float s = __fadd_rn(__fmul_rn(x0, k.ddim_sa), __fmul_rn(k.ddim_sb, eps2));
  if (k.noise_on && k.ddim_sigma != 0.0f) s = __fadd_rn(s, __fmul_rn(k.ddim_sigma, z));
  return s;
}

#define DPS_FOR4(body) \
  {                    \
    body(x) body(y) body(z) body(w) \
  }

template <bool kDdim, bool kExt>
__global__ void __launch_bounds__(kThreads) posterior_update_kernel(const UpdateArgs a) {
  const int n = blockIdx.y;
  __shared__ float s_coef;
  const int64_t base4 = (int64_t)blockIdx.x * (kThreads * kVecPerThread) + threadIdx.x;
  const float* x = a.x + n * a.x_stride;
  const float* eps = a.eps + n * a.eps_stride;
  const float* v = a.v ? a.v + n * a.v_stride : nullptr;
  const float* z = a.z ? a.z + n * a.chw4 * 4 : nullptr;
  const float* g = a.g ? a.g + n * a.g_stride : nullptr;
  const float* vjp = a.vjp ? a.vjp + n * a.chw4 * 4 : nullptr;
  const int64_t obase = n * a.chw4 * 4;

  float4 vx[kVecPerThread], ve[kVecPerThread], vv[kVecPerThread], vz[kVecPerThread],
      vg[kVecPerThread], vj[kVecPerThread];
  const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
  // issue every load before the first use
#pragma unroll
  for (int u = 0; u < kVecPerThread; ++u) {
    const int64_t i4 = base4 + (int64_t)u * kThreads;
    const bool ok = i4 < a.chw4;
    vx[u] = ok ? ldg_stream4(x + i4 * 4) : zero4;
    ve[u] = ok ? ldg_stream4(eps + i4 * 4) : zero4;
    vv[u] = (ok && v) ? ldg_stream4(v + i4 * 4) : zero4;
    vz[u] = (ok && z) ? ldg_stream4(z + i4 * 4) : zero4;
    vg[u] = (ok && g) ? ldg_stream4(g + i4 * 4) : zero4;
    vj[u] = (ok && vjp) ? ldg_stream4(vjp + i4 * 4) : zero4;
  }
  float coef = 1.0f;
  if constexpr (kExt) {
    if (a.ext.partials) {
      // while the loads are in flight: ‖r_n‖ from the residual kernel's partial sums — the reduction order of
      // particle_norms_kernel (lane-strided fp64 sums, xor-shuffle tree), so the value is bit-identical to it.  Warp 0
      // does it once per CTA; the barrier that publishes it comes after the noise generation below, so neither the
      // partial-sum round trip nor the reduction sits on the other warps' critical path.
      if (threadIdx.x < 32) {
        const float* p = a.ext.partials + (int64_t)n * a.ext.P * 2;
        double sq = 0.0;
        for (int i = threadIdx.x; i < a.ext.P; i += 32) sq += (double)__ldg(p + 2 * i);
        sq = warp_sum(sq);
        if (threadIdx.x == 0) {
          const float nrm = (float)sqrt(sq);
          s_coef = a.ext.coef_mode == DPS_COEF_NORM ? (nrm > 0.f ? -a.ext.scale / nrm : 0.f) : -2.0f * a.ext.scale;
          if (a.ext.l2_out && blockIdx.x == 0) a.ext.l2_out[n] = nrm;
        }
      }
    }
    if (a.ext.use_philox && !z && a.k.noise_on) {
#pragma unroll
      for (int u = 0; u < kVecPerThread; ++u)
        vz[u] = philox_normal4(base4 + (int64_t)u * kThreads, a.ext.particle_offset + n, a.ext);
    }
    if (a.ext.partials) {
      __syncthreads();
      coef = s_coef;
    }
  }
#pragma unroll
  for (int u = 0; u < kVecPerThread; ++u) {
    const int64_t i4 = base4 + (int64_t)u * kThreads;
    if (i4 >= a.chw4) continue;
    float4 x0v, sv, nv;
#define DPS_ELEM(c)                                                                   \
  {                                                                                   \
    const float x0 = x0_of(vx[u].c, ve[u].c, a.c1, a.c2, a.clip);                     \
    const float s = kDdim ? ddim_sample(vx[u].c, x0, vz[u].c, a.c1, a.c2, a.k)        \
                          : ddpm_sample(vx[u].c, ve[u].c, x0, vv[u].c, vz[u].c, a.k);          \
    /* ∇ₓ = c1·g − c2·VJP_ε(g)  (chain rule through c1·x − c2·ε(x), App. A.4) */      \
    const float grad = __fsub_rn(__fmul_rn(a.c1, vg[u].c), __fmul_rn(a.c2, vj[u].c)); \
    x0v.c = x0;                                                                       \
    sv.c = s;                                                                         \
    nv.c = g ? __fsub_rn(s, (kExt && a.ext.partials) ? __fmul_rn(coef, grad) : grad) : s; \
  }
    DPS_FOR4(DPS_ELEM)
#undef DPS_ELEM
    stg_stream4(a.x_next + obase + i4 * 4, nv);
    if (a.sample_out) stg_stream4(a.sample_out + obase + i4 * 4, sv);
    if (a.x0_out) stg_stream4(a.x0_out + obase + i4 * 4, x0v);
  }
}

__global__ void __launch_bounds__(kThreads) x0_from_eps_kernel(const float* __restrict__ xb,
                                                               const float* __restrict__ eb,
                                                               float* __restrict__ x0b,
                                                               int64_t x_stride, int64_t eps_stride,
                                                               int64_t chw4, float c1, float c2,
                                                               int clip) {
  const int n = blockIdx.y;
  const float* x = xb + n * x_stride;
  const float* e = eb + n * eps_stride;
  float* o = x0b + n * chw4 * 4;
  const int64_t base4 = (int64_t)blockIdx.x * (kThreads * kVecPerThread) + threadIdx.x;
  float4 vx[kVecPerThread], ve[kVecPerThread];
#pragma unroll
  for (int u = 0; u < kVecPerThread; ++u) {
    const int64_t i4 = base4 + (int64_t)u * kThreads;
    if (i4 < chw4) {
      vx[u] = ldg_stream4(x + i4 * 4);
      ve[u] = ldg_stream4(e + i4 * 4);
    }
  }
#pragma unroll
  for (int u = 0; u < kVecPerThread; ++u) {
    const int64_t i4 = base4 + (int64_t)u * kThreads;
    if (i4 >= chw4) continue;
    float4 r;
    r.x = x0_of(vx[u].x, ve[u].x, c1, c2, clip);
    r.y = x0_of(vx[u].y, ve[u].y, c1, c2, clip);
    r.z = x0_of(vx[u].z, ve[u].z, c1, c2, clip);
    r.w = x0_of(vx[u].w, ve[u].w, c1, c2, clip);
    stg_stream4(o + i4 * 4, r);
  }
}

__global__ void __launch_bounds__(kThreads) q_sample_kernel(const float* __restrict__ y,
                                                            const float* __restrict__ noise, float a,
                                                            float b, float* __restrict__ out,
                                                            int64_t n) {
  // coef1·x_start + coef2·noise                                          gaussian_diffusion.py:151
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x)
    out[i] = __fadd_rn(__fmul_rn(a, y[i]), __fmul_rn(b, noise[i]));
}

int check_source(const dps_source* s, const char* who) {
  DPS_REQUIRE(s && s->x, DPS_ERR_INVALID, "%s: null source", who);
  DPS_REQUIRE(dps_aligned16(s->x) && (s->x_stride % 4 == 0), DPS_ERR_ALIGN,
              "%s: x must be 16-byte aligned with a stride that is a multiple of 4", who);
  if (s->eps)
    DPS_REQUIRE(dps_aligned16(s->eps) && (s->eps_stride % 4 == 0), DPS_ERR_ALIGN,
                "%s: eps must be 16-byte aligned with a stride that is a multiple of 4", who);
  return DPS_OK;
}

template <bool kDdim>
int launch_update(const dps_source* src, const float* v, int64_t v_stride, const float* z,
                  const float* g, int64_t g_stride, const float* vjp, const dps_step_consts* k,
                  float* x_next, float* sample_out, float* x0_out, int n, int64_t chw,
                  dps_stream_t stream, const char* who, const dps_update_ext* ext = nullptr) {
  if (int rc = check_source(src, who)) return rc;
  DPS_REQUIRE(src->eps, DPS_ERR_INVALID, "%s: eps is required", who);
  DPS_REQUIRE(k && x_next, DPS_ERR_INVALID, "%s: null consts/output", who);
  DPS_REQUIRE(n > 0 && n <= 65535 && chw > 0, DPS_ERR_INVALID, "%s: bad sizes n=%d chw=%lld", who, n,
              (long long)chw);
  DPS_REQUIRE(chw % 4 == 0, DPS_ERR_UNSUPPORTED, "%s: C*H*W must be a multiple of 4", who);
  DPS_REQUIRE(dps_aligned16(x_next) && dps_aligned16(v) && dps_aligned16(z) && dps_aligned16(g) &&
                  dps_aligned16(vjp) && dps_aligned16(sample_out) && dps_aligned16(x0_out) &&
                  v_stride % 4 == 0 && g_stride % 4 == 0,
              DPS_ERR_ALIGN, "%s: tensors must be 16-byte aligned", who);
  const bool device_noise = ext && ext->use_philox;
  if (ext && ext->partials) {
    DPS_REQUIRE(ext->P > 0 && (ext->coef_mode == DPS_COEF_NORM || ext->coef_mode == DPS_COEF_NORM_SQ), DPS_ERR_INVALID,
                "%s: deferred coefficient needs P > 0 and coef_mode NORM or NORM_SQ", who);
    DPS_REQUIRE(g, DPS_ERR_INVALID, "%s: deferred coefficient without a cotangent g", who);
  }
  if (!kDdim) {
    DPS_REQUIRE(k->var_mode >= 0 && k->var_mode <= 2, DPS_ERR_INVALID, "%s: bad var_mode", who);
    DPS_REQUIRE(v || k->var_mode == 1 || !k->noise_on, DPS_ERR_INVALID,
                "%s: variance channels required", who);
    DPS_REQUIRE(z || device_noise || !k->noise_on, DPS_ERR_INVALID, "%s: noise z required when noise_on", who);
  } else {
    DPS_REQUIRE(z || device_noise || !k->noise_on || k->ddim_sigma == 0.0f, DPS_ERR_INVALID,
                "%s: noise z required when sigma != 0", who);
  }
  UpdateArgs a;
  a.x = src->x;
  a.eps = src->eps;
  a.v = v;
  a.z = z;
  a.g = g;
  a.vjp = g ? vjp : nullptr;
  a.x_next = x_next;
  a.sample_out = sample_out;
  a.x0_out = x0_out;
  a.x_stride = src->x_stride;
  a.eps_stride = src->eps_stride;
  a.v_stride = v_stride;
  a.g_stride = g_stride;
  a.chw4 = chw / 4;
  a.c1 = src->c1;
  a.c2 = src->c2;
  a.clip = src->clip;
  a.k = *k;
  a.ext = ext ? *ext : dps_update_ext{};
  const int per_block = kThreads * kVecPerThread;
  dim3 grid((unsigned)((a.chw4 + per_block - 1) / per_block), (unsigned)n);
  if (ext)
    posterior_update_kernel<kDdim, true><<<grid, kThreads, 0, (cudaStream_t)stream>>>(a);
  else
    posterior_update_kernel<kDdim, false><<<grid, kThreads, 0, (cudaStream_t)stream>>>(a);
  DPS_LAUNCH_CHECK(who);
  return DPS_OK;
}

}  // namespace

namespace {
// grad = c1·g − c2·vjp (kApply = false)   |   x' = sample − grad (kApply = true); b may be particle-broadcast
template <bool kApply>
__global__ void __launch_bounds__(kThreads) grad_kernel(const float* __restrict__ a, int64_t a_stride,
                                                        const float* __restrict__ b, int64_t b_stride, float c1,
                                                        float c2, float* __restrict__ out, int64_t chw4) {
  const int n = blockIdx.y;
  const float* ap = a + n * a_stride;
  const float* bp = b ? b + n * b_stride : nullptr;
  float* op = out + n * chw4 * 4;
  const int64_t base4 = (int64_t)blockIdx.x * (kThreads * kVecPerThread) + threadIdx.x;
  float4 va[kVecPerThread], vb[kVecPerThread];
#pragma unroll
  for (int u = 0; u < kVecPerThread; ++u) {
    const int64_t i4 = base4 + (int64_t)u * kThreads;
    va[u] = vb[u] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (i4 < chw4) {
      va[u] = ldg_stream4(ap + i4 * 4);
      if (bp) vb[u] = b_stride ? ldg_stream4(bp + i4 * 4) : ldg_ro4(bp + i4 * 4);
    }
  }
#pragma unroll
  for (int u = 0; u < kVecPerThread; ++u) {
    const int64_t i4 = base4 + (int64_t)u * kThreads;
    if (i4 >= chw4) continue;
    float4 r;
    if (kApply) {
      r.x = __fsub_rn(va[u].x, vb[u].x); r.y = __fsub_rn(va[u].y, vb[u].y);
      r.z = __fsub_rn(va[u].z, vb[u].z); r.w = __fsub_rn(va[u].w, vb[u].w);
    } else {
      r.x = __fsub_rn(__fmul_rn(c1, va[u].x), __fmul_rn(c2, vb[u].x));
      r.y = __fsub_rn(__fmul_rn(c1, va[u].y), __fmul_rn(c2, vb[u].y));
      r.z = __fsub_rn(__fmul_rn(c1, va[u].z), __fmul_rn(c2, vb[u].z));
      r.w = __fsub_rn(__fmul_rn(c1, va[u].w), __fmul_rn(c2, vb[u].w));
    }
    stg_stream4(op + i4 * 4, r);
  }
}

template <bool kApply>
int launch_grad(const float* a, int64_t a_stride, const float* b, int64_t b_stride, float c1, float c2, float* out,
                int n, int64_t chw, dps_stream_t stream, const char* who) {
  DPS_REQUIRE(a && out && (b || !kApply), DPS_ERR_INVALID, "%s: null tensor", who);
  DPS_REQUIRE(n > 0 && n <= 65535 && chw > 0 && chw % 4 == 0, DPS_ERR_INVALID, "%s: bad sizes", who);
  DPS_REQUIRE(a_stride >= chw && a_stride % 4 == 0 && (b_stride == 0 || (b_stride >= chw && b_stride % 4 == 0)),
              DPS_ERR_INVALID, "%s: bad particle strides", who);
  DPS_REQUIRE(dps_aligned16(a) && dps_aligned16(out) && (!b || dps_aligned16(b)), DPS_ERR_ALIGN,
              "%s: tensors must be 16-byte aligned", who);
  const int64_t chw4 = chw / 4;
  const int per_block = kThreads * kVecPerThread;
  dim3 grid((unsigned)((chw4 + per_block - 1) / per_block), (unsigned)n);
  grad_kernel<kApply><<<grid, kThreads, 0, (cudaStream_t)stream>>>(a, a_stride, b, b_stride, c1, c2, out, chw4);
  DPS_LAUNCH_CHECK(who);
  return DPS_OK;
}
}  // namespace

// Per-particle Σ(a − ref)² and Σ|a − ref| as P partial sums per particle (finished by dps_particle_norms): the PSNR /
// distance bookkeeping of the drivers (compute_metrics.py:93-98).  CTA p of particle n takes the float4 items
// p, p + P, … of the particle in steps of 256·P — a fixed assignment and a fixed tree, so the sums are reproducible.
__global__ void __launch_bounds__(256) sqdiff_kernel(const float* __restrict__ a, int64_t a_stride, const float* __restrict__ ref,
                                                     int64_t ref_stride, int64_t chw, float* __restrict__ partials) {
  __shared__ float red[64];
  const int n = blockIdx.y, P = gridDim.x;
  const float* ap = a + n * a_stride;
  const float* rp = ref + n * ref_stride;
  float sq = 0.f, ab = 0.f;
  for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < chw / 4; i += (int64_t)256 * P) {
    const float4 u = ldg_stream4(ap + 4 * i);
    const float4 v = ldg_ro4(rp + 4 * i);
    const float dx = u.x - v.x, dy = u.y - v.y, dz = u.z - v.z, dw = u.w - v.w;
    sq += dx * dx + dy * dy + dz * dz + dw * dw;
    ab += fabsf(dx) + fabsf(dy) + fabsf(dz) + fabsf(dw);
  }
  block_sum2(sq, ab, red);
  if (threadIdx.x == 0) {
    float* pp = partials + ((int64_t)n * P + blockIdx.x) * 2;
    pp[0] = sq;
    pp[1] = ab;
  }
}

extern "C" {

int dps_x0_from_eps(const dps_source* src, float* x0, int n, int64_t chw, dps_stream_t stream) {
  if (int rc = check_source(src, "dps_x0_from_eps")) return rc;
  DPS_REQUIRE(src->eps && x0, DPS_ERR_INVALID, "dps_x0_from_eps: eps and x0 are required");
  DPS_REQUIRE(n > 0 && n <= 65535 && chw > 0, DPS_ERR_INVALID, "dps_x0_from_eps: bad sizes");
  DPS_REQUIRE(chw % 4 == 0, DPS_ERR_UNSUPPORTED, "dps_x0_from_eps: C*H*W must be a multiple of 4");
  DPS_REQUIRE(dps_aligned16(x0), DPS_ERR_ALIGN, "dps_x0_from_eps: x0 must be 16-byte aligned");
  const int64_t chw4 = chw / 4;
  const int per_block = kThreads * kVecPerThread;
  dim3 grid((unsigned)((chw4 + per_block - 1) / per_block), (unsigned)n);
  x0_from_eps_kernel<<<grid, kThreads, 0, (cudaStream_t)stream>>>(
      src->x, src->eps, x0, src->x_stride, src->eps_stride, chw4, src->c1, src->c2, src->clip);
  DPS_LAUNCH_CHECK("dps_x0_from_eps");
  return DPS_OK;
}


int dps_guidance_grad(const float* g, int64_t g_stride, const float* vjp, float c1, float c2, float* grad, int n,
                      int64_t chw, dps_stream_t stream) {
  return launch_grad<false>(g, g_stride, vjp, chw, c1, c2, grad, n, chw, stream, "dps_guidance_grad");
}

int dps_apply_gradient(const float* sample, const float* grad, int64_t grad_stride, float* x_next, int n, int64_t chw,
                       dps_stream_t stream) {
  return launch_grad<true>(sample, chw, grad, grad_stride, 0.f, 0.f, x_next, n, chw, stream, "dps_apply_gradient");
}

int dps_posterior_update_ddpm(const dps_source* src, const float* v, int64_t v_stride,
                              const float* z, const float* g, int64_t g_stride, const float* vjp,
                              const dps_step_consts* k, float* x_next, float* sample_out,
                              float* x0_out, int n, int64_t chw, dps_stream_t stream) {
  return launch_update<false>(src, v, v_stride, z, g, g_stride, vjp, k, x_next, sample_out, x0_out,
                              n, chw, stream, "dps_posterior_update_ddpm");
}

int dps_posterior_update_ddim(const dps_source* src, const float* z, const float* g,
                              int64_t g_stride, const float* vjp, const dps_step_consts* k,
                              float* x_next, float* sample_out, float* x0_out, int n, int64_t chw,
                              dps_stream_t stream) {
  return launch_update<true>(src, nullptr, 0, z, g, g_stride, vjp, k, x_next, sample_out, x0_out, n,
                             chw, stream, "dps_posterior_update_ddim");
}

int dps_posterior_update_ddpm_ext(const dps_source* src, const float* v, int64_t v_stride, const float* z, const float* g,
                                  int64_t g_stride, const float* vjp, const dps_step_consts* k, const dps_update_ext* ext,
                                  float* x_next, int n, int64_t chw, dps_stream_t stream) {
  DPS_REQUIRE(ext, DPS_ERR_INVALID, "dps_posterior_update_ddpm_ext: null ext");
  return launch_update<false>(src, v, v_stride, z, g, g_stride, vjp, k, x_next, nullptr, nullptr, n, chw, stream,
                              "dps_posterior_update_ddpm_ext", ext);
}

int dps_posterior_update_ddim_ext(const dps_source* src, const float* z, const float* g, int64_t g_stride, const float* vjp,
                                  const dps_step_consts* k, const dps_update_ext* ext, float* x_next, int n, int64_t chw,
                                  dps_stream_t stream) {
  DPS_REQUIRE(ext, DPS_ERR_INVALID, "dps_posterior_update_ddim_ext: null ext");
  return launch_update<true>(src, nullptr, 0, z, g, g_stride, vjp, k, x_next, nullptr, nullptr, n, chw, stream,
                             "dps_posterior_update_ddim_ext", ext);
}

int dps_particle_sqdiff(const float* a, int64_t a_stride, const float* ref, int64_t ref_stride, int n_particles,
                        int64_t chw, float* partials, int P, dps_stream_t stream) {
  DPS_REQUIRE(a && ref && partials && n_particles > 0 && chw > 0 && chw % 4 == 0 && P > 0 && P <= 65535, DPS_ERR_INVALID,
              "dps_particle_sqdiff: bad arguments (chw must be a multiple of 4)");
  DPS_REQUIRE(dps_aligned16(a) && dps_aligned16(ref) && a_stride % 4 == 0 && ref_stride % 4 == 0, DPS_ERR_INVALID,
              "dps_particle_sqdiff: pointers and strides must be 16-byte aligned");
  sqdiff_kernel<<<dim3((unsigned)P, (unsigned)n_particles), 256, 0, (cudaStream_t)stream>>>(a, a_stride, ref, ref_stride,
                                                                                             chw, partials);
  DPS_LAUNCH_CHECK("dps_particle_sqdiff");
  return DPS_OK;
}

int dps_q_sample(const float* y, const float* noise, float a, float b, float* out, int64_t n,
                 dps_stream_t stream) {
  DPS_REQUIRE(y && noise && out && n > 0, DPS_ERR_INVALID, "dps_q_sample: bad arguments");
  const int blocks = (int)((n + 255) / 256 < 148 * 8 ? (n + 255) / 256 : 148 * 8);
  q_sample_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(y, noise, a, b, out, n);
  DPS_LAUNCH_CHECK("dps_q_sample");
  return DPS_OK;
}

}  // extern "C"
