// Inpainting operator (measurements.py:151-168): A x = mask ⊙ x, Aᵀ = A.  Pure streaming.
// Bytes per particle: forward 2T (x, ε) + T (write r) [mask/y are shared and stay in L2];
// adjoint T (r) + 2T (x, ε for the clamp mask) + T (write g).  HBM-bound.
#include "operator.cuh"

namespace {
constexpr int kThreads = 256;
constexpr int kVec = 2;                      // forward: float4 per thread (also fixes the partial-sum split)
constexpr int kPerBlock4 = kThreads * kVec;  // float4 per CTA
// adjoint: ONE float4 per thread and stream (4 loads in flight per thread, twice the CTAs) — measured against 2:
// 63.1 vs 65.0 µs at N = 128, 16.6 vs 17.4 µs at N = 32, 5.70 vs 6.05 µs at N = 8 (profiles/r1j_variants.md).  The
// forward keeps 2 (54.1 vs 58.7 µs): its block reduction amortises over twice the data.
constexpr int kVecAdj = 1;
constexpr int kPerBlockAdj4 = kThreads * kVecAdj;

__global__ void __launch_bounds__(kThreads) inpaint_fwd_kernel(const FwdArgs a, const float* __restrict__ mask,
                                                               int64_t chw4, int hw4, int P) {
  __shared__ float red[64];
  const int n = blockIdx.y;
  const float* x = a.src.x + n * a.src.x_stride;
  const float* eps = a.src.eps ? a.src.eps + n * a.src.eps_stride : nullptr;
  const float* y = a.y ? a.y + n * a.y_stride : nullptr;
  float* out = a.out + n * chw4 * 4;
  float sq = 0.f, ab = 0.f;
#pragma unroll
  for (int u = 0; u < kVec; ++u) {
    const int64_t i4 = (int64_t)blockIdx.x * kPerBlock4 + u * kThreads + threadIdx.x;
    if (i4 >= chw4) continue;
    const float4 x0 = src_load4(x, eps, i4 * 4, a.src.c1, a.src.c2, a.src.clip);
    const float4 m = *reinterpret_cast<const float4*>(mask + (i4 % hw4) * 4);
    float4 o = make_float4(__fmul_rn(x0.x, m.x), __fmul_rn(x0.y, m.y), __fmul_rn(x0.z, m.z),
                           __fmul_rn(x0.w, m.w));
    if (y) {
      const float4 yv = ldg_ro4(y + i4 * 4);
      o = make_float4(__fsub_rn(yv.x, o.x), __fsub_rn(yv.y, o.y), __fsub_rn(yv.z, o.z),
                      __fsub_rn(yv.w, o.w));
    }
    stg_stream4(out + i4 * 4, o);
    sq += o.x * o.x + o.y * o.y + o.z * o.z + o.w * o.w;
    ab += fabsf(o.x) + fabsf(o.y) + fabsf(o.z) + fabsf(o.w);
  }
  if (a.partials) {
    block_sum2(sq, ab, red);
    if (threadIdx.x == 0) {
      float* p = a.partials + ((int64_t)n * P + blockIdx.x) * 2;
      p[0] = sq;
      p[1] = ab;
    }
  }
}

__global__ void __launch_bounds__(kThreads) inpaint_adj_kernel(const AdjArgs a, const float* __restrict__ mask,
                                                               int64_t chw4, int hw4) {
  const int n = blockIdx.y;
  const float* r = a.r + n * chw4 * 4;
  const float coef = a.coef ? a.coef[n] : 1.0f;
  const float* extra = a.extra ? a.extra + n * a.extra_stride : nullptr;
  float* g = a.g + n * a.g_stride;
#pragma unroll
  for (int u = 0; u < kVecAdj; ++u) {
    const int64_t i4 = (int64_t)blockIdx.x * kPerBlockAdj4 + u * kThreads + threadIdx.x;
    if (i4 >= chw4) continue;
    const float4 rv = ldg_stream4(r + i4 * 4);
    const float4 m = *reinterpret_cast<const float4*>(mask + (i4 % hw4) * 4);
    const float4 pass = mask_load4(a.mask_src, a.has_mask, n, i4 * 4);
    float4 o = make_float4(coef * (m.x * rv.x), coef * (m.y * rv.y), coef * (m.z * rv.z),
                           coef * (m.w * rv.w));
    if (extra) {
      const float4 e = ldg_stream4(extra + i4 * 4);
      o.x += e.x; o.y += e.y; o.z += e.z; o.w += e.w;
    }
    o.x *= pass.x; o.y *= pass.y; o.z *= pass.z; o.w *= pass.w;
    stg_stream4(g + i4 * 4, o);
  }
}
// Fused guidance (dps_operator_guidance): r = y − m ⊙ x̂₀ and the unscaled masked cotangent g = 1[|pre| ≤ 1] ⊙ m ⊙ r in ONE
// streaming pass — x, ε and y read once, g written once (3T + M instead of 5T + 2M); r only leaves the chip if asked for.
// Same float4-per-thread split and block reduction as the forward kernel, so the partial sums are bit-identical to it.
__global__ void __launch_bounds__(kThreads) inpaint_guidance_kernel(const FwdArgs a, float* __restrict__ gb, int64_t g_stride,
                                                                    const float* __restrict__ mask, int64_t chw4, int hw4, int P) {
  __shared__ float red[64];
  const int n = blockIdx.y;
  const float* x = a.src.x + n * a.src.x_stride;
  const float* eps = a.src.eps + n * a.src.eps_stride;
  const float* y = a.y + n * a.y_stride;
  float* out = a.out ? a.out + n * chw4 * 4 : nullptr;
  float* g = gb + n * g_stride;
  float sq = 0.f, ab = 0.f;
  float4 xv[kVec], ev[kVec], yv[kVec], m[kVec];
#pragma unroll
  for (int u = 0; u < kVec; ++u) {  // every load in flight before the first use
    const int64_t i4 = (int64_t)blockIdx.x * kPerBlock4 + u * kThreads + threadIdx.x;
    if (i4 >= chw4) continue;
    xv[u] = ldg_stream4(x + i4 * 4);
    ev[u] = ldg_stream4(eps + i4 * 4);
    yv[u] = ldg_ro4(y + i4 * 4);
    m[u] = *reinterpret_cast<const float4*>(mask + (i4 % hw4) * 4);
  }
#pragma unroll
  for (int u = 0; u < kVec; ++u) {
    const int64_t i4 = (int64_t)blockIdx.x * kPerBlock4 + u * kThreads + threadIdx.x;
    if (i4 >= chw4) continue;
    float4 o, gv;
#define DPS_INP(c)                                                          \
  {                                                                         \
    const float pre = x0_pre(xv[u].c, ev[u].c, a.src.c1, a.src.c2);         \
    const float x0 = a.src.clip ? clamp1(pre) : pre;                        \
    o.c = __fsub_rn(yv[u].c, __fmul_rn(x0, m[u].c));                        \
    gv.c = (!a.src.clip || clamp_pass(pre) != 0.f) ? m[u].c * o.c : 0.f;    \
  }
    DPS_INP(x) DPS_INP(y) DPS_INP(z) DPS_INP(w)
#undef DPS_INP
    if (out) stg_stream4(out + i4 * 4, o);
    stg_stream4(g + i4 * 4, gv);
    sq += o.x * o.x + o.y * o.y + o.z * o.z + o.w * o.w;
    ab += fabsf(o.x) + fabsf(o.y) + fabsf(o.z) + fabsf(o.w);
  }
  if (a.partials) {
    block_sum2(sq, ab, red);
    if (threadIdx.x == 0) {
      float* p = a.partials + ((int64_t)n * P + blockIdx.x) * 2;
      p[0] = sq;
      p[1] = ab;
    }
  }
}
}  // namespace

int inpaint_guidance(const dps_operator* op, const FwdArgs& a, float* g, int64_t g_stride, cudaStream_t st) {
  DPS_REQUIRE(a.src.eps && a.y, DPS_ERR_INVALID, "inpainting guidance needs eps and the measurement");
  const int64_t chw4 = (int64_t)op->C * op->H * op->W / 4;
  dim3 grid((unsigned)op->P, (unsigned)a.n);
  inpaint_guidance_kernel<<<grid, kThreads, 0, st>>>(a, g, g_stride, op->mask_dev, chw4, op->H * op->W / 4, op->P);
  DPS_LAUNCH_CHECK("inpaint_guidance");
  return DPS_OK;
}

int inpaint_partials(int C, int H, int W) {
  const int64_t chw4 = (int64_t)C * H * W / 4;
  return (int)((chw4 + kPerBlock4 - 1) / kPerBlock4);
}

int inpaint_forward(const dps_operator* op, const FwdArgs& a, cudaStream_t st) {
  const int64_t chw4 = (int64_t)op->C * op->H * op->W / 4;
  dim3 grid((unsigned)op->P, (unsigned)a.n);
  inpaint_fwd_kernel<<<grid, kThreads, 0, st>>>(a, op->mask_dev, chw4, op->H * op->W / 4, op->P);
  DPS_LAUNCH_CHECK("inpaint_forward");
  return DPS_OK;
}

int inpaint_adjoint(const dps_operator* op, const AdjArgs& a, cudaStream_t st) {
  const int64_t chw4 = (int64_t)op->C * op->H * op->W / 4;
  dim3 grid((unsigned)((chw4 + kPerBlockAdj4 - 1) / kPerBlockAdj4), (unsigned)a.n);
  inpaint_adj_kernel<<<grid, kThreads, 0, st>>>(a, op->mask_dev, chw4, op->H * op->W / 4);
  DPS_LAUNCH_CHECK("inpaint_adjoint");
  return DPS_OK;
}
