// Operator plan (immutable device tables) shared by the operator translation units.
#pragma once
#include "common.cuh"

struct SepTables;     // blur_separable.cu
struct SparseTables;  // blur_sparse.cu
struct ResizeTables;  // resize.cu
struct PhaseTables;   // phase.cu
struct ResizeFused;   // resize_fused.cu
struct SepFused;      // blur_fused.cu

struct dps_operator {
  int kind = 0;
  int C = 0, H = 0, W = 0;
  int oC = 0, oH = 0, oW = 0;
  int P = 0;               // partial sums per particle written by forward
  int64_t aux_floats = 0;  // per particle
  int taps = 0;
  int device = 0;
  float* mask_dev = nullptr;  // inpainting: (H*W)
  SepTables* sep = nullptr;
  SparseTables* sparse = nullptr;
  ResizeTables* resize = nullptr;
  PhaseTables* phase = nullptr;
  SepFused* sepfused = nullptr;   // fused residual + cotangent kernel of the separable blur (blur_fused.cu)
  ResizeFused* rfused = nullptr;  // fused residual + cotangent kernel (resize_fused.cu), null when the shape is not covered
  int guidance_P = 0;             // partial sums per particle written by the fused guidance kernel (0: none)
};

// Arguments common to every forward / adjoint launch (already validated by operator.cu).
struct FwdArgs {
  dps_source src;
  const float* y;  // nullable
  int64_t y_stride;
  float* out;
  float* partials;  // nullable, (N, P, 2)
  float* aux;       // nullable
  int n;
};
struct AdjArgs {
  const float* r;
  const float* coef;  // nullable
  dps_source mask_src;
  int has_mask;
  const float* extra;  // nullable
  int64_t extra_stride;
  float* g;
  int64_t g_stride;
  const float* aux;
  int n;
};

// x̂₀ of one element given the source descriptor and per-particle base pointers
DPS_DEV float src_load(const float* x, const float* eps, int64_t i, float c1, float c2, int clip) {
  const float xv = ldg_stream(x + i);
  if (!eps) return xv;
  return x0_of(xv, ldg_stream(eps + i), c1, c2, clip);
}
DPS_DEV float4 src_load4(const float* x, const float* eps, int64_t i, float c1, float c2, int clip) {
  float4 xv = ldg_stream4(x + i);
  if (!eps) return xv;
  const float4 ev = ldg_stream4(eps + i);
  xv.x = x0_of(xv.x, ev.x, c1, c2, clip);
  xv.y = x0_of(xv.y, ev.y, c1, c2, clip);
  xv.z = x0_of(xv.z, ev.z, c1, c2, clip);
  xv.w = x0_of(xv.w, ev.w, c1, c2, clip);
  return xv;
}
// clamp-backward mask of one element (1 when no mask source / no eps / no clipping)
DPS_DEV float mask_load(const dps_source& s, int has_mask, int n, int64_t i) {
  if (!has_mask || !s.eps || !s.clip) return 1.0f;
  const float xv = ldg_stream(s.x + n * s.x_stride + i);
  const float ev = ldg_stream(s.eps + n * s.eps_stride + i);
  return clamp_pass(x0_pre(xv, ev, s.c1, s.c2));
}
DPS_DEV float4 mask_load4(const dps_source& s, int has_mask, int n, int64_t i) {
  if (!has_mask || !s.eps || !s.clip) return make_float4(1.f, 1.f, 1.f, 1.f);
  const float4 xv = ldg_stream4(s.x + n * s.x_stride + i);
  const float4 ev = ldg_stream4(s.eps + n * s.eps_stride + i);
  return make_float4(clamp_pass(x0_pre(xv.x, ev.x, s.c1, s.c2)), clamp_pass(x0_pre(xv.y, ev.y, s.c1, s.c2)),
                     clamp_pass(x0_pre(xv.z, ev.z, s.c1, s.c2)), clamp_pass(x0_pre(xv.w, ev.w, s.c1, s.c2)));
}

// 64-bit streaming load of a column pair
DPS_DEV float2 ldg_stream2(const float2* p) {
  float2 r;
  asm("ld.global.nc.L1::no_allocate.v2.f32 {%0,%1}, [%2];" : "=f"(r.x), "=f"(r.y) : "l"(p));
  return r;
}
// x̂₀ of a column pair, bit-identical to the scalar path (mul, mul, sub, clamp).  ptxas contracts mul.rn.f32x2 +
// add.rn.f32x2 into ONE FFMA2 (seen in SASS, CUDA 12.9), which would round c1·x only once and break the bit-identity
// with x0_pre / the reference.  The subtraction is therefore written as fma(a, 1, −b) with the 1 read from constant
// memory, a value ptxas cannot see: FMUL2, FMUL2, FFMA2 — three packed instructions, every product rounded on its own.
static __constant__ float dps_opaque_one = 1.0f;
DPS_DEV float2 x0_pair_pre(float2 x, float2 e, float c1, float c2) {
  const float2 a = __fmul2_rn(make_float2(c1, c1), x);
  const float2 b = __fmul2_rn(make_float2(c2, c2), e);
  const float one = dps_opaque_one;
  return __ffma2_rn(a, make_float2(one, one), make_float2(-b.x, -b.y));
}
DPS_DEV float2 x0_pair(float2 x, float2 e, float c1, float c2, int clip) {
  float2 v = x0_pair_pre(x, e, c1, c2);
  if (clip) { v.x = clamp1(v.x); v.y = clamp1(v.y); }
  return v;
}
// the same with the clamp bounds as data (±1 or ±inf): no branch in an unrolled row loop
DPS_DEV float2 x0_pair_bounds(float2 x, float2 e, float c1, float c2, float lo, float hi) {
  float2 v = x0_pair_pre(x, e, c1, c2);
  v.x = fminf(fmaxf(v.x, lo), hi);
  v.y = fminf(fmaxf(v.y, lo), hi);
  return v;
}

// per-operator launchers (each returns DPS_OK / error)
int inpaint_forward(const dps_operator* op, const FwdArgs& a, cudaStream_t st);
int inpaint_adjoint(const dps_operator* op, const AdjArgs& a, cudaStream_t st);
int inpaint_guidance(const dps_operator* op, const FwdArgs& a, float* g, int64_t g_stride, cudaStream_t st);

int sep_create(dps_operator* op, const float* taps1d_v, const float* taps1d_h, int rv, int rh);
void sep_destroy(dps_operator* op);
int sep_forward(const dps_operator* op, const FwdArgs& a, cudaStream_t st);
int sep_adjoint(const dps_operator* op, const AdjArgs& a, cudaStream_t st);

int sparse_create(dps_operator* op, const float* kernel, int ksize);
void sparse_destroy(dps_operator* op);
int sparse_forward(const dps_operator* op, const FwdArgs& a, cudaStream_t st);
int sparse_adjoint(const dps_operator* op, const AdjArgs& a, cudaStream_t st);

int resize_create(dps_operator* op, const int32_t* fov_h, const float* w_h, int taps_h, int out_h,
                  const int32_t* fov_w, const float* w_w, int taps_w, int out_w);
void resize_destroy(dps_operator* op);
int resize_forward(const dps_operator* op, const FwdArgs& a, cudaStream_t st);
int resize_adjoint(const dps_operator* op, const AdjArgs& a, cudaStream_t st);

int sep_fused_create(dps_operator* op, const float* taps_v, int rv, const float* taps_h, int rh);
void sep_fused_destroy(dps_operator* op);
int sep_fused_guidance(const dps_operator* op, const dps_source& src, const float* y, int64_t y_stride, float* r_out, float* g,
                       int64_t g_stride, float* partials, int n, cudaStream_t st);
void resize_fused_destroy(dps_operator* op);
int resize_fused_guidance(const dps_operator* op, const dps_source& src, const float* y, int64_t y_stride, float* r_out, float* g,
                          int64_t g_stride, float* partials, int n, cudaStream_t st);

int phase_create(dps_operator* op, int pad);
void phase_destroy(dps_operator* op);
int phase_forward(const dps_operator* op, const FwdArgs& a, cudaStream_t st);
int phase_adjoint(const dps_operator* op, const AdjArgs& a, cudaStream_t st);
int phase_guidance(const dps_operator* op, const dps_source& src, const float* y, int64_t y_stride, float* r_out, float* g,
                   int64_t g_stride, float* partials, float* aux, int n, cudaStream_t st);
