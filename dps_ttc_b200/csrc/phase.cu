// Phase retrieval operator (measurements.py:179-189, util/img_utils.py:26-30, util/fastmri_utils.py:67-89):
//   A(x) = | fftshift( FFT2_ortho( ifftshift( zero-pad(x) ) ) ) |  =  fftshift( |FFT2(pad(x))| ) / L
// (the input ifftshift only multiplies the spectrum by (−1)^{k1+k2}), and its
// Jacobian-transpose
//   Jᵀg [p] = (1/L)·Re Σ_k G[k]·conj(F[k])/|F[k]|·e^{−2πi k·p/L},   G = ifftshift(g),  zero where |F| = 0.
//
// L = H + 2·pad = 384 = 8·8·6 (256² images; also 256 = 8·8·4 for 128² and 192 = 8·8·3 for 64²: phase_impl.cuh is compiled once
// per length): hand-written mixed-radix Stockham FFT, radices 8, 8, L/64.  What makes it cheap:
//   * real input  → two image rows ride one complex FFT; only the half spectrum k2 ∈ [0,192] is kept and
//     the other half of the magnitude is written by Hermitian symmetry |F[−k]| = |F[k]|;
//   * zero padding → only the 256 non-zero rows are row-transformed, the column pass reads 256 of 384;
//   * Re(·) in the VJP → G is symmetrised (G[k]+G[−k])/2 so the back-transform is Hermitian too and again
//     runs on the half spectrum with two rows per complex FFT, cropped to the 256×256 image.
// forward : K1 rows  (x, ε) → Rt[k2][row]        (scratch, L2-resident at small N)
//           K2 cols  Rt → |F|/L → out = y − A(x̂₀) (or A(x̂₀)), Σr², Σ|r|, unit phase conj(F)/|F| → aux
// adjoint : A1 cols  (r, phase) → T[row][k2]      (scratch)
//           A2 rows  T → g = clamp-mask ⊙ (coef/L·Re(·) + extra)
// guidance (dps_operator_guidance): K1 (+ clamp-mask bytes) → fused columns (both transforms, residual and cotangent on chip) → A2.
// Two generations of kernels live here.  The default since the end of round 2 keeps the butterflies in REGISTERS (a thread owns a
// butterfly, shared memory is only the exchange between stages: phase_colsreg.cuh, phase_rowsreg.cuh — 512-thread CTAs, 40
// registers, 3 CTAs per SM); the earlier kernels run every Stockham stage through shared memory (fft_batch / fft_inplace in
// phase_impl.cuh) and remain selectable (DPSTTC_PHASE_{COLS,ROWS,FWD,ADJ}_REG=0) as the baseline the new ones are gated against.
// Roofline: ~16 MFLOP per particle against ≥ 13 MB of traffic, yet instruction- and barrier-bound: 20–26 % of the
// HBM peak on the algorithmic bytes; the scratch round trips (1.5T each way) stay in L2 at the particle counts the configs use.
#include <math.h>

#include <vector>

#include "operator.cuh"


struct PhaseTables {
  // [0, L) exp(−2πi j/L) (the shared-memory kernels stage its first half) · [L, L+64) W64^{k·r} at [8r+k] · [L+64, 2L+64) the full
  // table rebuilt from the half table (tw[j+L/2] = −tw[j]) · the last two again as (w.x, w.y, −w.y, w.x) quadruples (PHASE_PACKED=2)
  float2* tw = nullptr;
};


namespace {
constexpr int kThreads = 256;

// for (i = tid; i < kItems; i += kThreads) store(i, load(i)) with the loads of kBatch iterations issued before the
// first store: a "load, then store to shared" loop otherwise serialises one global round trip per iteration.
template <int kItems, int kBatch, class Load, class Store>
DPS_DEV void batched_copy(int tid, Load load, Store store) {
#pragma unroll 1
  for (int i0 = tid; i0 < kItems; i0 += kBatch * kThreads) {
    decltype(load(0)) v[kBatch];
#pragma unroll
    for (int b = 0; b < kBatch; ++b) {
      const int i = i0 + b * kThreads;
      if (i < kItems) v[b] = load(i);
    }
#pragma unroll
    for (int b = 0; b < kBatch; ++b) {
      const int i = i0 + b * kThreads;
      if (i < kItems) store(i, v[b]);
    }
  }
}

#include "phase_math.cuh"

DPS_DEV void stg_u8(unsigned char* p, unsigned v) { asm volatile("st.global.u8 [%0], %1;" ::"l"(p), "r"(v)); }
// read-only loads the compiler must not move across a barrier (issued early on purpose: in flight while a stage runs)
DPS_DEV unsigned char ldg_u8_pinned(const unsigned char* p) {
  unsigned v;
  // ld.volatile: ptxas moves a non-coherent load below the barrier it is meant to be in flight across (seen in SASS)
  asm volatile("ld.volatile.global.u8 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return (unsigned char)v;
}
DPS_DEV float ldg_ro_pinned(const float* p) {
  float v;
  asm volatile("ld.global.nc.f32 %0, [%1];" : "=f"(v) : "l"(p));
  return v;
}

// clamp-mask bytes of the fused guidance path (1 = gradient passes): written by the row kernel, read by the last kernel,
// so that x and ε are read ONCE per step (T/4 bytes each way instead of 2T)
DPS_DEV unsigned pack_pass4(const float* x, const float* eps, int64_t i, float c1, float c2, int clip) {
  if (!eps || !clip) return 0x01010101u;
  const float4 xv = ldg_stream4(x + i), ev = ldg_stream4(eps + i);
  return (clamp_pass(x0_pre(xv.x, ev.x, c1, c2)) != 0.f ? 1u : 0u) | (clamp_pass(x0_pre(xv.y, ev.y, c1, c2)) != 0.f ? 0x100u : 0u) |
         (clamp_pass(x0_pre(xv.z, ev.z, c1, c2)) != 0.f ? 0x10000u : 0u) | (clamp_pass(x0_pre(xv.w, ev.w, c1, c2)) != 0.f ? 0x1000000u : 0u);
}

// ---- K1: row transforms of the 256 image rows, two real rows per complex FFT ---------------------
int set_smem(const void* fn, size_t bytes) {
  DPS_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
  return DPS_OK;
}

}  // namespace

#define PHASE_R3 6
namespace ph384 {
#include "phase_impl.cuh"
}
#undef PHASE_R3
#define PHASE_R3 4
namespace ph256 {
#include "phase_impl.cuh"
}
#undef PHASE_R3
#define PHASE_R3 3
namespace ph192 {
#include "phase_impl.cuh"
}
#undef PHASE_R3

// The reference pads by int((oversample / 8)·256) = 64 whatever the image size (measurements.py:181-183): 256² → 384²
// (the shipped configs), 128² → 256², 64² → 192² (the size of the reference-made fixtures).
#define DPS_PHASE_DISPATCH(op, call)                                                              \
  switch ((op)->H) {                                                                              \
    case 256: return ph384::call;                                                                 \
    case 128: return ph256::call;                                                                 \
    case 64: return ph192::call;                                                                  \
  }                                                                                               \
  dps_set_error("phase retrieval: no kernels for %dx%d images", (op)->H, (op)->W);               \
  return DPS_ERR_UNSUPPORTED;

int phase_create(dps_operator* op, int pad) {
  DPS_REQUIRE(op->H == op->W && (op->H == 256 || op->H == 128 || op->H == 64) && pad == 64, DPS_ERR_UNSUPPORTED,
              "phase retrieval: kernels are built for 256x256, 128x128 and 64x64 images padded by 64 (got %dx%d, pad %d)", op->H,
              op->W, pad);
  DPS_PHASE_DISPATCH(op, create(op))
}

void phase_destroy(dps_operator* op) {
  if (!op->phase) return;
  cudaFree(op->phase->tw);
  delete op->phase;
  op->phase = nullptr;
}

int phase_forward(const dps_operator* op, const FwdArgs& a, cudaStream_t st) { DPS_PHASE_DISPATCH(op, forward(op, a, st)) }

int phase_adjoint(const dps_operator* op, const AdjArgs& a, cudaStream_t st) { DPS_PHASE_DISPATCH(op, adjoint(op, a, st)) }

int phase_guidance(const dps_operator* op, const dps_source& src, const float* y, int64_t y_stride, float* r_out, float* g,
                   int64_t g_stride, float* partials, float* aux, int n, cudaStream_t st) {
  DPS_PHASE_DISPATCH(op, guidance(op, src, y, y_stride, r_out, g, g_stride, partials, aux, n, st))
}
