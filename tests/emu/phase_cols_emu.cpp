// CPU emulation of phase_cols_fused_reg (dps_ttc_b200/csrc/phase_colsreg.cuh) — TEST INFRASTRUCTURE ONLY.
// The kernel is written as barrier-free phases that take the thread index as an argument; this file gives the device
// intrinsics host shims, includes the very same headers and runs the phases thread by thread (a loop over tid stands for
// the CTA, the end of each loop for __syncthreads) for every column group of one plane.  It checks the index logic of the
// register-resident FFT (thread roles, Stockham maps, exchange layout, output positions) where no GPU is available;
// tests/test_phase_emu.py compares the result with a plain numpy DFT.  Built by the test with
//   g++ -O1 -shared -fPIC -DPHASE_R3={6,4,3} -I dps_ttc_b200/csrc
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

struct float2 { float x, y; };
struct float4 { float x, y, z, w; };
static inline float2 make_float2(float x, float y) { return float2{x, y}; }
static inline float4 make_float4(float x, float y, float z, float w) { return float4{x, y, z, w}; }
#define DPS_DEV static inline
// packed fp32 intrinsics of sm_100 (PHASE_PACKED builds): componentwise, every product / sum rounded as on the device
static inline float2 __fadd2_rn(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
static inline float2 __fmul2_rn(float2 a, float2 b) { return make_float2(a.x * b.x, a.y * b.y); }
static inline float2 __ffma2_rn(float2 a, float2 b, float2 c) { return make_float2(std::fmaf(a.x, b.x, c.x), std::fmaf(a.y, b.y, c.y)); }
static inline float emu_rsqrtf(float x) { return 1.0f / std::sqrt(x); }
#define rsqrtf emu_rsqrtf
static inline float __fsub_rn(float a, float b) { return a - b; }
static inline float2 ldg_stream2(const float2* p) { return *p; }
static inline float ldg_stream(const float* p) { return *p; }
static inline float ldg_ro(const float* p) { return *p; }
static inline float ldg_ro_pinned(const float* p) { return *p; }
static inline void stg_stream(float* p, float v) { *p = v; }
static inline void stg_stream2(float* p, const float2& v) { p[0] = v.x; p[1] = v.y; }

#include "phase_math.cuh"
#include "phase_dims.cuh"
#include "phase_colsreg.cuh"


// twiddle tables exactly as ph*::create builds them (float2 entries, or (w.x, w.y, −w.y, w.x) quadruples under PHASE_PACKED == 2)
static inline tw_t emu_tw_entry(float2 w) {
#if PHASE_PACKED == 2
  return make_float4(w.x, w.y, -w.y, w.x);
#else
  return w;
#endif
}
struct EmuTables {
  std::vector<tw_t> twf, w64;
  EmuTables() : twf(kL), w64(64) {
    std::vector<float2> tw(kL);
    for (int j = 0; j < kL; ++j) {
      const double a = -2.0 * M_PI * j / kL;
      tw[j] = make_float2((float)cos(a), (float)sin(a));
    }
    for (int r = 0; r < 8; ++r)
      for (int k = 0; k < 8; ++k) {
        const int j = kR3 * k * r;
        w64[8 * r + k] = emu_tw_entry(j >= kTW ? make_float2(-tw[j - kTW].x, -tw[j - kTW].y) : tw[j]);
      }
    for (int j = 0; j < kL; ++j) twf[j] = emu_tw_entry(j >= kTW ? make_float2(-tw[j - kTW].x, -tw[j - kTW].y) : tw[j]);
  }
};

extern "C" int emu_dims(int* out) {
  out[0] = kL; out[1] = kImg; out[2] = kHalf; out[3] = kColGroups;
  return 0;
}

// rt: Rt[k2][row] (kHalf × kImg complex), y: L×L, r_out: L×L or null, t: T[row][k2] (kImg × kHalf complex),
// partials: kColGroups × 2
extern "C" int emu_cols(const float* rt, const float* y, float* r_out, float* t, float* partials) {
  EmuTables T;
  std::vector<float2> A(kSeq * kLQ), B(kSeq * kLQ);
  std::vector<ColsRegs> R(kT2);
  std::vector<ColsY> Y(kT2);
  for (int grp = 0; grp < kColGroups; ++grp) {
    // poison the exchange buffers: a read of a word nobody wrote shows up as NaN in the result
    for (auto& e : A) e = make_float2(NAN, NAN);
    for (auto& e : B) e = make_float2(NAN, NAN);
    ColsCtx cx;
    cx.A = A.data();
    cx.B = B.data();
    cx.tw = T.twf.data();
    cx.w64 = T.w64.data();
    cx.k20 = grp * kColsPerCta;
    cx.ncols = kColsPerCta < kHalf - cx.k20 ? kColsPerCta : kHalf - cx.k20;
    cx.rt = reinterpret_cast<const float2*>(rt);
    cx.y = y;
    cx.outp = r_out;
    cx.t = reinterpret_cast<float2*>(t);
    for (int tid = 0; tid < kT2; ++tid) {
      R[tid].sq = R[tid].ab = 0.f;
      for (int r = 0; r < 8; ++r) R[tid].v[r] = make_float2(NAN, NAN);
      cr_load(tid, R[tid], cx);
      cr_stage_a(tid, R[tid], cx.A, cx.ncols);
    }
    for (int tid = 0; tid < kT2; ++tid) {
      cr_yload(tid, Y[tid], cx);
      cr_stage_b(tid, R[tid], cx.A, cx.B, cx.w64, cx.ncols);
    }
    for (int tid = 0; tid < kT2; ++tid) {
      if (r_out) cr_epilogue<true>(tid, R[tid], Y[tid], cx);
      else cr_epilogue<false>(tid, R[tid], Y[tid], cx);
    }
    double sq = 0.0, ab = 0.0;
    for (int tid = 0; tid < kT2; ++tid) { sq += R[tid].sq; ab += R[tid].ab; }
    partials[2 * grp] = (float)sq;
    partials[2 * grp + 1] = (float)ab;
    for (int tid = 0; tid < kT2; ++tid) {
      cr_read_a(tid, R[tid], cx.A, cx.ncols);
      cr_stage_a(tid, R[tid], cx.B, cx.ncols);
    }
    for (int tid = 0; tid < kT2; ++tid) cr_stage_b(tid, R[tid], cx.B, cx.A, cx.w64, cx.ncols);
    for (int tid = 0; tid < kT2; ++tid) cr_store(tid, R[tid], cx);
  }
  return 0;
}

// forward (two-kernel) path: rt → out = y − |F|/L (y null: |F|/L), partial sums, unit phase ph[k2][k1] (kHalf × L complex)
extern "C" int emu_cols_fwd(const float* rt, const float* y, float* out, float* ph, float* partials) {
  EmuTables T;
  std::vector<float2> A(kSeq * kLQ), B(kSeq * kLQ);
  std::vector<ColsRegs> R(kT2);
  std::vector<ColsY> Y(kT2);
  for (int grp = 0; grp < kColGroups; ++grp) {
    for (auto& e : A) e = make_float2(NAN, NAN);
    for (auto& e : B) e = make_float2(NAN, NAN);
    ColsCtx cx;
    cx.A = A.data();
    cx.B = B.data();
    cx.tw = T.twf.data();
    cx.w64 = T.w64.data();
    cx.k20 = grp * kColsPerCta;
    cx.ncols = kColsPerCta < kHalf - cx.k20 ? kColsPerCta : kHalf - cx.k20;
    cx.rt = reinterpret_cast<const float2*>(rt);
    cx.y = y;
    cx.outp = out;
    cx.t = nullptr;
    for (int tid = 0; tid < kT2; ++tid) {
      R[tid].sq = R[tid].ab = 0.f;
      for (int r = 0; r < 8; ++r) R[tid].v[r] = make_float2(NAN, NAN);
      cr_load(tid, R[tid], cx);
      cr_stage_a(tid, R[tid], cx.A, cx.ncols);
    }
    for (int tid = 0; tid < kT2; ++tid) {
      if (y) cr_yload(tid, Y[tid], cx);
      cr_stage_b(tid, R[tid], cx.A, cx.B, cx.w64, cx.ncols);
    }
    for (int tid = 0; tid < kT2; ++tid) {
      if (y) cr_fwd_epilogue<true>(tid, R[tid], Y[tid], cx, reinterpret_cast<float2*>(ph));
      else cr_fwd_epilogue<false>(tid, R[tid], Y[tid], cx, reinterpret_cast<float2*>(ph));
    }
    double sq = 0.0, ab = 0.0;
    for (int tid = 0; tid < kT2; ++tid) { sq += R[tid].sq; ab += R[tid].ab; }
    partials[2 * grp] = (float)sq;
    partials[2 * grp + 1] = (float)ab;
  }
  return 0;
}

// adjoint (two-kernel) path, columns: r plane (L×L) and unit phase ph[k2][k1] → T[row][k2] (kImg × kHalf complex)
extern "C" int emu_cols_adj(const float* rplane, const float* ph, float* t) {
  EmuTables T;
  std::vector<float2> A(kSeq * kLQ), B(kSeq * kLQ);
  std::vector<ColsRegs> R(kT2);
  for (int grp = 0; grp < kColGroups; ++grp) {
    for (auto& e : A) e = make_float2(NAN, NAN);
    for (auto& e : B) e = make_float2(NAN, NAN);
    ColsCtx cx;
    cx.A = A.data();
    cx.B = B.data();
    cx.tw = T.twf.data();
    cx.w64 = T.w64.data();
    cx.k20 = grp * kColsPerCta;
    cx.ncols = kColsPerCta < kHalf - cx.k20 ? kColsPerCta : kHalf - cx.k20;
    cx.rt = nullptr;
    cx.y = nullptr;
    cx.outp = nullptr;
    cx.t = reinterpret_cast<float2*>(t);
    for (int tid = 0; tid < kT2; ++tid) {
      for (int r = 0; r < 8; ++r) R[tid].v[r] = make_float2(NAN, NAN);
      ca_load(tid, rplane, reinterpret_cast<const float2*>(ph), cx);
    }
    for (int tid = 0; tid < kT2; ++tid) {
      cr_read_a(tid, R[tid], cx.A, cx.ncols);
      cr_stage_a(tid, R[tid], cx.B, cx.ncols);
    }
    for (int tid = 0; tid < kT2; ++tid) cr_stage_b(tid, R[tid], cx.B, cx.A, cx.w64, cx.ncols);
    for (int tid = 0; tid < kT2; ++tid) cr_store(tid, R[tid], cx);
  }
  return 0;
}
