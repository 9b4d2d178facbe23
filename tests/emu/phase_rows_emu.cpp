// CPU emulation of phase_rows_fwd_reg / phase_rows_adj_reg (dps_ttc_b200/csrc/phase_rowsreg.cuh) — TEST INFRASTRUCTURE ONLY.
// Same scheme as phase_cols_emu.cpp: host shims for the device intrinsics, the kernel's own headers, the barrier-free phases
// run thread by thread (one loop over tid per barrier interval) for every row group of one plane.
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
static inline unsigned char ldg_u8_pinned(const unsigned char* p) { return *p; }
static inline void stg_u8(unsigned char* p, unsigned v) { *p = (unsigned char)v; }
static inline void stg_stream(float* p, float v) { *p = v; }
static inline void stg_stream2(float* p, const float2& v) { p[0] = v.x; p[1] = v.y; }
static inline void stg_stream4(float* p, const float4& v) { p[0] = v.x; p[1] = v.y; p[2] = v.z; p[3] = v.w; }
// common.cuh: x̂₀ = clamp(c1·x − c2·ε) with every product rounded on its own (build with -ffp-contract=off)
static inline float x0_pre(float x, float e, float c1, float c2) { return c1 * x - c2 * e; }
static inline float clamp1(float v) { return std::fmin(std::fmax(v, -1.0f), 1.0f); }
static inline float clamp_pass(float pre) { return (pre >= -1.0f && pre <= 1.0f) ? 1.0f : 0.0f; }

#include "phase_math.cuh"
#include "phase_dims.cuh"
#include "phase_colsreg.cuh"
#include "phase_rowsreg.cuh"


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
namespace {
typedef EmuTables Tables;
void poison(std::vector<float2>& b) { for (auto& e : b) e = make_float2(NAN, NAN); }
}  // namespace

extern "C" int emu_dims(int* out) {
  out[0] = kL; out[1] = kImg; out[2] = kHalf; out[3] = kImg / kRowsReg;
  return 0;
}

// x, eps: H×H planes; rt: Rt[k2][row] (kHalf × H complex); maskb: H×H bytes
extern "C" int emu_rows_fwd(const float* x, const float* eps, float c1, float c2, int clip, float* rt, unsigned char* maskb) {
  Tables T;
  std::vector<float2> A(kSeq * kLQ), B(kSeq * kLQ);
  std::vector<ColsRegs> R(kT2);
  for (int grp = 0; grp < kImg / kRowsReg; ++grp) {
    poison(A); poison(B);
    RowsFwdCtx cx;
    cx.A = A.data(); cx.B = B.data(); cx.tw = T.twf.data(); cx.w64 = T.w64.data();
    cx.x = x; cx.eps = eps; cx.c1 = c1; cx.c2 = c2; cx.clip = clip;
    cx.maskb = maskb; cx.rt = reinterpret_cast<float2*>(rt); cx.r0 = grp * kRowsReg;
    for (int tid = 0; tid < kT2; ++tid) {
      for (int r = 0; r < 8; ++r) R[tid].v[r] = make_float2(NAN, NAN);
      rf_load(tid, R[tid], cx);
      cr_stage_a(tid, R[tid], cx.A);
    }
    for (int tid = 0; tid < kT2; ++tid) cr_stage_b(tid, R[tid], cx.A, cx.B, cx.w64);
    for (int tid = 0; tid < kT2; ++tid) rf_spectrum(tid, R[tid], cx);
    for (int tid = 0; tid < kT2; ++tid) rf_split_store(tid, cx);
  }
  return 0;
}

// t: T[row][k2] (H × kHalf complex); clamp mask from maskb (H×H bytes) or recomputed from mx, meps (H×H planes) or none;
// extra: H×H plane or null; g: H×H
extern "C" int emu_rows_adj(const float* t, const unsigned char* maskb, const float* mx, const float* meps, float mc1, float mc2,
                            const float* extra, float coef, float* g) {
  Tables T;
  std::vector<float2> A(kSeq * kLQ), B(kSeq * kLQ);
  std::vector<ColsRegs> R(kT2);
  std::vector<RowsMask> M(kT2);
  for (int grp = 0; grp < kImg / kRowsReg; ++grp) {
    poison(A); poison(B);
    RowsAdjCtx cx;
    cx.A = A.data(); cx.B = B.data(); cx.tw = T.twf.data(); cx.w64 = T.w64.data();
    cx.t = reinterpret_cast<const float2*>(t); cx.maskb = maskb; cx.mx = mx; cx.meps = meps; cx.mc1 = mc1; cx.mc2 = mc2;
    cx.extra = extra; cx.g = g; cx.coef = coef; cx.r0 = grp * kRowsReg;
    for (int tid = 0; tid < kT2; ++tid) {
      for (int r = 0; r < 8; ++r) R[tid].v[r] = make_float2(NAN, NAN);
      ra_load(tid, R[tid], cx);
      cr_stage_a(tid, R[tid], cx.A);
    }
    for (int tid = 0; tid < kT2; ++tid) {
      cr_stage_b(tid, R[tid], cx.A, cx.B, cx.w64);
      ra_maskload(tid, M[tid], cx);
    }
    for (int tid = 0; tid < kT2; ++tid) ra_store(tid, R[tid], M[tid], cx);
  }
  return 0;
}
