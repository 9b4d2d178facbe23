// Size-dependent constants and index helpers of the phase-retrieval kernels for one transform length L = 64*PHASE_R3.
// Host/device-clean (constexpr + DPS_DEV functions): phase_impl.cuh includes it once per length inside a namespace, and
// tests/emu/phase_cols_emu.cpp includes it with host shims to run the register-FFT column kernel on the CPU.
constexpr int kR3 = PHASE_R3;
constexpr int kL = 64 * kR3;
constexpr int kL8 = kL / 8;           // butterflies per sequence in the two radix-8 stages
constexpr int kS8 = 9 * kR3;          // P(j + kL8·r) − P(j): shared-memory stride of a radix-8 butterfly's inputs
constexpr int kHalf = kL / 2 + 1;     // 193 / 129 / 97
constexpr int kLP = kL + kL / 8 + 1;  // padded length of a sequence in shared memory (433 / 289 / 217): ODD so that the same element of
                                      // consecutive sequences falls into different banks (loops that run over the sequence index)
constexpr int kLF = kL + 1;           // row stride of the float planes staged per sequence (same reason)
constexpr int kPad = 64;              // int((oversample / 8) · 256) with oversample = 2, whatever the image size (measurements.py:181)
constexpr int kImg = kL - 2 * kPad;
constexpr int kRowsPerCta = 16;  // K1 / A2: image rows per CTA → 8 packed FFTs (55.9 KB of shared memory at L = 384 → 4 CTAs per SM)
constexpr int kColsPerCta = 8;   // K2 / A1: spectrum columns per CTA
constexpr int kColGroups = (kHalf + kColsPerCta - 1) / kColsPerCta;  // 25 / 17 / 13
// the adjoint prefers wider CTAs: its scattered reads of r coalesce into 64-byte runs with 16 columns
constexpr int kRowsAdj = 32;
constexpr int kColsAdj = 16;
constexpr int kColGroupsAdj = (kHalf + kColsAdj - 1) / kColsAdj;  // 13 / 9 / 7
static_assert(kL8 % 8 == 0 && kImg % kRowsAdj == 0 && kImg % kRowsPerCta == 0 && kL % 32 == 0, "supported lengths");

// Shared-memory sequences are padded by one element every 8: element i lives at i + (i >> 3).  With 8-byte
// elements this makes the stride-8 / stride-64 scatter of the Stockham stages conflict-free (stride 9 / 72).
DPS_DEV int P(int i) { return i + (i >> 3); }

// last stage: radix kR3
DPS_DEV void dft_last(float2* v) {
  if constexpr (kR3 == 6) dft6(v);
  else if constexpr (kR3 == 4) dft4(v);
  else dft3(v[0], v[1], v[2]);
}

template <bool kPackedMul>
DPS_DEV void dft_last_p(float2* v) {  // packed variant (phase_math.cuh)
  if constexpr (kR3 == 6) dft6p<kPackedMul>(v);
  else if constexpr (kR3 == 4) dft4p(v);
  else dft3p(v[0], v[1], v[2]);
}

// Twiddle exp(−2πi j/384), j ∈ [0,384), from the half table in shared memory: tw[j+192] = −tw[j].  Halving the table
// (1.5 KB instead of 3 KB) is what lets FOUR 8-sequence CTAs (57 KB each) share an SM instead of three.
constexpr int kTW = kL / 2;
DPS_DEV float2 twid(const float2* tw, int j) {
  const bool hi = j >= kTW;
  const float2 t = tw[hi ? j - kTW : j];
  return hi ? make_float2(-t.x, -t.y) : t;
}

DPS_DEV int shift_idx(int k) { return k + kL / 2 >= kL ? k - kL / 2 : k + kL / 2; }  // fftshift position of bin k

