// Register-resident row kernels of the fused phase-retrieval guidance (included by phase_impl.cuh once per transform length,
// after phase_colsreg.cuh whose J-role stages cr_stage_a / cr_stage_b and exchange layout they share):
//   K1  image rows (x, ε → x̂₀, two real rows per complex sequence) → half spectrum Rt[k2][row] + clamp-pass bytes
//   K3  T[row][k2] (Hermitian half rows) → masked cotangent rows g
// Same idea as the column kernel: global data goes straight into the registers of the thread that owns the first-stage
// butterfly (no staging pass), shared memory is only the exchange between stages.  A CTA of 512 threads owns 16 image rows
// = 8 sequences.
//   J role (radix-8 stages): sequence f = tid / (L/8), butterfly j = tid % (L/8) → a warp's global accesses run along a row;
//   G role (last stage, all 512 threads): sequence f = tid >> 6, butterfly j = tid & 63 → bins j + 64r: a warp again runs
//          along a row, so K3's stores are 128-byte runs; K1 puts the spectrum back into A in natural order and a last pass
//          (8 consecutive lanes = the 8 sequences of one bin) splits Z into the two rows' spectra: 128-byte runs of Rt.
// Every function is a barrier-free phase taking the thread index (tests/emu/phase_rows_emu.cpp runs them on the CPU).
struct RowsFwdCtx {
  float2* A;
  float2* B;
  const tw_t* tw;        // full table (see ColsCtx)
  const tw_t* w64;
  const float* x;        // plane of this particle and channel (H×H)
  const float* eps;      // plane; null: x̂₀ = x (forward path only, operator.forward(x))
  float c1, c2;
  int clip;
  unsigned char* maskb;  // clamp-pass bytes of the plane (1 = gradient passes); null: not wanted (forward path)
  float2* rt;            // Rt[k2][row] of the plane
  int r0;                // first image row of the CTA
};
constexpr int kRowsReg = 2 * kSeq;  // image rows per CTA
static_assert(kImg % kRowsReg == 0, "row groups");

// slices r of the first stage whose positions j + (L/8)·r are all zero padding
DPS_DEV constexpr bool rr_all_pad(int r) { return kL8 * r + kL8 - 1 < kPad || kL8 * r >= kPad + kImg; }

// K1, J role: x̂₀ of the row pair at columns j + (L/8)·r − 64 (packed as re / im), clamp-pass bytes written on the way
DPS_DEV void rf_load(int tid, ColsRegs& R, const RowsFwdCtx& c) {
  if (tid >= kJ) return;
  const int f = tid / kL8, j = tid - f * kL8;
  const int64_t row0 = (int64_t)(c.r0 + 2 * f) * kImg;
  float xa[8], ea[8], xb[8], eb[8];
  const bool he = c.eps != nullptr;
#pragma unroll
  for (int r = 0; r < 8; ++r) {
    if (rr_all_pad(r)) continue;
    const int col = j + kL8 * r - kPad;
    const bool ok = col >= 0 && col < kImg;
    const int64_t off = row0 + col;
    xa[r] = ok ? ldg_stream(c.x + off) : 0.f;
    ea[r] = (ok && he) ? ldg_stream(c.eps + off) : 0.f;
    xb[r] = ok ? ldg_stream(c.x + off + kImg) : 0.f;
    eb[r] = (ok && he) ? ldg_stream(c.eps + off + kImg) : 0.f;
  }
  const bool clip = he && c.clip;  // x̂₀ = x is never clamped (x0_of / src_load)
#pragma unroll
  for (int r = 0; r < 8; ++r) {
    if (rr_all_pad(r)) {
      R.v[r] = make_float2(0.f, 0.f);
      continue;
    }
    const int col = j + kL8 * r - kPad;
    const bool ok = col >= 0 && col < kImg;
    const float pa = he ? x0_pre(xa[r], ea[r], c.c1, c.c2) : xa[r], pb = he ? x0_pre(xb[r], eb[r], c.c1, c.c2) : xb[r];
    R.v[r] = ok ? make_float2(clip ? clamp1(pa) : pa, clip ? clamp1(pb) : pb) : make_float2(0.f, 0.f);
    if (ok && c.maskb) {
      const int64_t off = row0 + col;
      stg_u8(c.maskb + off, (!clip || clamp_pass(pa) != 0.f) ? 1u : 0u);
      stg_u8(c.maskb + off + kImg, (!clip || clamp_pass(pb) != 0.f) ? 1u : 0u);
    }
  }
}
// K1 / K3, G role: last stage (R = L/64, Ns = 64) of sequence tid >> 6, bins (tid & 63) + 64r in natural order
DPS_DEV void rr_stage_c(int tid, float2* v, const float2* srcbuf, const tw_t* tw) {
  const int f = tid >> 6, j = tid & 63;
  const float2* src = srcbuf + f * kLQ + P(j);
#pragma unroll
  for (int r = 0; r < kR3; ++r) {
    v[r] = src[72 * r];
    if (r) v[r] = cr_cmul(v[r], tw[j * r]);
  }
  cr_dft_last(v);
}
// K1, G role: spectrum Z of the packed row pair back into a buffer, natural order
DPS_DEV void rf_spectrum(int tid, ColsRegs& R, const RowsFwdCtx& c) {
  rr_stage_c(tid, R.v, c.B, c.tw);
  const int f = tid >> 6, j = tid & 63;
  float2* dst = c.A + f * kLQ + P(j);
#pragma unroll
  for (int r = 0; r < kR3; ++r) dst[72 * r] = R.v[r];
}
// K1: Z = A + iB (A, B the spectra of the even / odd row): A[k] = ½(Z[k] + conj Z[−k]), B[k] = (Z[k] − conj Z[−k]) / 2i for
// k ≤ L/2 → Rt[k][row], Rt[k][row + 1]
DPS_DEV void rf_split_store(int tid, const RowsFwdCtx& c) {
  constexpr int kItems = kHalf * kSeq;
#pragma unroll
  for (int q = 0; q < (kItems + kT2 - 1) / kT2; ++q) {
    const int i = tid + q * kT2;
    if (i < kItems) {
      const int k = i >> 3, f = i & 7;
      const float2* zb = c.A + f * kLQ;
      const float2 z = zb[P(k)];
      const float2 zc = cconj(zb[P(k ? kL - k : 0)]);
      const float2 a = make_float2(0.5f * (z.x + zc.x), 0.5f * (z.y + zc.y));
      const float2 d = make_float2(0.5f * (z.x - zc.x), 0.5f * (z.y - zc.y));
      stg_stream4(reinterpret_cast<float*>(c.rt + (int64_t)k * kImg + c.r0 + 2 * f), make_float4(a.x, a.y, d.y, -d.x));
    }
  }
}
static_assert(kSeq == 8, "rf_split_store: i & 7");

struct RowsAdjCtx {
  float2* A;
  float2* B;
  const tw_t* tw;
  const tw_t* w64;
  const float2* t;             // T[row][k2] of the plane, row stride L/2 + 1
  const unsigned char* maskb;  // clamp-pass bytes of the plane (fused path) or null
  const float* mx;             // two-kernel path: planes of x and ε the clamp mask is recomputed from (null: no mask)
  const float* meps;
  float mc1, mc2;
  const float* extra;          // plane added before the mask (null: none)
  float* g;                    // cotangent plane (H×H)
  float coef;                  // 1/L (× the per-particle coefficient where one is given)
  int r0;
};
struct RowsMask {
  float a[kR3], b[kR3];    // 1 = gradient passes
  float ea[kR3], eb[kR3];  // extra term
};

// K3, J role: X[k] = T1[k] + i·T2[k] of the row pair, T[L − k] = conj(T[k]) for the upper half
DPS_DEV void ra_load(int tid, ColsRegs& R, const RowsAdjCtx& c) {
  if (tid >= kJ) return;
  const int f = tid / kL8, j = tid - f * kL8;
  const float2* t1 = c.t + (int64_t)(c.r0 + 2 * f) * kHalf;
  const float2* t2 = t1 + kHalf;
  float2 u1[8], u2[8];
#pragma unroll
  for (int r = 0; r < 8; ++r) {
    const int pos = j + kL8 * r;
    const int kk = pos < kHalf ? pos : kL - pos;
    u1[r] = ldg_stream2(t1 + kk);
    u2[r] = ldg_stream2(t2 + kk);
  }
#pragma unroll
  for (int r = 0; r < 8; ++r) {
    const int pos = j + kL8 * r;
    const float sg = pos >= kHalf ? -1.f : 1.f;  // conj for the mirrored half
    R.v[r] = make_float2(u1[r].x - sg * u2[r].y, sg * u1[r].y + u2[r].x);
  }
}
// K3, G role: the clamp mask (bytes of the fused path, or recomputed from x and ε) and the extra term of the thread's outputs,
// requested before the last barrier
DPS_DEV void ra_maskload(int tid, RowsMask& M, const RowsAdjCtx& c) {
  const int f = tid >> 6, j = tid & 63;
  const int64_t o = (int64_t)(c.r0 + 2 * f) * kImg + j - kPad;
#pragma unroll
  for (int r = 0; r < kR3; ++r) {
    if (64 * r >= kPad && 64 * r + 63 < kPad + kImg) {
      if (c.maskb) {
        M.a[r] = ldg_u8_pinned(c.maskb + o + 64 * r) ? 1.f : 0.f;
        M.b[r] = ldg_u8_pinned(c.maskb + o + 64 * r + kImg) ? 1.f : 0.f;
      } else if (c.mx) {
        M.a[r] = clamp_pass(x0_pre(ldg_stream(c.mx + o + 64 * r), ldg_stream(c.meps + o + 64 * r), c.mc1, c.mc2));
        M.b[r] = clamp_pass(x0_pre(ldg_stream(c.mx + o + 64 * r + kImg), ldg_stream(c.meps + o + 64 * r + kImg), c.mc1, c.mc2));
      } else {
        M.a[r] = M.b[r] = 1.f;
      }
      M.ea[r] = c.extra ? ldg_stream(c.extra + o + 64 * r) : 0.f;
      M.eb[r] = c.extra ? ldg_stream(c.extra + o + 64 * r + kImg) : 0.f;
    }
  }
}
// K3, G role: last stage; Re → even row, Im → odd row, padded columns 64 .. 64 + H − 1: (coefficient · value + extra) × clamp mask
DPS_DEV void ra_store(int tid, ColsRegs& R, const RowsMask& M, const RowsAdjCtx& c) {
  rr_stage_c(tid, R.v, c.B, c.tw);
  const int f = tid >> 6, j = tid & 63;
  float* g = c.g + (int64_t)(c.r0 + 2 * f) * kImg + j - kPad;
#pragma unroll
  for (int r = 0; r < kR3; ++r) {
    if (64 * r >= kPad && 64 * r + 63 < kPad + kImg) {
      stg_stream(g + 64 * r, (c.coef * R.v[r].x + M.ea[r]) * M.a[r]);
      stg_stream(g + 64 * r + kImg, (c.coef * R.v[r].y + M.eb[r]) * M.b[r]);
    }
  }
}
