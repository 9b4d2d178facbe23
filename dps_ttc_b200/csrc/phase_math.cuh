// Complex helpers and the in-register DFT kernels (radix 8, 6, 4, 3; forward sign) of the phase-retrieval FFTs.
// Included inside phase.cu's anonymous namespace; tests/emu/phase_cols_emu.cpp includes it with host shims for DPS_DEV,
// float2 and make_float2 to run the column kernel's index logic on the CPU (test infrastructure, never a product path).
DPS_DEV float2 cmul(float2 a, float2 b) { return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
DPS_DEV float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
DPS_DEV float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
DPS_DEV float2 cconj(float2 a) { return make_float2(a.x, -a.y); }
DPS_DEV float2 mul_mi(float2 a) { return make_float2(a.y, -a.x); }  // a·(−i)

// in-register DFT-8, forward sign, natural order in and out
DPS_DEV void dft8(float2* v) {
  const float h = 0.70710678118654752440f;
  // stage 1 (stride 4)
  float2 a0 = cadd(v[0], v[4]), a4 = csub(v[0], v[4]);
  float2 a1 = cadd(v[1], v[5]), a5 = csub(v[1], v[5]);
  float2 a2 = cadd(v[2], v[6]), a6 = csub(v[2], v[6]);
  float2 a3 = cadd(v[3], v[7]), a7 = csub(v[3], v[7]);
  // twiddles on the odd half: W8^0, W8^1, W8^2, W8^3
  a5 = make_float2(h * (a5.x + a5.y), h * (a5.y - a5.x));   // ·(1−i)/√2
  a6 = mul_mi(a6);                                          // ·(−i)
  a7 = make_float2(h * (a7.y - a7.x), -h * (a7.x + a7.y));  // ·(−1−i)/√2
  // stage 2 (two DFT-4 halves)
  float2 b0 = cadd(a0, a2), b2 = csub(a0, a2);
  float2 b1 = cadd(a1, a3), b3 = mul_mi(csub(a1, a3));
  float2 b4 = cadd(a4, a6), b6 = csub(a4, a6);
  float2 b5 = cadd(a5, a7), b7 = mul_mi(csub(a5, a7));
  // stage 3 → natural order: even outputs from the first half, odd outputs from the second
  v[0] = cadd(b0, b1); v[4] = csub(b0, b1);
  v[2] = cadd(b2, b3); v[6] = csub(b2, b3);
  v[1] = cadd(b4, b5); v[5] = csub(b4, b5);
  v[3] = cadd(b6, b7); v[7] = csub(b6, b7);
}

// in-register DFT-3 (forward sign)
DPS_DEV void dft3(float2& x0, float2& x1, float2& x2) {
  const float s = 0.86602540378443864676f;  // sin(2π/3)
  const float2 t = cadd(x1, x2);
  const float2 d = csub(x1, x2);
  const float2 m = make_float2(x0.x - 0.5f * t.x, x0.y - 0.5f * t.y);
  const float2 r = make_float2(s * d.y, -s * d.x);  // −i·s·d
  x0 = cadd(x0, t);
  x1 = cadd(m, r);
  x2 = csub(m, r);
}

// in-register DFT-6 (forward sign): V[q] = E[q mod 3] + W6^q·O[q mod 3]
DPS_DEV void dft6(float2* v) {
  float2 e0 = v[0], e1 = v[2], e2 = v[4];
  float2 o0 = v[1], o1 = v[3], o2 = v[5];
  dft3(e0, e1, e2);
  dft3(o0, o1, o2);
  const float s = 0.86602540378443864676f;
  const float2 w1 = make_float2(0.5f, -s), w2 = make_float2(-0.5f, -s);  // W6^1, W6^2
  const float2 t1 = cmul(o1, w1), t2 = cmul(o2, w2);
  v[0] = cadd(e0, o0); v[3] = csub(e0, o0);   // W6^3 = −1
  v[1] = cadd(e1, t1); v[4] = csub(e1, t1);   // W6^4 = −W6^1
  v[2] = cadd(e2, t2); v[5] = csub(e2, t2);   // W6^5 = −W6^2
}

// in-register DFT-4 (forward sign)
DPS_DEV void dft4(float2* v) {
  const float2 a0 = cadd(v[0], v[2]), a1 = csub(v[0], v[2]);
  const float2 a2 = cadd(v[1], v[3]), a3 = mul_mi(csub(v[1], v[3]));
  v[0] = cadd(a0, a2); v[2] = csub(a0, a2);
  v[1] = cadd(a1, a3); v[3] = csub(a1, a3);
}

// ---- packed variants for the register-resident kernels (PHASE_PACKED) ---------------------------------------------------------
// A complex number is an aligned register pair, so a complex add / subtract is ONE FADD2 (the negation folds into the operand),
// and a multiplication by a twiddle stored as (w.x, w.y, −w.y, w.x) is FMUL2 + FFMA2 with the two halves of the data as
// scalar-broadcast operands: a·w = w·a.x + (−w.y, w.x)·a.y.  Additions by ∓i·z keep their scalar form (the swap of the halves
// would cost moves).  Every value is formed by the same operations in the same order as in dft8 / dft6 / dft4 / dft3 above,
// except the twiddle product, whose second product (not the first) is the rounded one.
DPS_DEV float2 padd(float2 a, float2 b) { return __fadd2_rn(a, b); }
DPS_DEV float2 psub(float2 a, float2 b) { return __fadd2_rn(a, make_float2(-b.x, -b.y)); }
DPS_DEV float2 cmul_tw(float2 a, float4 w) {
  const float2 m = __fmul2_rn(make_float2(w.z, w.w), make_float2(a.y, a.y));
  return __ffma2_rn(make_float2(w.x, w.y), make_float2(a.x, a.x), m);
}
DPS_DEV float2 add_mi(float2 a, float2 d) { return make_float2(a.x + d.y, a.y - d.x); }  // a + (−i)·d
DPS_DEV float2 sub_mi(float2 a, float2 d) { return make_float2(a.x - d.y, a.y + d.x); }  // a − (−i)·d

DPS_DEV void dft8p(float2* v) {
  const float h = 0.70710678118654752440f;
  const float2 a0 = padd(v[0], v[4]), a4 = psub(v[0], v[4]);
  const float2 a1 = padd(v[1], v[5]), q5 = psub(v[1], v[5]);
  const float2 a2 = padd(v[2], v[6]), a6 = psub(v[2], v[6]);
  const float2 a3 = padd(v[3], v[7]), q7 = psub(v[3], v[7]);
  const float2 a5 = make_float2(h * (q5.x + q5.y), h * (q5.y - q5.x));   // ·(1−i)/√2
  const float2 a7 = make_float2(h * (q7.y - q7.x), -h * (q7.x + q7.y));  // ·(−1−i)/√2
  const float2 b0 = padd(a0, a2), b2 = psub(a0, a2);
  const float2 b1 = padd(a1, a3), d13 = psub(a1, a3);
  const float2 b4 = add_mi(a4, a6), b6 = sub_mi(a4, a6);
  const float2 b5 = padd(a5, a7), d57 = psub(a5, a7);
  v[0] = padd(b0, b1); v[4] = psub(b0, b1);
  v[2] = add_mi(b2, d13); v[6] = sub_mi(b2, d13);
  v[1] = padd(b4, b5); v[5] = psub(b4, b5);
  v[3] = add_mi(b6, d57); v[7] = sub_mi(b6, d57);
}
DPS_DEV void dft3p(float2& x0, float2& x1, float2& x2) {
  const float s = 0.86602540378443864676f;
  const float2 t = padd(x1, x2);
  const float2 d = psub(x1, x2);
  const float2 m = __ffma2_rn(t, make_float2(-0.5f, -0.5f), x0);  // x0 − t/2 (the halving is exact)
  const float2 r = make_float2(s * d.y, -s * d.x);                 // −i·s·d
  x0 = padd(x0, t);
  x1 = padd(m, r);
  x2 = psub(m, r);
}
template <bool kPackedMul>
DPS_DEV void dft6p(float2* v) {
  float2 e0 = v[0], e1 = v[2], e2 = v[4];
  float2 o0 = v[1], o1 = v[3], o2 = v[5];
  dft3p(e0, e1, e2);
  dft3p(o0, o1, o2);
  const float s = 0.86602540378443864676f;
  // ·W6^1 = ½ − i·s, ·W6^2 = −½ − i·s
  const float2 t1 = kPackedMul ? cmul_tw(o1, make_float4(0.5f, -s, s, 0.5f)) : cmul(o1, make_float2(0.5f, -s));
  const float2 t2 = kPackedMul ? cmul_tw(o2, make_float4(-0.5f, -s, s, -0.5f)) : cmul(o2, make_float2(-0.5f, -s));
  v[0] = padd(e0, o0); v[3] = psub(e0, o0);
  v[1] = padd(e1, t1); v[4] = psub(e1, t1);
  v[2] = padd(e2, t2); v[5] = psub(e2, t2);
}
DPS_DEV void dft4p(float2* v) {
  const float2 a0 = padd(v[0], v[2]), a1 = psub(v[0], v[2]);
  const float2 a2 = padd(v[1], v[3]), d = psub(v[1], v[3]);
  v[0] = padd(a0, a2); v[2] = psub(a0, a2);
  v[1] = add_mi(a1, d); v[3] = sub_mi(a1, d);
}
