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
