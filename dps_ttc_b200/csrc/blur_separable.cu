// Separable blur (Gaussian): ReflectionPad2d(k/2) + depthwise cross-correlation with a rank-1 kernel
// (measurements.py:129-149, util/img_utils.py:268-308), and its exact adjoint.
//
// Formulation.  In 1-D, "reflect-pad then correlate" is the banded matrix
//     A[i][m] = Σ_{d : reflect(i+d) = m} w[d+r],   |i − m| ≤ r,
// which equals the plain taps w[m−i+r] except next to each border, where the mirrored taps fold back
// onto the band: rows i < r of A, and rows m ≤ r of Aᵀ (column r of A still receives w[0] from row 0) —
// hence R+1 border rows per side in the tables below.  Forward applies A_v ⊗ A_h, the adjoint A_vᵀ ⊗ A_hᵀ: the SAME kernel
// with different tap tables (interior taps as kernel parameters → constant-bank FFMA operands,
// border rows from a small table in shared memory).  No padded image is ever materialised.
//
// One CTA = one (particle, channel, strip of kRows output rows):
//   0. stage the strip + r halo rows in shared memory, 128-bit loads, x̂₀ = clamp(c1·x − c2·ε) applied
//      on the fly (forward) — the only global read of the particle;
//   1. vertical pass in place, one thread per column, register-blocked 8 outputs per 8+2r loads;
//   2. horizontal pass, one thread per 4 adjacent outputs, (4+2r)/4 LDS.128;
//   3. epilogue in registers: residual y − A x̂₀ + per-CTA Σr², Σ|r|   (forward)
//                            clamp mask ⊙ (coef·Aᵀr + extra)          (adjoint).
// Roofline: 2·(2r+1) FMA per pixel (50 for σ=3) against 12-16 B per pixel — the FMA pipe and HBM are
// within 1.5× of each other on B200, so this kernel is co-limited; see DESIGN.md.
#include <vector>

#include "operator.cuh"

namespace {

constexpr int kThreads = 256;
constexpr int kRows = 32;  // output rows per CTA
constexpr int kGroup = 8;  // vertical outputs per register block

template <int R>
struct SepParams {
  float wv[2 * R + 1];  // interior vertical taps, index e+R multiplies input row i+e
  float wh[2 * R + 1];
  const float* bv;  // border rows: (2(R+1), 2R+1); rows [0,R] top, then rows [L-1-R, L-1] bottom
  const float* bh;
  int C, H, W;
  int strips;  // ceil(H / kRows)
};

struct SepSet {
  std::vector<float> wv, wh;  // (2R+1)
  float* bv = nullptr;        // device
  float* bh = nullptr;
};

}  // namespace

struct SepTables {
  int R = 0;  // template radius (multiple of 4, >= true radius)
  SepSet fwd, adj;
};

namespace {

template <int R, bool kAdjoint>
__global__ void __launch_bounds__(kThreads) sep_kernel(const SepParams<R> p, const FwdArgs fa,
                                                       const AdjArgs aa) {
  extern __shared__ __align__(16) float smem[];
  const int H = p.H, W = p.W;
  const int SW = W + 2 * R;             // tile row stride (multiple of 4)
  const int tile_rows = kRows + 2 * R;  // staged rows
  float* tile = smem;
  constexpr int kBorder = 2 * (R + 1) * (2 * R + 1);
  float* bvs = tile + tile_rows * SW;
  float* bhs = bvs + kBorder;
  float* red = bhs + kBorder;  // 64 floats

  const int strip = blockIdx.x % p.strips;
  const int c = blockIdx.x / p.strips;
  const int n = blockIdx.y;
  const int r0 = strip * kRows;
  const int tid = threadIdx.x;
  const int64_t plane = (int64_t)c * H * W;

  for (int i = tid; i < kBorder; i += kThreads) {
    bvs[i] = p.bv[i];
    bhs[i] = p.bh[i];
  }

  // ---- phase 0: stage rows [r0-R, r0+kRows+R) with zero fill outside the image -----------------
  {
    const float* x;
    const float* eps = nullptr;
    float c1 = 1.f, c2 = 0.f;
    int clip = 0;
    if (kAdjoint) {
      x = aa.r + (int64_t)n * p.C * H * W + plane;
    } else {
      x = fa.src.x + n * fa.src.x_stride + plane;
      if (fa.src.eps) eps = fa.src.eps + n * fa.src.eps_stride + plane;
      c1 = fa.src.c1; c2 = fa.src.c2; clip = fa.src.clip;
    }
    const int w4 = W / 4;
    for (int i = tid; i < tile_rows * w4; i += kThreads) {
      const int tr = i / w4, q = i - tr * w4;
      const int row = r0 - R + tr;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (row >= 0 && row < H) v = src_load4(x, eps, (int64_t)row * W + q * 4, c1, c2, clip);
      *reinterpret_cast<float4*>(tile + tr * SW + R + q * 4) = v;
    }
    // zero the column halos
    for (int i = tid; i < tile_rows * 2 * R; i += kThreads) {
      const int tr = i / (2 * R), q = i - tr * (2 * R);
      tile[tr * SW + (q < R ? q : W + q)] = 0.f;
    }
  }
  __syncthreads();

  // ---- phase 1: vertical pass, in place (a column is touched by one thread only) ---------------
  for (int col = tid; col < W; col += kThreads) {
    float* colp = tile + R + col;
#pragma unroll 1
    for (int g0 = 0; g0 < kRows; g0 += kGroup) {
      float in[kGroup + 2 * R];
#pragma unroll
      for (int s = 0; s < kGroup + 2 * R; ++s) in[s] = colp[(g0 + s) * SW];
#pragma unroll
      for (int o = 0; o < kGroup; ++o) {
        const int row = r0 + g0 + o;  // image row of this output (uniform across the CTA)
        float acc = 0.f;
        if (row > R && row < H - 1 - R) {
#pragma unroll
          for (int k = 0; k <= 2 * R; ++k) acc = fmaf(p.wv[k], in[o + k], acc);
        } else if (row < H) {
          const float* bt = bvs + (row <= R ? row : (R + 1) + (row - (H - 1 - R))) * (2 * R + 1);
#pragma unroll
          for (int k = 0; k <= 2 * R; ++k) acc = fmaf(bt[k], in[o + k], acc);
        }
        colp[(g0 + o) * SW] = acc;
      }
    }
  }
  __syncthreads();

  // ---- phase 2 + epilogue: horizontal pass, 4 adjacent outputs per thread ---------------------
  const int w4 = W / 4;
  float sq = 0.f, ab = 0.f;
  for (int i = tid; i < kRows * w4; i += kThreads) {
    const int tr = i / w4, q = i - tr * w4;
    const int row = r0 + tr;
    if (row >= H) continue;
    const int col = q * 4;
    const float* rowp = tile + tr * SW + col;  // window starts at image column col-R
    float in[4 + 2 * R];
#pragma unroll
    for (int s = 0; s < (4 + 2 * R) / 4; ++s) {
      const float4 v = *reinterpret_cast<const float4*>(rowp + s * 4);
      in[s * 4 + 0] = v.x; in[s * 4 + 1] = v.y; in[s * 4 + 2] = v.z; in[s * 4 + 3] = v.w;
    }
    float o[4];
    if (col > R && col + 3 < W - 1 - R) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float acc = 0.f;
#pragma unroll
        for (int k = 0; k <= 2 * R; ++k) acc = fmaf(p.wh[k], in[j + k], acc);
        o[j] = acc;
      }
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int cc = col + j;
        float acc = 0.f;
        if (cc > R && cc < W - 1 - R) {
#pragma unroll
          for (int k = 0; k <= 2 * R; ++k) acc = fmaf(p.wh[k], in[j + k], acc);
        } else {
          const float* bt = bhs + (cc <= R ? cc : (R + 1) + (cc - (W - 1 - R))) * (2 * R + 1);
#pragma unroll
          for (int k = 0; k <= 2 * R; ++k) acc = fmaf(bt[k], in[j + k], acc);
        }
        o[j] = acc;
      }
    }
    const int64_t off = plane + (int64_t)row * W + col;
    if (!kAdjoint) {
      float4 res = make_float4(o[0], o[1], o[2], o[3]);
      if (fa.y) {
        const float4 yv = *reinterpret_cast<const float4*>(fa.y + n * fa.y_stride + off);
        res = make_float4(__fsub_rn(yv.x, res.x), __fsub_rn(yv.y, res.y), __fsub_rn(yv.z, res.z),
                          __fsub_rn(yv.w, res.w));
      }
      stg_stream4(fa.out + (int64_t)n * p.C * H * W + off, res);
      sq += res.x * res.x + res.y * res.y + res.z * res.z + res.w * res.w;
      ab += fabsf(res.x) + fabsf(res.y) + fabsf(res.z) + fabsf(res.w);
    } else {
      const float coef = aa.coef ? aa.coef[n] : 1.0f;
      float4 res = make_float4(coef * o[0], coef * o[1], coef * o[2], coef * o[3]);
      if (aa.extra) {
        const float4 e = ldg_stream4(aa.extra + n * aa.extra_stride + off);
        res.x += e.x; res.y += e.y; res.z += e.z; res.w += e.w;
      }
      const float4 pass = mask_load4(aa.mask_src, aa.has_mask, n, off);
      res.x *= pass.x; res.y *= pass.y; res.z *= pass.z; res.w *= pass.w;
      stg_stream4(aa.g + n * aa.g_stride + off, res);
    }
  }
  if (!kAdjoint && fa.partials) {
    block_sum2(sq, ab, red);
    if (tid == 0) {
      float* pp = fa.partials + ((int64_t)n * (p.C * p.strips) + blockIdx.x) * 2;
      pp[0] = sq;
      pp[1] = ab;
    }
  }
}

template <int R>
size_t sep_smem_bytes(int W) {
  return sizeof(float) * ((size_t)(kRows + 2 * R) * (W + 2 * R) + 2 * (size_t)(2 * (R + 1)) * (2 * R + 1) + 64);
}

template <int R, bool kAdjoint>
int sep_launch(const dps_operator* op, const SepSet& set, const FwdArgs& fa, const AdjArgs& aa, int n,
               cudaStream_t st) {
  SepParams<R> p;
  for (int k = 0; k <= 2 * R; ++k) {
    p.wv[k] = set.wv[k];
    p.wh[k] = set.wh[k];
  }
  p.bv = set.bv;
  p.bh = set.bh;
  p.C = op->C; p.H = op->H; p.W = op->W;
  p.strips = (op->H + kRows - 1) / kRows;
  const size_t smem = sep_smem_bytes<R>(op->W);
  static bool attr_set = false;  // per template instantiation
  if (!attr_set) {
    DPS_CUDA(cudaFuncSetAttribute(sep_kernel<R, kAdjoint>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)smem > 48 * 1024 ? 227 * 1024 : 48 * 1024));
    attr_set = true;
  }
  dim3 grid((unsigned)(p.C * p.strips), (unsigned)n);
  sep_kernel<R, kAdjoint><<<grid, kThreads, smem, st>>>(p, fa, aa);
  DPS_LAUNCH_CHECK(kAdjoint ? "sep_blur_adjoint" : "sep_blur_forward");
  return DPS_OK;
}

template <bool kAdjoint>
int sep_dispatch(const dps_operator* op, const FwdArgs& fa, const AdjArgs& aa, int n, cudaStream_t st) {
  const SepTables* t = op->sep;
  const SepSet& set = kAdjoint ? t->adj : t->fwd;
  switch (t->R) {
    case 4: return sep_launch<4, kAdjoint>(op, set, fa, aa, n, st);
    case 8: return sep_launch<8, kAdjoint>(op, set, fa, aa, n, st);
    case 12: return sep_launch<12, kAdjoint>(op, set, fa, aa, n, st);
    case 16: return sep_launch<16, kAdjoint>(op, set, fa, aa, n, st);
    case 24: return sep_launch<24, kAdjoint>(op, set, fa, aa, n, st);
    case 32: return sep_launch<32, kAdjoint>(op, set, fa, aa, n, st);
  }
  dps_set_error("separable blur: unsupported radius %d", t->R);
  return DPS_ERR_UNSUPPORTED;
}

// Dense 1-D operator matrix of "reflect-pad r then correlate with w (2r+1 taps)" on length L.
std::vector<double> band_matrix(const std::vector<double>& w, int r, int L) {
  std::vector<double> A((size_t)L * L, 0.0);
  for (int i = 0; i < L; ++i)
    for (int d = -r; d <= r; ++d) {
      int m = i + d;
      if (m < 0) m = -m;
      if (m >= L) m = 2 * (L - 1) - m;
      A[(size_t)i * L + m] += w[d + r];
    }
  return A;
}

// interior + border tap tables of A (transpose=false) or Aᵀ (transpose=true), zero-padded to radius R
int build_set(const std::vector<double>& wv, int rv, int H, const std::vector<double>& wh, int rh, int W,
              int R, bool transpose, SepSet* out) {
  auto build = [&](const std::vector<double>& w, int r, int L, std::vector<float>* interior,
                   float** border_dev) -> int {
    std::vector<double> A = band_matrix(w, r, L);
    auto at = [&](int i, int m) -> double {  // operator entry: output i, input m
      if (m < 0 || m >= L) return 0.0;
      return transpose ? A[(size_t)m * L + i] : A[(size_t)i * L + m];
    };
    interior->assign(2 * R + 1, 0.f);
    const int mid = L / 2;  // an interior row (L >= 2R+1 checked by the caller)
    for (int e = -R; e <= R; ++e) (*interior)[e + R] = (float)at(mid, mid + e);
    std::vector<float> border((size_t)2 * (R + 1) * (2 * R + 1), 0.f);
    for (int b = 0; b < 2 * (R + 1); ++b) {
      const int i = b <= R ? b : (L - 1 - R) + (b - (R + 1));
      for (int e = -R; e <= R; ++e) border[(size_t)b * (2 * R + 1) + e + R] = (float)at(i, i + e);
    }
    DPS_CUDA(cudaMalloc(border_dev, border.size() * sizeof(float)));
    DPS_CUDA(cudaMemcpy(*border_dev, border.data(), border.size() * sizeof(float), cudaMemcpyHostToDevice));
    return DPS_OK;
  };
  if (int rc = build(wv, rv, H, &out->wv, &out->bv)) return rc;
  if (int rc = build(wh, rh, W, &out->wh, &out->bh)) return rc;
  return DPS_OK;
}

}  // namespace

// taps1d_v / taps1d_h: (2rv+1) / (2rh+1) cross-correlation taps, index d+r multiplies x[i+d]
int sep_create(dps_operator* op, const float* taps1d_v, const float* taps1d_h, int rv, int rh) {
  const int r = rv > rh ? rv : rh;
  int R = 0;
  for (int cand : {4, 8, 12, 16, 24, 32})
    if (cand >= r) { R = cand; break; }
  DPS_REQUIRE(R > 0, DPS_ERR_UNSUPPORTED, "separable blur: radius %d > 32", r);
  DPS_REQUIRE(op->H >= 2 * R + 3 && op->W >= 2 * R + 4 && op->W % 4 == 0, DPS_ERR_UNSUPPORTED,
              "separable blur: image %dx%d too small / W not a multiple of 4 for radius %d", op->H, op->W, R);
  size_t smem = sizeof(float) * ((size_t)(kRows + 2 * R) * (op->W + 2 * R) + 2 * (size_t)(2 * (R + 1)) * (2 * R + 1) + 64);
  DPS_REQUIRE(smem <= 227 * 1024, DPS_ERR_UNSUPPORTED, "separable blur: tile of %zu bytes exceeds shared memory", smem);
  std::vector<double> wv(taps1d_v, taps1d_v + 2 * rv + 1), wh(taps1d_h, taps1d_h + 2 * rh + 1);
  SepTables* t = new SepTables();
  t->R = R;
  op->sep = t;
  if (int rc = build_set(wv, rv, op->H, wh, rh, op->W, R, false, &t->fwd)) return rc;
  if (int rc = build_set(wv, rv, op->H, wh, rh, op->W, R, true, &t->adj)) return rc;
  op->P = op->C * ((op->H + kRows - 1) / kRows);
  op->taps = 2 * r + 1;
  return DPS_OK;
}

void sep_destroy(dps_operator* op) {
  if (!op->sep) return;
  for (SepSet* s : {&op->sep->fwd, &op->sep->adj}) {
    cudaFree(s->bv);
    cudaFree(s->bh);
  }
  delete op->sep;
  op->sep = nullptr;
}

int sep_forward(const dps_operator* op, const FwdArgs& a, cudaStream_t st) {
  AdjArgs dummy = {};
  return sep_dispatch<false>(op, a, dummy, a.n, st);
}
int sep_adjoint(const dps_operator* op, const AdjArgs& a, cudaStream_t st) {
  FwdArgs dummy = {};
  return sep_dispatch<true>(op, dummy, a, a.n, st);
}
