// Non-separable blur (motion kernels): ReflectionPad2d(k/2) + depthwise cross-correlation with an
// arbitrary (k,k) kernel (measurements.py:93-126, util/img_utils.py:268-308) and its exact adjoint.
//
// The motion kernels of the reference are thin paths: a few hundred non-zero taps of 3 721.  The plan
// keeps only the non-zero taps (dy, dx, w); one CTA stages a strip of the image with its halo in
// shared memory (reflect-filled for A, zero-filled for Aᵀ) and every thread accumulates 8 vertically
// adjacent outputs of one column, one LDS + one FMA per (tap, output).
//
// Adjoint with reflect padding.  out[i] = Σ_d w[d]·x[refl(i+d)] scatters w[d]·u[i] into refl(i+d), so
//   g[m] = Σ_d w[d]·( U(m−d) + [m≥1]·U(−m−d) + [m≤L−2]·U(2(L−1)−m−d) ),  U zero outside [0,L),
// per axis; the 2-D adjoint is the product of the row and column variants.  Only pixels within the
// kernel radius of a border have more than the first term, so interior CTAs run the plain path.
//
// Roofline: T_nz taps → 2·T_nz flop per pixel per direction against 8-16 B: LDS/FMA-bound, NOT
// HBM-bound, for T_nz ≳ 40; stated as such in DESIGN.md.
#include <vector>

#include "operator.cuh"

namespace {
constexpr int kThreads = 256;
constexpr int kRows = 32;
constexpr int kGroup = 8;

struct Tap {
  int dydx;  // (dy << 16) | (dx & 0xffff)
  float w;
};
}  // namespace

struct SparseTables {
  int ntaps = 0;
  int Ry = 0, Rx = 0;  // halo (Rx rounded up to a multiple of 4)
  Tap* taps_dev = nullptr;
};

namespace {

DPS_DEV int tap_dy(int v) { return v >> 16; }
DPS_DEV int tap_dx(int v) { return (int)(short)(v & 0xffff); }

template <bool kAdjoint>
__global__ void __launch_bounds__(kThreads) sparse_kernel(const Tap* __restrict__ taps_g, int ntaps, int Ry,
                                                          int Rx, int C, int H, int W, int strips,
                                                          const FwdArgs fa, const AdjArgs aa) {
  extern __shared__ __align__(16) float smem[];
  const int SW = W + 2 * Rx;
  const int tile_rows = kRows + 2 * Ry;
  float* tile = smem;
  Tap* taps = reinterpret_cast<Tap*>(tile + tile_rows * SW);
  float* red = reinterpret_cast<float*>(taps + ntaps);

  const int strip = blockIdx.x % strips;
  const int c = blockIdx.x / strips;
  const int n = blockIdx.y;
  const int r0 = strip * kRows;
  const int tid = threadIdx.x;
  const int64_t plane = (int64_t)c * H * W;

  for (int i = tid; i < ntaps; i += kThreads) taps[i] = taps_g[i];

  // ---- stage the strip: rows [r0-Ry, r0+kRows+Ry), cols [-Rx, W+Rx) ---------------------------
  {
    const float* x;
    const float* eps = nullptr;
    float c1 = 1.f, c2 = 0.f;
    int clip = 0;
    if (kAdjoint) {
      x = aa.r + (int64_t)n * C * H * W + plane;
    } else {
      x = fa.src.x + n * fa.src.x_stride + plane;
      if (fa.src.eps) eps = fa.src.eps + n * fa.src.eps_stride + plane;
      c1 = fa.src.c1; c2 = fa.src.c2; clip = fa.src.clip;
    }
    const int w4 = W / 4;
    for (int i = tid; i < tile_rows * w4; i += kThreads) {
      const int tr = i / w4, q = i - tr * w4;
      int row = r0 - Ry + tr;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (kAdjoint) {
        if (row >= 0 && row < H) v = src_load4(x, eps, (int64_t)row * W + q * 4, c1, c2, clip);
      } else {
        row = reflect_idx(row, H);
        v = src_load4(x, eps, (int64_t)row * W + q * 4, c1, c2, clip);
      }
      *reinterpret_cast<float4*>(tile + tr * SW + Rx + q * 4) = v;
    }
    __syncthreads();
    // column halos: reflect (forward) from the staged interior, zero (adjoint)
    for (int i = tid; i < tile_rows * 2 * Rx; i += kThreads) {
      const int tr = i / (2 * Rx), q = i - tr * (2 * Rx);
      const int colimg = q < Rx ? q - Rx : W + (q - Rx);  // image column of this halo cell
      float v = 0.f;
      if (!kAdjoint) v = tile[tr * SW + Rx + reflect_idx(colimg, W)];
      tile[tr * SW + Rx + colimg] = v;
    }
  }
  __syncthreads();

  const bool row_variants = kAdjoint && (r0 <= Ry || r0 + kRows - 1 >= H - 1 - Ry);
  float sq = 0.f, ab = 0.f;
  for (int col = tid; col < W; col += kThreads) {
    const bool col_variants = kAdjoint && (col <= Rx || col >= W - 1 - Rx);
#pragma unroll 1
    for (int g0 = 0; g0 < kRows; g0 += kGroup) {
      float acc[kGroup];
#pragma unroll
      for (int j = 0; j < kGroup; ++j) acc[j] = 0.f;
      if (!(row_variants || col_variants)) {
        // plain path: forward reads x̂₀[m+d], adjoint reads u[m−d]
        const float* base = tile + (g0 + Ry) * SW + Rx + col;
#pragma unroll 2
        for (int t = 0; t < ntaps; ++t) {
          const Tap tp = taps[t];
          const int dy = tap_dy(tp.dydx), dx = tap_dx(tp.dydx);
          const float* p = kAdjoint ? base - dy * SW - dx : base + dy * SW + dx;
#pragma unroll
          for (int j = 0; j < kGroup; ++j) acc[j] = fmaf(tp.w, p[j * SW], acc[j]);
        }
      } else {
        // border path (adjoint only): up to 2 row variants × 2 column variants per tap
        for (int t = 0; t < ntaps; ++t) {
          const Tap tp = taps[t];
          const int dy = tap_dy(tp.dydx), dx = tap_dx(tp.dydx);
          int xs[3];
          int nx = 0;
          xs[nx++] = col - dx;  // may fall in the zero halo
          if (col >= 1) { const int x1 = -col - dx; if (x1 >= 0 && x1 < W) xs[nx++] = x1; }
          if (col <= W - 2) { const int x2 = 2 * (W - 1) - col - dx; if (x2 >= 0 && x2 < W) xs[nx++] = x2; }
#pragma unroll
          for (int j = 0; j < kGroup; ++j) {
            const int my = r0 + g0 + j;
            int ys[3];
            int ny = 0;
            { const int y0 = my - dy; if (y0 >= 0 && y0 < H) ys[ny++] = y0; }
            if (my >= 1) { const int y1 = -my - dy; if (y1 >= 0 && y1 < H) ys[ny++] = y1; }
            if (my <= H - 2) { const int y2 = 2 * (H - 1) - my - dy; if (y2 >= 0 && y2 < H) ys[ny++] = y2; }
            float s = 0.f;
            for (int a = 0; a < ny; ++a)
              for (int b = 0; b < nx; ++b) s += tile[(ys[a] - (r0 - Ry)) * SW + Rx + xs[b]];
            acc[j] = fmaf(tp.w, s, acc[j]);
          }
        }
      }
      // ---- epilogue for these kGroup outputs of column `col` ---------------------------------
#pragma unroll
      for (int j = 0; j < kGroup; ++j) {
        const int row = r0 + g0 + j;
        if (row >= H) continue;
        const int64_t off = plane + (int64_t)row * W + col;
        if (!kAdjoint) {
          float res = acc[j];
          if (fa.y) res = __fsub_rn(fa.y[n * fa.y_stride + off], res);
          fa.out[(int64_t)n * C * H * W + off] = res;
          sq += res * res;
          ab += fabsf(res);
        } else {
          float res = (aa.coef ? aa.coef[n] : 1.0f) * acc[j];
          if (aa.extra) res += ldg_stream(aa.extra + n * aa.extra_stride + off);
          res *= mask_load(aa.mask_src, aa.has_mask, n, off);
          aa.g[n * aa.g_stride + off] = res;
        }
      }
    }
  }
  if (!kAdjoint && fa.partials) {
    block_sum2(sq, ab, red);
    if (tid == 0) {
      float* pp = fa.partials + ((int64_t)n * (C * strips) + blockIdx.x) * 2;
      pp[0] = sq;
      pp[1] = ab;
    }
  }
}

size_t sparse_smem(const dps_operator* op) {
  const SparseTables* t = op->sparse;
  return sizeof(float) * ((size_t)(kRows + 2 * t->Ry) * (op->W + 2 * t->Rx) + 64) + sizeof(Tap) * (size_t)t->ntaps;
}

template <bool kAdjoint>
int sparse_launch(const dps_operator* op, const FwdArgs& fa, const AdjArgs& aa, int n, cudaStream_t st) {
  const SparseTables* t = op->sparse;
  const size_t smem = sparse_smem(op);
  static bool attr_set = false;
  if (!attr_set) {
    DPS_CUDA(cudaFuncSetAttribute(sparse_kernel<kAdjoint>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    attr_set = true;
  }
  const int strips = (op->H + kRows - 1) / kRows;
  dim3 grid((unsigned)(op->C * strips), (unsigned)n);
  sparse_kernel<kAdjoint><<<grid, kThreads, smem, st>>>(t->taps_dev, t->ntaps, t->Ry, t->Rx, op->C, op->H,
                                                         op->W, strips, fa, aa);
  DPS_LAUNCH_CHECK(kAdjoint ? "sparse_blur_adjoint" : "sparse_blur_forward");
  return DPS_OK;
}
}  // namespace

int sparse_create(dps_operator* op, const float* kernel, int ksize) {
  const int r0 = ksize / 2;
  std::vector<Tap> taps;
  int Ry = 0, Rx = 0;
  for (int a = 0; a < ksize; ++a)
    for (int b = 0; b < ksize; ++b) {
      const float w = kernel[a * ksize + b];
      if (w == 0.0f) continue;
      const int dy = a - r0, dx = b - r0;
      taps.push_back({(int)(((unsigned)dy << 16) | ((unsigned)dx & 0xffffu)), w});
      Ry = abs(dy) > Ry ? abs(dy) : Ry;
      Rx = abs(dx) > Rx ? abs(dx) : Rx;
    }
  DPS_REQUIRE(!taps.empty(), DPS_ERR_INVALID, "blur: kernel is all zero");
  Rx = (Rx + 3) / 4 * 4;
  if (Rx == 0) Rx = 4;
  DPS_REQUIRE(Ry <= kRows - 1 && Rx <= 32, DPS_ERR_UNSUPPORTED, "sparse blur: radius (%d,%d) > 31", Ry, Rx);
  DPS_REQUIRE(op->H % kRows == 0 && op->W % 4 == 0 && op->H > Ry && op->W > Rx, DPS_ERR_UNSUPPORTED,
              "sparse blur: need H %% 32 == 0, W %% 4 == 0 and the kernel radius below the image size");
  SparseTables* t = new SparseTables();
  t->ntaps = (int)taps.size();
  t->Ry = Ry;
  t->Rx = Rx;
  op->sparse = t;
  DPS_REQUIRE(sparse_smem(op) <= 227 * 1024, DPS_ERR_UNSUPPORTED, "sparse blur: tile exceeds shared memory");
  DPS_CUDA(cudaMalloc(&t->taps_dev, taps.size() * sizeof(Tap)));
  DPS_CUDA(cudaMemcpy(t->taps_dev, taps.data(), taps.size() * sizeof(Tap), cudaMemcpyHostToDevice));
  op->P = op->C * ((op->H + kRows - 1) / kRows);
  op->taps = t->ntaps;
  return DPS_OK;
}

void sparse_destroy(dps_operator* op) {
  if (!op->sparse) return;
  cudaFree(op->sparse->taps_dev);
  delete op->sparse;
  op->sparse = nullptr;
}

int sparse_forward(const dps_operator* op, const FwdArgs& a, cudaStream_t st) {
  AdjArgs dummy = {};
  return sparse_launch<false>(op, a, dummy, a.n, st);
}
int sparse_adjoint(const dps_operator* op, const AdjArgs& a, cudaStream_t st) {
  FwdArgs dummy = {};
  return sparse_launch<true>(op, dummy, a, a.n, st);
}
