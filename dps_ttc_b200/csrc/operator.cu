// Operator plans: creation, validation and dispatch of forward / adjoint launches.
#include <math.h>

#include <vector>

#include "operator.cuh"

int inpaint_partials(int C, int H, int W);  // inpaint.cu

namespace {

int new_op(int kind, int C, int H, int W, dps_operator** out, const char* who) {
  DPS_REQUIRE(out, DPS_ERR_INVALID, "%s: null output handle", who);
  DPS_REQUIRE(C > 0 && H > 0 && W > 0, DPS_ERR_INVALID, "%s: bad shape (%d,%d,%d)", who, C, H, W);
  DPS_REQUIRE(((int64_t)H * W) % 4 == 0, DPS_ERR_UNSUPPORTED, "%s: H*W must be a multiple of 4", who);
  dps_operator* op = new dps_operator();
  op->kind = kind;
  op->C = op->oC = C;
  op->H = op->oH = H;
  op->W = op->oW = W;
  cudaError_t e = cudaGetDevice(&op->device);
  if (e != cudaSuccess) {
    delete op;
    dps_set_error("%s: cudaGetDevice failed: %s", who, cudaGetErrorString(e));
    return DPS_ERR_CUDA;
  }
  *out = op;
  return DPS_OK;
}

// The plan's tables live on op->device and the launch goes to the CALLER's current device: they must be the same one.
int check_device(const dps_operator* op, const char* who) {
  int cur = -1;
  DPS_CUDA(cudaGetDevice(&cur));
  DPS_REQUIRE(cur == op->device, DPS_ERR_INVALID, "%s: the operator was created on device %d but the current device is %d", who,
              op->device, cur);
  return DPS_OK;
}

int fail(dps_operator** out, int rc) {
  if (out && *out) {
    dps_operator_destroy(*out);
    *out = nullptr;
  }
  return rc;
}

}  // namespace

extern "C" {

int dps_operator_create_inpainting(const float* mask_host, int C, int H, int W, dps_operator** out) {
  DPS_REQUIRE(mask_host, DPS_ERR_INVALID, "dps_operator_create_inpainting: null mask");
  if (int rc = new_op(DPS_OP_INPAINT, C, H, W, out, "dps_operator_create_inpainting")) return rc;
  dps_operator* op = *out;
  cudaError_t e = cudaMalloc(&op->mask_dev, sizeof(float) * H * W);
  if (e == cudaSuccess) e = cudaMemcpy(op->mask_dev, mask_host, sizeof(float) * H * W, cudaMemcpyHostToDevice);
  if (e != cudaSuccess) {
    dps_set_error("dps_operator_create_inpainting: %s", cudaGetErrorString(e));
    return fail(out, DPS_ERR_CUDA);
  }
  op->P = inpaint_partials(C, H, W);
  op->guidance_P = op->P;  // dps_operator_guidance: one streaming kernel (inpaint_guidance)
  return DPS_OK;
}

int dps_operator_create_blur(const float* kernel_host, int ksize, int C, int H, int W, int mode, dps_operator** out) {
  DPS_REQUIRE(kernel_host && ksize > 0 && (ksize % 2) == 1, DPS_ERR_INVALID,
              "dps_operator_create_blur: kernel must be (k,k) with k odd");
  DPS_REQUIRE(mode >= 0 && mode <= 2, DPS_ERR_INVALID, "dps_operator_create_blur: bad mode %d", mode);
  // rank-1 test: K ≈ rowsum ⊗ colsum / total
  const int r0 = ksize / 2;
  std::vector<double> rs(ksize, 0.0), cs(ksize, 0.0);
  double total = 0.0, kmax = 0.0;
  for (int a = 0; a < ksize; ++a)
    for (int b = 0; b < ksize; ++b) {
      const double v = kernel_host[a * ksize + b];
      rs[a] += v;
      cs[b] += v;
      total += v;
      kmax = fmax(kmax, fabs(v));
    }
  bool separable = total != 0.0 && kmax > 0.0;
  if (separable) {
    double err = 0.0;
    for (int a = 0; a < ksize; ++a)
      for (int b = 0; b < ksize; ++b) err = fmax(err, fabs(rs[a] * cs[b] / total - (double)kernel_host[a * ksize + b]));
    separable = err <= 2e-7 * kmax;  // fp32 rounding of the (k,k) weights themselves is ~6e-8 relative
  }
  DPS_REQUIRE(!(mode == 1 && !separable), DPS_ERR_UNSUPPORTED, "dps_operator_create_blur: kernel is not rank-1");
  const bool use_sep = (mode == 1) || (mode == 0 && separable);
  if (int rc = new_op(use_sep ? DPS_OP_BLUR_SEPARABLE : DPS_OP_BLUR_SPARSE, C, H, W, out, "dps_operator_create_blur"))
    return rc;
  dps_operator* op = *out;
  int rc;
  if (use_sep) {
    // 1-D taps: v[a]·h[b] = rs[a]·cs[b]/total ; split the normalisation symmetrically
    const double s = sqrt(fabs(total));
    const double sign = total < 0 ? -1.0 : 1.0;
    std::vector<float> tv(ksize), th(ksize);
    int rv = 0, rh = 0;
    for (int a = 0; a < ksize; ++a) {
      tv[a] = (float)(rs[a] / s);
      th[a] = (float)(sign * cs[a] / s);
      if (tv[a] != 0.f) rv = abs(a - r0) > rv ? abs(a - r0) : rv;
      if (th[a] != 0.f) rh = abs(a - r0) > rh ? abs(a - r0) : rh;
    }
    rc = sep_create(op, tv.data() + (r0 - rv), th.data() + (r0 - rh), rv, rh);
  } else {
    rc = sparse_create(op, kernel_host, ksize);
  }
  return rc ? fail(out, rc) : DPS_OK;
}

int dps_operator_create_resize(const int32_t* fov_h, const float* w_h, int taps_h, int out_h, const int32_t* fov_w,
                               const float* w_w, int taps_w, int out_w, int C, int H, int W, dps_operator** out) {
  DPS_REQUIRE(fov_h && w_h && fov_w && w_w && taps_h > 0 && taps_w > 0, DPS_ERR_INVALID,
              "dps_operator_create_resize: null tables");
  if (int rc = new_op(DPS_OP_RESIZE, C, H, W, out, "dps_operator_create_resize")) return rc;
  int rc = resize_create(*out, fov_h, w_h, taps_h, out_h, fov_w, w_w, taps_w, out_w);
  return rc ? fail(out, rc) : DPS_OK;
}

int dps_operator_create_phase(int pad, int C, int H, int W, dps_operator** out) {
  DPS_REQUIRE(pad >= 0, DPS_ERR_INVALID, "dps_operator_create_phase: negative pad");
  if (int rc = new_op(DPS_OP_PHASE, C, H, W, out, "dps_operator_create_phase")) return rc;
  int rc = phase_create(*out, pad);
  return rc ? fail(out, rc) : DPS_OK;
}

void dps_operator_destroy(dps_operator* op) {
  if (!op) return;
  int cur = 0;
  cudaGetDevice(&cur);
  if (cur != op->device) cudaSetDevice(op->device);
  cudaFree(op->mask_dev);
  sep_destroy(op);
  sep_fused_destroy(op);
  sparse_destroy(op);
  resize_destroy(op);
  resize_fused_destroy(op);
  phase_destroy(op);
  if (cur != op->device) cudaSetDevice(cur);
  delete op;
}

int dps_operator_get_info(const dps_operator* op, dps_operator_info* info) {
  DPS_REQUIRE(op && info, DPS_ERR_INVALID, "dps_operator_get_info: null argument");
  info->kind = op->kind;
  info->C = op->C; info->H = op->H; info->W = op->W;
  info->out_C = op->oC; info->out_H = op->oH; info->out_W = op->oW;
  info->partials_per_particle = op->P;
  info->aux_floats_per_particle = op->aux_floats;
  info->taps = op->taps;
  info->guidance_partials = op->guidance_P;
  return DPS_OK;
}

int dps_operator_forward(const dps_operator* op, const dps_source* src, const float* y, int64_t y_stride, float* out,
                         float* partials, float* aux, int n, dps_stream_t stream) {
  DPS_REQUIRE(op && src && src->x, DPS_ERR_INVALID, "dps_operator_forward: null operator/source");
  DPS_REQUIRE(n > 0 && n <= 65535, DPS_ERR_INVALID, "dps_operator_forward: bad particle count %d", n);
  if (int rc = check_device(op, "dps_operator_forward")) return rc;
  DPS_REQUIRE(out || (op->kind == DPS_OP_PHASE && aux), DPS_ERR_INVALID, "dps_operator_forward: null output");
  DPS_REQUIRE(dps_aligned16(src->x) && dps_aligned16(src->eps) && dps_aligned16(y) && dps_aligned16(out) &&
                  dps_aligned16(aux) && src->x_stride % 4 == 0 && (!src->eps || src->eps_stride % 4 == 0) &&
                  y_stride % 4 == 0,
              DPS_ERR_ALIGN, "dps_operator_forward: tensors must be 16-byte aligned, strides multiples of 4");
  DPS_REQUIRE(op->aux_floats == 0 || aux || op->kind == DPS_OP_PHASE || op->kind == DPS_OP_BLUR_SPARSE, DPS_ERR_INVALID,
              "dps_operator_forward: this operator needs the aux workspace");  // (sparse blur: the workspace is the adjoint's)
  FwdArgs a;
  a.src = *src;
  a.y = y;
  a.y_stride = y_stride;
  a.out = out;
  a.partials = partials;
  a.aux = aux;
  a.n = n;
  cudaStream_t st = (cudaStream_t)stream;
  switch (op->kind) {
    case DPS_OP_INPAINT: return inpaint_forward(op, a, st);
    case DPS_OP_BLUR_SEPARABLE: return sep_forward(op, a, st);
    case DPS_OP_BLUR_SPARSE: return sparse_forward(op, a, st);
    case DPS_OP_RESIZE: return resize_forward(op, a, st);
    case DPS_OP_PHASE: return phase_forward(op, a, st);
  }
  dps_set_error("dps_operator_forward: unknown operator kind %d", op->kind);
  return DPS_ERR_INVALID;
}

int dps_operator_guidance(const dps_operator* op, const dps_source* src, const float* y, int64_t y_stride, float* r_out,
                          float* g, int64_t g_stride, float* partials, float* aux, int n, dps_stream_t stream) {
  DPS_REQUIRE(op && src && src->x && g, DPS_ERR_INVALID, "dps_operator_guidance: null operator/source/output");
  DPS_REQUIRE(n > 0 && n <= 65535, DPS_ERR_INVALID, "dps_operator_guidance: bad particle count %d", n);
  if (int rc = check_device(op, "dps_operator_guidance")) return rc;
  if (op->guidance_P > 0) {
    DPS_REQUIRE(dps_aligned16(src->x) && dps_aligned16(src->eps) && dps_aligned16(y) && dps_aligned16(g) && dps_aligned16(r_out) &&
                    src->x_stride % 4 == 0 && (!src->eps || src->eps_stride % 4 == 0) && y_stride % 4 == 0 && g_stride % 4 == 0,
                DPS_ERR_ALIGN, "dps_operator_guidance: tensors must be 16-byte aligned, strides multiples of 4");
    if (op->kind == DPS_OP_RESIZE) return resize_fused_guidance(op, *src, y, y_stride, r_out, g, g_stride, partials, n, (cudaStream_t)stream);
    if (op->kind == DPS_OP_BLUR_SEPARABLE && op->sepfused && src->eps)
      return sep_fused_guidance(op, *src, y, y_stride, r_out, g, g_stride, partials, n, (cudaStream_t)stream);
    if (op->kind == DPS_OP_INPAINT && src->eps && y) {
      FwdArgs fa;
      fa.src = *src;
      fa.y = y;
      fa.y_stride = y_stride;
      fa.out = r_out;
      fa.partials = partials;
      fa.aux = nullptr;
      fa.n = n;
      return inpaint_guidance(op, fa, g, g_stride, (cudaStream_t)stream);
    }
    if (op->kind == DPS_OP_PHASE && src->eps)
      return phase_guidance(op, *src, y, y_stride, r_out, g, g_stride, partials, aux, n, (cudaStream_t)stream);
  }
  // no fused kernel for this operator / shape: the residual kernel, then the adjoint kernel without a coefficient
  DPS_REQUIRE(r_out, DPS_ERR_INVALID, "dps_operator_guidance: this operator has no fused kernel and needs the residual buffer r_out");
  if (int rc = dps_operator_forward(op, src, y, y_stride, r_out, partials, aux, n, stream)) return rc;
  return dps_operator_adjoint(op, r_out, nullptr, src->eps ? src : nullptr, nullptr, 0, g, g_stride, aux, n, stream);
}

int dps_operator_adjoint(const dps_operator* op, const float* r, const float* coef, const dps_source* mask_src,
                         const float* extra, int64_t extra_stride, float* g, int64_t g_stride, const float* aux, int n,
                         dps_stream_t stream) {
  DPS_REQUIRE(op && g, DPS_ERR_INVALID, "dps_operator_adjoint: null operator/output");
  DPS_REQUIRE(r || (op->kind == DPS_OP_PHASE && aux), DPS_ERR_INVALID, "dps_operator_adjoint: null residual");
  DPS_REQUIRE(n > 0 && n <= 65535, DPS_ERR_INVALID, "dps_operator_adjoint: bad particle count %d", n);
  if (int rc = check_device(op, "dps_operator_adjoint")) return rc;
  DPS_REQUIRE(dps_aligned16(r) && dps_aligned16(extra) && dps_aligned16(g) && dps_aligned16(aux) && g_stride % 4 == 0 &&
                  extra_stride % 4 == 0,
              DPS_ERR_ALIGN, "dps_operator_adjoint: tensors must be 16-byte aligned, strides multiples of 4");
  AdjArgs a;
  a.r = r;
  a.coef = coef;
  a.has_mask = mask_src != nullptr;
  if (mask_src) {
    a.mask_src = *mask_src;
    DPS_REQUIRE(dps_aligned16(mask_src->x) && dps_aligned16(mask_src->eps) && mask_src->x_stride % 4 == 0 &&
                    (!mask_src->eps || mask_src->eps_stride % 4 == 0),
                DPS_ERR_ALIGN, "dps_operator_adjoint: mask source must be 16-byte aligned");
  } else {
    a.mask_src = dps_source{};
  }
  a.extra = extra;
  a.extra_stride = extra_stride;
  a.g = g;
  a.g_stride = g_stride;
  a.aux = aux;
  a.n = n;
  cudaStream_t st = (cudaStream_t)stream;
  switch (op->kind) {
    case DPS_OP_INPAINT: return inpaint_adjoint(op, a, st);
    case DPS_OP_BLUR_SEPARABLE: return sep_adjoint(op, a, st);
    case DPS_OP_BLUR_SPARSE: return sparse_adjoint(op, a, st);
    case DPS_OP_RESIZE: return resize_adjoint(op, a, st);
    case DPS_OP_PHASE: return phase_adjoint(op, a, st);
  }
  dps_set_error("dps_operator_adjoint: unknown operator kind %d", op->kind);
  return DPS_ERR_INVALID;
}

}  // extern "C"
