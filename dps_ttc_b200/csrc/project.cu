// Operator.project / ortho_project behind one C entry point, dps_operator_project.
//
// The reference composes them from forward / transpose and elementwise ops (measurements.py:48-54), with a "transpose"
// that is the identity for the blur and inpainting operators (:70-71, :123-124, :146-147) and nearest-neighbour
// up-sampling for super-resolution (:80-88):
//   LinearOperator.ortho_project(d)      = d − T(A d)
//   LinearOperator.project(d, y)         = ortho_project(y) − A d        = (y − A y) − A d          (T = identity)
//   SuperResolutionOperator.project(d,y) = d − up(A d) + up(y)                                       (:90-91)
//   NonLinearOperator.project(d, y)      = d + y − A(d)                                               (:175-177)
// Here the elementwise part rides in the epilogue of the operator kernels:
//   identity transpose — the forward kernels already evaluate `y − A(src)` in their epilogue (the DPS residual), so
//     ortho_project(d) is ONE forward launch with the data as its own "measurement", and project(d, y) is two:
//     t = y − A y (measurement-sized batch, usually 1 plane set), out = t − A d.  Same roundings as the reference
//     (fp32 subtractions in the same order).
//   super-resolution ×4/×8 at 256² — one launch of the cluster kernel of resize_fused.cu in PROJ mode (A·d and y stay in
//     shared memory, the epilogue writes (d − up(A d)) + up(y)); other shapes: forward launch + the combine kernel below.
//   phase retrieval — measurement and image shapes differ: the reference's expression cannot be evaluated either
//     (oracle/make_golden.py gen_project) → DPS_ERR_UNSUPPORTED.
#include "operator.cuh"

int resize_fused_project(const dps_operator* op, const float* data, int64_t data_stride, const float* y, int64_t y_stride, float* out,
                         int64_t out_stride, int n, cudaStream_t st);

namespace {

// out = (d − up(a)) + up(y):  a = A·d (n, C, oH, oW) dense, y nullable (stride 0 = broadcast), up = nearest ×F
__global__ void __launch_bounds__(256) upsample_combine_kernel(const float* __restrict__ d, int64_t d_stride, const float* __restrict__ a,
                                                               const float* __restrict__ y, int64_t y_stride, float* __restrict__ out,
                                                               int64_t out_stride, int C, int H, int W, int oH, int oW) {
  const int n = blockIdx.y;
  const int64_t i4 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;  // float4 index inside the particle
  const int64_t chw4 = (int64_t)C * H * W / 4;
  if (i4 >= chw4) return;
  const int Fh = H / oH, Fw = W / oW;
  const int64_t e = i4 * 4;
  const int col = (int)(e % W), row = (int)((e / W) % H), c = (int)(e / ((int64_t)W * H));
  const float4 dv = ldg_stream4(d + n * d_stride + e);
  const float dd[4] = {dv.x, dv.y, dv.z, dv.w};
  float r[4];
  const int64_t obase = ((int64_t)c * oH + row / Fh) * oW;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int64_t o = obase + (col + k) / Fw;
    const float av = __ldg(a + (int64_t)n * C * oH * oW + o);
    const float yv = y ? __ldg(y + n * y_stride + o) : 0.f;
    r[k] = __fadd_rn(__fsub_rn(dd[k], av), yv);
  }
  stg_stream4(out + n * out_stride + e, make_float4(r[0], r[1], r[2], r[3]));
}

}  // namespace

extern "C" int dps_operator_project(const dps_operator* op, const float* data, int64_t data_stride, const float* y, int64_t y_stride,
                                    int n_y, float* out, int64_t out_stride, float* scratch, int n, dps_stream_t stream) {
  DPS_REQUIRE(op && data && out, DPS_ERR_INVALID, "dps_operator_project: null operator/data/output");
  DPS_REQUIRE(n > 0 && n <= 65535, DPS_ERR_INVALID, "dps_operator_project: bad particle count %d", n);
  DPS_REQUIRE(!y || n_y == 1 || n_y == n, DPS_ERR_INVALID, "dps_operator_project: measurement batch %d must be 1 or %d", n_y, n);
  DPS_REQUIRE(dps_aligned16(data) && dps_aligned16(y) && dps_aligned16(out) && dps_aligned16(scratch) && data_stride % 4 == 0 &&
                  y_stride % 4 == 0 && out_stride % 4 == 0,
              DPS_ERR_ALIGN, "dps_operator_project: tensors must be 16-byte aligned, strides multiples of 4");
  int cur = -1;
  DPS_CUDA(cudaGetDevice(&cur));
  DPS_REQUIRE(cur == op->device, DPS_ERR_INVALID, "dps_operator_project: the operator was created on device %d but the current device is %d",
              op->device, cur);
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t chw = (int64_t)op->C * op->H * op->W, m = (int64_t)op->oC * op->oH * op->oW;
  dps_source src = {};
  src.x = data;
  src.x_stride = data_stride;
  switch (op->kind) {
    case DPS_OP_INPAINT:
    case DPS_OP_BLUR_SEPARABLE:
    case DPS_OP_BLUR_SPARSE: {
      DPS_REQUIRE(out_stride == chw, DPS_ERR_INVALID, "dps_operator_project: the output must be dense for this operator");
      float* aux = nullptr;  // (the sparse blur's workspace belongs to its adjoint; the forward kernels need none)
      if (!y)  // ortho_project(d) = d − A d: the residual epilogue with the data as its own measurement
        return dps_operator_forward(op, &src, data, data_stride, out, nullptr, aux, n, stream);
      DPS_REQUIRE(scratch, DPS_ERR_INVALID, "dps_operator_project: project() needs a scratch buffer of n_y measurement-sized planes");
      dps_source ys = {};
      ys.x = y;
      ys.x_stride = y_stride ? y_stride : m;
      if (int rc = dps_operator_forward(op, &ys, y, ys.x_stride, scratch, nullptr, aux, n_y, stream)) return rc;  // t = y − A y
      return dps_operator_forward(op, &src, scratch, n_y == 1 ? 0 : m, out, nullptr, aux, n, stream);             // t − A d
    }
    case DPS_OP_RESIZE: {
      if (op->rfused) return resize_fused_project(op, data, data_stride, y, n_y == 1 ? 0 : y_stride, out, out_stride, n, st);
      DPS_REQUIRE(scratch, DPS_ERR_INVALID, "dps_operator_project: this shape has no fused kernel and needs a scratch buffer of n measurement planes");
      DPS_REQUIRE(op->H % op->oH == 0 && op->W % op->oW == 0 && op->W % 4 == 0, DPS_ERR_UNSUPPORTED,
                  "dps_operator_project: nearest up-sampling needs an integer factor");
      if (int rc = dps_operator_forward(op, &src, nullptr, 0, scratch, nullptr, nullptr, n, stream)) return rc;   // a = A d
      dim3 grid((unsigned)((chw / 4 + 255) / 256), (unsigned)n);
      upsample_combine_kernel<<<grid, 256, 0, st>>>(data, data_stride, scratch, y, n_y == 1 ? 0 : y_stride, out, out_stride, op->C, op->H,
                                                    op->W, op->oH, op->oW);
      DPS_LAUNCH_CHECK("upsample_combine");
      return DPS_OK;
    }
    default:
      dps_set_error("dps_operator_project: measurement and image shapes differ for this operator (the reference's expression "
                    "data + measurement - A(data) cannot be evaluated either)");
      return DPS_ERR_UNSUPPORTED;
  }
}
