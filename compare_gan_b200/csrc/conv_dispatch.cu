// Public conv entry points: choose between the exact-fp32 gather-GEMM (gemm.cu) and the tcgen05
// tensor-core implicit GEMM (conv_tc.cu) according to the context's math mode and the shape.
#include "common.cuh"

int cgan_conv2d_fwd(cgan_ctx* ctx, const cgan_conv_desc* d, const float* x, const float* w, const float* bias, float* y) {
  if (!ctx) return CGAN_ERR_ARG;
  CGAN_REQUIRE(ctx, d && x && w && y, "null pointer");
  return cgan_conv2d_fwd_simt(ctx, d, x, w, bias, y);
}

int cgan_conv2d_dgrad(cgan_ctx* ctx, const cgan_conv_desc* d, const float* dy, const float* w, float* dx) {
  if (!ctx) return CGAN_ERR_ARG;
  CGAN_REQUIRE(ctx, d && dy && w && dx, "null pointer");
  return cgan_conv2d_dgrad_simt(ctx, d, dy, w, dx);
}
