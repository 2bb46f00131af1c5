// EXPERIMENTAL op-level entry points of the weight-gradient GEMM (see wgrad_tcgen05.cuh: not working yet).
#include "c2w_b200.h"
#include "common.cuh"
#include "wgrad_tcgen05.cuh"

using namespace c2w;

int c2w_num_sms();  // op_conv.cu

extern "C" {

int c2w_op_transpose_bf16(const void* in, void* out, int64_t rows, int32_t cols, void* stream) {
  C2W_REQUIRE(in && out && rows >= 1 && cols >= 1, "c2w_op_transpose_bf16: bad argument");
  dim3 grid(static_cast<unsigned>((rows + 63) / 64), static_cast<unsigned>((cols + 63) / 64));
  transpose_bf16_kernel<<<grid, dim3(32, 8), 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __nv_bfloat16*>(in), static_cast<__nv_bfloat16*>(out), rows, cols);
  C2W_CUDA(cudaGetLastError());
  return C2W_OK;
}

int c2w_op_wgrad(const void* x_t, const void* dy_t, int32_t n_img, int32_t H, int32_t W, int32_t cin, int32_t cout,
                 float* dw, void* stream) {
  C2W_REQUIRE(x_t && dy_t && dw && n_img >= 1, "c2w_op_wgrad: bad argument");
  const int sms = c2w_num_sms();
  C2W_REQUIRE(sms > 0, "c2w_op_wgrad: no CUDA device");
  char msg[300];
  if (wgrad_launch(static_cast<const __nv_bfloat16*>(x_t), static_cast<const __nv_bfloat16*>(dy_t), n_img, H, W, cin, cout,
                   dw, sms, static_cast<cudaStream_t>(stream), msg, sizeof(msg)) != 0)
    return fail(C2W_ERR_INVALID, "c2w_op_wgrad: %s", msg);
  return C2W_OK;
}

}  // extern "C"
