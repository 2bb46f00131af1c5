// Op-level C-ABI entries of the weight-gradient kernels (parity tests of single kernels; the training step in
// train.cu launches the same kernels with prebuilt tensor maps).
#include "c2w_b200.h"
#include "common.cuh"
#include "wgrad_tcgen05.cuh"

using namespace c2w;

int c2w_num_sms();  // op_conv.cu

extern "C" {

int c2w_op_wgrad(const void* x, const void* dy, int32_t n_img, int32_t H, int32_t W, int32_t cin_pad, int32_t cout_pad,
                 int32_t stride, int32_t conv3x3, float* scratch, int64_t scratch_floats, float* dw, int32_t cin,
                 int32_t cout, int32_t accumulate, void* stream) {
  C2W_REQUIRE(x && dy && dw && scratch && n_img >= 1 && cin >= 1 && cout >= 1 && cin <= cin_pad && cout <= cout_pad,
              "c2w_op_wgrad: bad argument");
  const int sms = c2w_num_sms();
  C2W_REQUIRE(sms > 0, "c2w_op_wgrad: no CUDA device");
  WgradLaunch L;
  if (!wgrad_launch_init(&L, conv3x3 != 0, static_cast<const __nv_bfloat16*>(x), static_cast<const __nv_bfloat16*>(dy), n_img,
                         H, W, cin_pad, cout_pad, stride ? stride : 1, sms, scratch, static_cast<size_t>(scratch_floats)))
    return fail(C2W_ERR_INVALID, "c2w_op_wgrad: cannot build launch (n=%d H=%d W=%d cin=%d cout=%d stride=%d; channel "
                "counts must be multiples of 64, output images multiples of 8 x 8, GEMM rows a multiple of 64; scratch "
                "%lld floats) %s", n_img, H, W, cin_pad, cout_pad, stride, (long long)scratch_floats, tmap_error_slot());
  C2W_CUDA(wgrad_run(L, dw, cout, cin, accumulate, 1.0f, static_cast<cudaStream_t>(stream)));
  return C2W_OK;
}

int c2w_op_colsum(const void* x_bf16, float* out, int64_t rows, int32_t C, int64_t rows_per_group, int32_t out_stride,
                  float scale, void* stream) {
  C2W_REQUIRE(x_bf16 && out && rows >= 1 && C >= 2 && C % 2 == 0 && C <= 512 && rows_per_group >= 1,
              "c2w_op_colsum: bad argument");
  C2W_CUDA(colsum_run(static_cast<const __nv_bfloat16*>(x_bf16), out, rows, C, rows_per_group, out_stride, scale,
                      static_cast<cudaStream_t>(stream)));
  return C2W_OK;
}

}  // extern "C"
