// Op-level C-ABI entry for K1 (used by the parity tests and by tools/; the engine calls
// conv_launch_init/conv_launch directly with prebuilt tensor maps).
#include "c2w_b200.h"
#include "common.cuh"
#include "conv_tcgen05.cuh"

using namespace c2w;

static int g_num_sms = 0;

int c2w_num_sms() {
  if (g_num_sms == 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 0;
    cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev);
  }
  return g_num_sms;
}

extern "C" {

// x: bf16 NHWC [n_img, H, W, cin]  (conv3x3 != 0)   or   bf16 [n_img*H*W, cin] row-major (conv3x3 == 0)
// w_packed: bf16 [cout_pad, taps*cin] with k = (r*3+s)*cin + c ; bias: fp32 [cout_pad]
// mode: EpiMode (0 bias, 1 bias+silu, 2 bias+residual, 4 fp32 out); res/out: bf16 [M, cout_pad]
// bn: N tile (0 = pick) ; max_ctas: 0 = one per SM
int c2w_op_conv(const void* x, int n_img, int H, int W, int cin, const void* w_packed, int cout_pad,
                const float* bias, int mode, const void* res, void* out, float* out_f32, int conv3x3, int bn,
                int max_ctas, void* stream) {
  C2W_REQUIRE(mode == EPI_BIAS || mode == EPI_BIAS_SILU || mode == EPI_BIAS_RES || mode == EPI_F32,
              "c2w_op_conv: unsupported mode %d", mode);
  const int sms = c2w_num_sms();
  C2W_REQUIRE(sms > 0, "c2w_op_conv: no CUDA device");
  const int dbg = bn >> 12;  // diagnostics: bit 12 of `bn` = skip TMA loads once the ring is primed
  bn &= 0xfff;
  if (bn == 0) bn = conv_pick_bn(cout_pad);
  ConvLaunch L;
  if (!conv_launch_init(&L, conv3x3 != 0, static_cast<const __nv_bfloat16*>(x), n_img, H, W, cin,
                        static_cast<const __nv_bfloat16*>(w_packed), cout_pad, bn, max_ctas > 0 ? max_ctas : sms))
    return fail(C2W_ERR_INVALID, "c2w_op_conv: cannot build launch (n=%d H=%d W=%d cin=%d cout=%d bn=%d)", n_img, H,
                W, cin, cout_pad, bn);
  L.p.mode = mode;
  L.p.bias = bias;
  L.p.res = static_cast<const __nv_bfloat16*>(res);
  L.p.out = static_cast<__nv_bfloat16*>(out);
  L.p.out_f32 = out_f32;
  L.p.dbg_skip_loads = dbg & 1;
  C2W_CUDA(conv_launch(L, static_cast<cudaStream_t>(stream)));
  return C2W_OK;
}

}  // extern "C"
