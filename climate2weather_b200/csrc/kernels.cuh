// Fused, coalesced, vectorised CUDA kernels around K1 (all HBM- or latency-bound; no tensor cores):
//   K0  gather_windows      src/thor/score.py:68-74,143-154   unfold -> bf16 NHWC window batch
//   K2  channel_layernorm   model/nn.py:154,183,44 (zuko LayerNorm over C) + modulation add (+ 2x nearest upsample)
//   K3  time embedding MLP + all modulation projections        model/score.py:14-34,61-67 ; model/nn.py:149
//   K4  attention core      model/nn.py:64-85
//   K6  guided eps + predictor update                          src/thor/score.py:44-60,24-35 ; pipelines.py:41-46
//   K7  corrector (guided eps, ||eps||^2, Langevin update)     src/thor/pipelines.py:81-88
//   layout converters between the reference's NCHW fp32 tensors and the device layouts
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace c2w {

typedef __nv_bfloat16 bf16;

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ uint32_t pack_bf16x2(float a, float b) {
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ float2 unpack_bf16x2(uint32_t w) {
  __nv_bfloat162 h = *reinterpret_cast<__nv_bfloat162*>(&w);
  return make_float2(__low2float(h), __high2float(h));
}

// ------------------------------------------------------------------------------------------------ K0
// traj: fp32 [frames, HW, C]; window i of this launch starts at local frame f0 + i.
// out : bf16 [n, HW, cin_pad], channel tau*C + c  <-  traj[f0 + i + tau, pix, c]; channels >= wC are zero.
// One thread writes 8 channels (16 B): writes are fully coalesced, reads are 16 B (C = 4) sectors.
__global__ void gather_windows_kernel(const float* __restrict__ traj, bf16* __restrict__ out, int n, int hw, int C,
                                      int wC, int cin_pad, int f0) {
  const int groups = cin_pad >> 3;
  const long long total = static_cast<long long>(n) * hw * groups;
  for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int g = static_cast<int>(idx % groups);
    const long long pw = idx / groups;
    const int pix = static_cast<int>(pw % hw);
    const int i = static_cast<int>(pw / hw);
    float v[8];
    if (C == 4) {
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int tau = 2 * g + h;
        float4 q = make_float4(0.f, 0.f, 0.f, 0.f);
        if (tau * 4 < wC)
          q = __ldg(reinterpret_cast<const float4*>(traj + (static_cast<long long>(f0 + i + tau) * hw + pix) * 4));
        v[4 * h + 0] = q.x;
        v[4 * h + 1] = q.y;
        v[4 * h + 2] = q.z;
        v[4 * h + 3] = q.w;
      }
    } else {
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const int ch = 8 * g + e;
        const int tau = ch / C, c = ch - tau * C;
        v[e] = (ch < wC) ? __ldg(traj + (static_cast<long long>(f0 + i + tau) * hw + pix) * C + c) : 0.f;
      }
    }
    uint4 o = make_uint4(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]), pack_bf16x2(v[4], v[5]),
                         pack_bf16x2(v[6], v[7]));
    *reinterpret_cast<uint4*>(out + (static_cast<long long>(i) * hw + pix) * cin_pad + 8 * g) = o;
  }
}

// ------------------------------------------------------------------------------------------------ K2
// y = (v - mean_C(v)) / sqrt(var_C(v) + eps), v = x + mod, unbiased variance (torch.var_mean default), no affine.
// One warp per pixel; lane owns NCH chunks of VEC channels: chunk j = channels [j*32*VEC + lane*VEC, +VEC)
// so every chunk is one coalesced warp access.  upsample != 0 writes each pixel to its 2x2 nearest-neighbour
// block of a [n, 2H, 2W, C] tensor (model/nn.py:183-184 fused).
template <int C>
__global__ void channel_layernorm_kernel(const bf16* __restrict__ x, const float* __restrict__ mod,
                                         bf16* __restrict__ out, long long npix, int H, int W, int upsample,
                                         float eps) {
  constexpr int VEC = (C % 128 == 0) ? 4 : 2;
  constexpr int NCH = C / (32 * VEC);
  static_assert(C % 64 == 0 && NCH >= 1, "C must be a multiple of 64");
  const int lane = threadIdx.x & 31;
  const long long warp0 = (blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x) >> 5;
  const long long nwarps = (static_cast<long long>(gridDim.x) * blockDim.x) >> 5;
  float m[NCH * VEC];
#pragma unroll
  for (int j = 0; j < NCH; ++j)
#pragma unroll
    for (int e = 0; e < VEC; ++e) m[j * VEC + e] = mod ? __ldg(mod + j * 32 * VEC + lane * VEC + e) : 0.f;

  for (long long pix = warp0; pix < npix; pix += nwarps) {
    const bf16* px = x + pix * C;
    float v[NCH * VEC];
#pragma unroll
    for (int j = 0; j < NCH; ++j) {
      const int c0 = j * 32 * VEC + lane * VEC;
      if (VEC == 4) {
        const uint2 r = *reinterpret_cast<const uint2*>(px + c0);
        const float2 a = unpack_bf16x2(r.x), b = unpack_bf16x2(r.y);
        v[j * 4 + 0] = a.x + m[j * 4 + 0];
        v[j * 4 + 1] = a.y + m[j * 4 + 1];
        v[j * 4 + 2] = b.x + m[j * 4 + 2];
        v[j * 4 + 3] = b.y + m[j * 4 + 3];
      } else {
        const float2 a = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(px + c0));
        v[j * 2 + 0] = a.x + m[j * 2 + 0];
        v[j * 2 + 1] = a.y + m[j * 2 + 1];
      }
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < NCH * VEC; ++i) s += v[i];
    const float mean = warp_sum(s) * (1.0f / C);
    float ss = 0.f;
#pragma unroll
    for (int i = 0; i < NCH * VEC; ++i) {
      v[i] -= mean;
      ss += v[i] * v[i];
    }
    const float var = warp_sum(ss) * (1.0f / (C - 1));
    const float inv = 1.0f / sqrtf(var + eps);
    long long obase[4];
    int nout = 1;
    if (upsample) {
      const int w = static_cast<int>(pix % W);
      const long long t = pix / W;
      const int h = static_cast<int>(t % H);
      const long long n = t / H;
      const long long o00 = ((n * 2 * H + 2 * h) * 2 * W + 2 * w);
      obase[0] = o00;
      obase[1] = o00 + 1;
      obase[2] = o00 + 2 * W;
      obase[3] = o00 + 2 * W + 1;
      nout = 4;
    } else {
      obase[0] = pix;
    }
#pragma unroll
    for (int j = 0; j < NCH; ++j) {
      const int c0 = j * 32 * VEC + lane * VEC;
      if (VEC == 4) {
        const uint2 o = make_uint2(pack_bf16x2(v[j * 4] * inv, v[j * 4 + 1] * inv),
                                   pack_bf16x2(v[j * 4 + 2] * inv, v[j * 4 + 3] * inv));
#pragma unroll
        for (int q = 0; q < 4; ++q)
          if (q < nout) *reinterpret_cast<uint2*>(out + obase[q] * C + c0) = o;
      } else {
        const uint32_t o = pack_bf16x2(v[j * 2] * inv, v[j * 2 + 1] * inv);
#pragma unroll
        for (int q = 0; q < 4; ++q)
          if (q < nout) *reinterpret_cast<uint32_t*>(out + obase[q] * C + c0) = o;
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------ K3
// h0 = silu(W0 * [cos(t f), sin(t f)] + b0)      (model/score.py:14-34, :62-63)
__global__ void time_embed_kernel(float t, const float* __restrict__ W0, const float* __restrict__ b0,
                                  float* __restrict__ h0, int E, int nf) {
  extern __shared__ float e[];
  const int half = nf / 2;
  if (threadIdx.x < half) {
    const float f = expf(-9.210340371976184f * static_cast<float>(threadIdx.x) / static_cast<float>(half));
    const float a = t * f;
    e[threadIdx.x] = cosf(a);
    e[half + threadIdx.x] = sinf(a);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < E; i += blockDim.x) {
    float acc = b0[i];
    for (int j = 0; j < nf; ++j) acc += W0[i * nf + j] * e[j];
    h0[i] = acc / (1.0f + expf(-acc));
  }
}
// y[r] = act(b[r] + W[r, :] . x)   warp per row, float4 lanes.  act: 0 none, 1 SiLU
__global__ void matvec_kernel(const float* __restrict__ W, const float* __restrict__ b, const float* __restrict__ x,
                              float* __restrict__ y, int rows, int cols, int act) {
  const int lane = threadIdx.x & 31;
  const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (row >= rows) return;
  const float4* w4 = reinterpret_cast<const float4*>(W + static_cast<size_t>(row) * cols);
  const float4* x4 = reinterpret_cast<const float4*>(x);
  float acc = 0.f;
  for (int j = lane; j < cols / 4; j += 32) {
    const float4 a = __ldg(w4 + j), v = x4[j];
    acc += a.x * v.x + a.y * v.y + a.z * v.z + a.w * v.w;
  }
  acc = warp_sum(acc);
  if (lane == 0) {
    acc += b[row];
    y[row] = act ? acc / (1.0f + expf(-acc)) : acc;
  }
}

// ------------------------------------------------------------------------------------------------ K4
// qkv: bf16 [n*T, 3C] (q | k | v along channels, model/nn.py:74), out: bf16 [n*T, C].
// CTA = (query block of QB, window).  S^T = (K Q^T) / sqrt(C) in fp32 smem, fp32 softmax over keys, O = P V.
// QB = 64 when there are enough windows to fill the SMs, 16 otherwise (4x the CTAs; K and V are re-read from L2).
constexpr int kAttnThreads = 256;
inline size_t attention_smem_bytes(int T, int C, int QB = 64) {
  const size_t pitch = (C + 2) * 2;
  return (QB + 2 * static_cast<size_t>(T)) * pitch + static_cast<size_t>(T) * (QB + 4) * 4;
}
template <int QB>
__global__ void __launch_bounds__(kAttnThreads)
attention_kernel(const bf16* __restrict__ qkv, bf16* __restrict__ out, int T, int C, float scale2) {
  extern __shared__ __align__(16) uint8_t smraw[];
  const int pitch = C + 2;  // bf16 elements; +2 shifts consecutive rows by one bank
  bf16* sq = reinterpret_cast<bf16*>(smraw);
  bf16* sk = sq + QB * pitch;
  bf16* sv = sk + static_cast<size_t>(T) * pitch;
  const int sp = QB + 4;  // pitch of S^T rows (floats), keeps float4 alignment
  float* st = reinterpret_cast<float*>(sv + static_cast<size_t>(T) * pitch);
  const int q0 = blockIdx.x * QB;
  const int nq = min(QB, T - q0);
  const bf16* base = qkv + static_cast<size_t>(blockIdx.y) * T * 3 * C;
  const int tid = threadIdx.x;
  const int c8 = C / 8;
  // ---- load q block, k, v (16 B global reads, 4 B smem writes because of the padded pitch)
  for (int idx = tid; idx < (QB + 2 * T) * c8; idx += kAttnThreads) {
    const int row = idx / c8, g = idx - row * c8;
    const bf16* src;
    bf16* dst;
    bool ok = true;
    if (row < QB) {
      ok = row < nq;
      src = base + static_cast<size_t>(q0 + row) * 3 * C + g * 8;
      dst = sq + row * pitch + g * 8;
    } else if (row < QB + T) {
      const int r = row - QB;
      src = base + static_cast<size_t>(r) * 3 * C + C + g * 8;
      dst = sk + static_cast<size_t>(r) * pitch + g * 8;
    } else {
      const int r = row - QB - T;
      src = base + static_cast<size_t>(r) * 3 * C + 2 * C + g * 8;
      dst = sv + static_cast<size_t>(r) * pitch + g * 8;
    }
    uint4 v = make_uint4(0, 0, 0, 0);
    if (ok) v = *reinterpret_cast<const uint4*>(src);
    uint32_t* d = reinterpret_cast<uint32_t*>(dst);
    d[0] = v.x;
    d[1] = v.y;
    d[2] = v.z;
    d[3] = v.w;
  }
  __syncthreads();
  // ---- S^T[s][tq] = scale2 * q[tq] . k[s]   (MQ x 4 register micro-tiles; 16 query groups x T/4 key groups)
  constexpr int MQ = QB / 16;
  const int tiles = 16 * (T / 4);
  for (int mt = tid; mt < tiles; mt += kAttnThreads) {
    const int tqg = mt % 16, s4 = mt / 16;
    float acc[MQ][4];
#pragma unroll
    for (int i = 0; i < MQ; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    const uint32_t* qp = reinterpret_cast<const uint32_t*>(sq + (tqg * MQ) * pitch);
    const uint32_t* kp = reinterpret_cast<const uint32_t*>(sk + static_cast<size_t>(s4 * 4) * pitch);
    const int pw = pitch / 2;
    for (int c = 0; c < C / 2; ++c) {
      float2 qv[MQ], kv[4];
#pragma unroll
      for (int i = 0; i < MQ; ++i) qv[i] = unpack_bf16x2(qp[i * pw + c]);
#pragma unroll
      for (int j = 0; j < 4; ++j) kv[j] = unpack_bf16x2(kp[j * pw + c]);
#pragma unroll
      for (int i = 0; i < MQ; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] += qv[i].x * kv[j].x + qv[i].y * kv[j].y;
    }
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
      for (int i = 0; i < MQ; ++i) st[(s4 * 4 + j) * sp + tqg * MQ + i] = acc[i][j] * scale2;
  }
  __syncthreads();
  // ---- softmax over keys s for each query column tq (fp32, model/nn.py:82)
  {
    const int warp = tid >> 5, lane = tid & 31;
    for (int tq = warp; tq < QB; tq += kAttnThreads / 32) {
      float mx = -INFINITY;
      for (int s = lane; s < T; s += 32) mx = fmaxf(mx, st[s * sp + tq]);
      mx = warp_max(mx);
      float sum = 0.f;
      for (int s = lane; s < T; s += 32) {
        const float e = expf(st[s * sp + tq] - mx);
        st[s * sp + tq] = e;
        sum += e;
      }
      sum = warp_sum(sum);
      const float inv = 1.0f / sum;
      for (int s = lane; s < T; s += 32) st[s * sp + tq] *= inv;
    }
  }
  __syncthreads();
  // ---- O[tq][c] = sum_s P[tq][s] v[s][c]; thread = (channel pair, 8 queries)
  const int pairs = C / 2;
  for (int item = tid; item < pairs * (QB / 8); item += kAttnThreads) {
    const int c2 = item % pairs, qb = item / pairs;
    float a0[8], a1[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) a0[i] = a1[i] = 0.f;
    const uint32_t* vp = reinterpret_cast<const uint32_t*>(sv) + c2;
    const int pw = pitch / 2;
    for (int s = 0; s < T; ++s) {
      const float2 vv = unpack_bf16x2(vp[static_cast<size_t>(s) * pw]);
      const float4 p0 = *reinterpret_cast<const float4*>(st + s * sp + qb * 8);
      const float4 p1 = *reinterpret_cast<const float4*>(st + s * sp + qb * 8 + 4);
      const float pr[8] = {p0.x, p0.y, p0.z, p0.w, p1.x, p1.y, p1.z, p1.w};
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        a0[i] += pr[i] * vv.x;
        a1[i] += pr[i] * vv.y;
      }
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int tq = qb * 8 + i;
      if (tq < nq)
        *reinterpret_cast<uint32_t*>(out + (static_cast<size_t>(blockIdx.y) * T + q0 + tq) * C + 2 * c2) =
            pack_bf16x2(a0[i], a1[i]);
    }
  }
}

// ------------------------------------------------------------------------------------------------ layout converters
// fp32 NCHW [F, C, HW]  ->  fp32 [F, HW, C]   (trajectory state layout; C = 4 -> one float4 per pixel)
__global__ void nchw_to_fhwc_kernel(const float* __restrict__ in, float* __restrict__ out, long long F, int C, int hw) {
  const long long total = F * hw;
  for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long f = idx / hw;
    const int pix = static_cast<int>(idx - f * hw);
    for (int c = 0; c < C; ++c) out[idx * C + c] = in[(f * C + c) * hw + pix];
  }
}
__global__ void fhwc_to_nchw_kernel(const float* __restrict__ in, float* __restrict__ out, long long F, int C, int hw) {
  const long long total = F * hw;
  for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long f = idx / hw;
    const int pix = static_cast<int>(idx - f * hw);
    for (int c = 0; c < C; ++c) out[(f * C + c) * hw + pix] = in[idx * C + c];
  }
}
// fp32 NCHW [n, C, HW] -> bf16 [n, HW, cpad] (zero padded channels); smem transpose keeps both sides coalesced
__global__ void nchw_to_nhwc_bf16_kernel(const float* __restrict__ in, bf16* __restrict__ out, int C, int hw, int cpad) {
  __shared__ float tile[32][33];
  const long long img = blockIdx.z;
  const int p0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int c = c0 + i, p = p0 + threadIdx.x;
    tile[i][threadIdx.x] = (c < C && p < hw) ? in[(img * C + c) * hw + p] : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int p = p0 + i, c = c0 + threadIdx.x;
    if (p < hw && c < cpad) out[(img * hw + p) * cpad + c] = __float2bfloat16_rn(tile[threadIdx.x][i]);
  }
}
// fp32 [n, HW, cpad] -> fp32 NCHW [n, C, HW]
__global__ void nhwc_f32_to_nchw_kernel(const float* __restrict__ in, float* __restrict__ out, int C, int hw, int cpad) {
  __shared__ float tile[32][33];
  const long long img = blockIdx.z;
  const int p0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int p = p0 + i, c = c0 + threadIdx.x;
    tile[i][threadIdx.x] = (p < hw && c < cpad) ? in[(img * hw + p) * cpad + c] : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int c = c0 + i, p = p0 + threadIdx.x;
    if (c < C && p < hw) out[(img * C + c) * hw + p] = tile[threadIdx.x][i];
  }
}

// ------------------------------------------------------------------------------------------------ K6 / K7
// State layout: x, eps, z are fp32 [frames, H, W, 4] (one float4 per pixel).  Observation y: fp32 [n_obs, 4, Hs, Ws]
// exactly as the reference builds it (exp/downscaling.py:129-132: every t_step-th frame, s x s tile means).
struct GuideParams {
  float* x;
  const float* eps;  // unguided window-composed score (K1 compose epilogue)
  float* eps_out;    // mode 1: guided eps
  const float* y;    // null => unconditioned
  float std2[4];     // likelihood std^2 per variable
  float gamma[4];
  float mu, sigma;            // at the time the score was evaluated
  float mu_next, sigma_next;  // mode 0: target of the predictor step
  int t_step, s_step, H, W;
  int frame_global0;  // global index of local frame 0
  int own_lo;         // first local frame this launch updates (blockIdx.y = 0)
  int mode;           // 0: predictor update of x in place; 1: guided eps -> eps_out + partial sum of squares
  float* partials;    // mode 1: one float per CTA
  int* nan_flag;
};

// CTA = one s-row strip of one frame; one warp per s x s observation tile (warp-shuffle tile mean, deterministic).
__global__ void guided_step_kernel(const GuideParams p) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int s = p.s_step;
  const int fl = p.own_lo + blockIdx.y;
  const int fg = p.frame_global0 + fl;
  const int h0 = blockIdx.x * s, w0 = warp * s;
  const long long fbase = (static_cast<long long>(fl) * p.H) * p.W;
  const bool observed = (p.y != nullptr) && (fg % p.t_step == 0);
  const float inv_mu = 1.0f / p.mu;
  float4 corr = make_float4(0.f, 0.f, 0.f, 0.f);  // eps_guided = eps - corr   (src/thor/score.py:35)
  if (observed) {
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int i = lane; i < s * s; i += 32) {
      const long long o = (fbase + static_cast<long long>(h0 + i / s) * p.W + (w0 + i % s)) * 4;
      const float4 xv = *reinterpret_cast<const float4*>(p.x + o);
      const float4 ev = *reinterpret_cast<const float4*>(p.eps + o);
      acc.x += (xv.x - p.sigma * ev.x) * inv_mu;
      acc.y += (xv.y - p.sigma * ev.y) * inv_mu;
      acc.z += (xv.z - p.sigma * ev.z) * inv_mu;
      acc.w += (xv.w - p.sigma * ev.w) * inv_mu;
    }
    const float inv_area = 1.0f / static_cast<float>(s * s);
    const float mean[4] = {warp_sum(acc.x) * inv_area, warp_sum(acc.y) * inv_area, warp_sum(acc.z) * inv_area,
                           warp_sum(acc.w) * inv_area};
    const int Hs = p.H / s, Ws = p.W / s;
    const int m = fg / p.t_step;
    const float r2 = (p.sigma * inv_mu) * (p.sigma * inv_mu);
    float c[4];
#pragma unroll
    for (int ch = 0; ch < 4; ++ch) {
      const float yv = __ldg(p.y + ((static_cast<long long>(m) * 4 + ch) * Hs + blockIdx.x) * Ws + warp);
      const float err = yv - mean[ch];
      const float var = p.std2[ch] + p.gamma[ch] * r2;
      // J = A^T(err / var) / mu  (each pixel of the tile gets err/var / s^2);  corr = sigma * J
      c[ch] = p.sigma * (err / var) * inv_area * inv_mu;
    }
    corr = make_float4(c[0], c[1], c[2], c[3]);
  }
  float sq = 0.f;
  bool bad = false;
  for (int i = lane; i < s * s; i += 32) {
    const long long o = (fbase + static_cast<long long>(h0 + i / s) * p.W + (w0 + i % s)) * 4;
    float4 ev = *reinterpret_cast<const float4*>(p.eps + o);
    ev.x -= corr.x;
    ev.y -= corr.y;
    ev.z -= corr.z;
    ev.w -= corr.w;
    if (p.mode == 0) {
      float4 xv = *reinterpret_cast<const float4*>(p.x + o);
      // x0 = (x - sigma eps)/mu ; x <- mu' x0 + sigma' eps     (src/thor/pipelines.py:41-46)
      xv.x = p.mu_next * ((xv.x - p.sigma * ev.x) * inv_mu) + p.sigma_next * ev.x;
      xv.y = p.mu_next * ((xv.y - p.sigma * ev.y) * inv_mu) + p.sigma_next * ev.y;
      xv.z = p.mu_next * ((xv.z - p.sigma * ev.z) * inv_mu) + p.sigma_next * ev.z;
      xv.w = p.mu_next * ((xv.w - p.sigma * ev.w) * inv_mu) + p.sigma_next * ev.w;
      bad |= !(isfinite(xv.x) && isfinite(xv.y) && isfinite(xv.z) && isfinite(xv.w));
      *reinterpret_cast<float4*>(p.x + o) = xv;
    } else {
      sq += ev.x * ev.x + ev.y * ev.y + ev.z * ev.z + ev.w * ev.w;
      *reinterpret_cast<float4*>(p.eps_out + o) = ev;
    }
  }
  if (p.mode == 0) {
    if (__any_sync(0xffffffffu, bad) && lane == 0) atomicOr(p.nan_flag, 1);
  } else {
    __shared__ float wsum[32];
    sq = warp_sum(sq);
    if (lane == 0) wsum[warp] = sq;
    __syncthreads();
    if (warp == 0) {
      float v = (lane < (blockDim.x >> 5)) ? wsum[lane] : 0.f;
      v = warp_sum(v);
      if (lane == 0) p.partials[blockIdx.y * gridDim.x + blockIdx.x] = v;
    }
  }
}

// Deterministic final reduction of the per-CTA partials (fixed order, double accumulate).
__global__ void reduce_partials_kernel(const float* __restrict__ partials, int n, double* __restrict__ out) {
  __shared__ double sh[256];
  double acc = 0.0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) acc += static_cast<double>(partials[i]);
  sh[threadIdx.x] = acc;
  __syncthreads();
  for (int o = blockDim.x / 2; o > 0; o >>= 1) {
    if (threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) out[0] = sh[0];
}

// Philox4x32-10 (counter-based; keyed by seed, counter = (global pixel index, step id)).
__device__ __forceinline__ uint4 philox4x32_10(uint4 ctr, uint2 key) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, ctr.x), lo0 = 0xD2511F53u * ctr.x;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, ctr.z), lo1 = 0xCD9E8D57u * ctr.z;
    ctr = make_uint4(hi1 ^ ctr.y ^ key.x, lo1, hi0 ^ ctr.w ^ key.y, lo0);
    key.x += 0x9E3779B9u;
    key.y += 0xBB67AE85u;
  }
  return ctr;
}
__device__ __forceinline__ float2 box_muller(uint32_t a, uint32_t b) {
  const float u1 = (static_cast<float>(a) + 0.5f) * 2.3283064365386963e-10f;  // (0, 1)
  const float u2 = (static_cast<float>(b) + 0.5f) * 2.3283064365386963e-10f;
  const float r = sqrtf(-2.0f * logf(u1));
  float sn, cs;
  sincospif(2.0f * u2, &sn, &cs);
  return make_float2(r * cs, r * sn);
}

// x <- x - (delta * eps + sqrt(2 delta) z) * sigma' ,  delta = tau / mean(eps^2)   (src/thor/pipelines.py:84-87)
// sumsq: the trajectory-global sum of eps^2 (already all-reduced when time-sharded); count = L*C*H*W.
// z == null: draw z on chip with Philox keyed by the GLOBAL pixel index (identical for any sharding).
__global__ void corrector_update_kernel(float* __restrict__ x, const float* __restrict__ eps,
                                        const float* __restrict__ z, const double* __restrict__ sumsq, double count,
                                        float tau, float sigma_next, long long pix0_global, long long npix,
                                        unsigned long long seed, unsigned int step_id, int* nan_flag) {
  const float delta = tau / static_cast<float>(sumsq[0] / count);
  const float zs = sqrtf(2.0f * delta);
  bool bad = false;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < npix;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    float4 xv = *reinterpret_cast<const float4*>(x + i * 4);
    const float4 ev = *reinterpret_cast<const float4*>(eps + i * 4);
    float4 zv;
    if (z) {
      zv = *reinterpret_cast<const float4*>(z + i * 4);
    } else {
      const unsigned long long g = static_cast<unsigned long long>(pix0_global + i);
      const uint4 r = philox4x32_10(make_uint4(static_cast<uint32_t>(g), static_cast<uint32_t>(g >> 32), step_id, 0u),
                                    make_uint2(static_cast<uint32_t>(seed), static_cast<uint32_t>(seed >> 32)));
      const float2 n0 = box_muller(r.x, r.y), n1 = box_muller(r.z, r.w);
      zv = make_float4(n0.x, n0.y, n1.x, n1.y);
    }
    xv.x -= (delta * ev.x + zs * zv.x) * sigma_next;
    xv.y -= (delta * ev.y + zs * zv.y) * sigma_next;
    xv.z -= (delta * ev.z + zs * zv.z) * sigma_next;
    xv.w -= (delta * ev.w + zs * zv.w) * sigma_next;
    bad |= !(isfinite(xv.x) && isfinite(xv.y) && isfinite(xv.z) && isfinite(xv.w));
    *reinterpret_cast<float4*>(x + i * 4) = xv;
  }
  if (bad) atomicOr(nan_flag, 1);
}

}  // namespace c2w
