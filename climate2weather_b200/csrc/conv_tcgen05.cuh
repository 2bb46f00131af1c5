// K1 — implicit-GEMM 3x3 convolution / plain GEMM on tcgen05 tensor cores (sm_100a).
//
// Computes every Conv2d of the reference UNet (model/nn.py:155,157,169,185,193,194) and the
// attention 1x1 Conv1d (model/nn.py:45,47) as  D[M, N] = A[M, K] * B[N, K]^T  with
//   M = pixels (n, h, w flattened, NHWC activations, bf16),
//   N = output channels, K = taps * Cin  (k index = (r*3 + s) * Cin + c).
//
// Feeding:   A is never materialised.  For filter tap (r, s) the 128-pixel A tile (whole output rows)
//            is a SHIFTED BOX of the NHWC input: one 4-D TMA load {64 ch, W, tile_h, tile_n} at
//            (c0, s-1, h0+r-1, n0); out-of-bounds rows/columns are zero-filled by the TMA unit,
//            which is exactly padding=1 / padding_mode=zeros (configs/sda_unet.yml:16).
//            B (packed weights [Cout, 9*Cin], K-major) is a 2-D TMA load {64, BN}.
//            Both land in 128B-swizzled K-major tiles that tcgen05.mma consumes directly.
// Pipeline:  warp 0 = TMA producer, warp 1 = MMA issuer (single thread), warp 2 = TMEM allocator,
//            warps 4..7 = epilogue.  STAGES-deep smem ring (full/empty mbarriers) and a
//            double-buffered TMEM accumulator (tmem_full/tmem_empty) so the epilogue of tile i
//            overlaps the main loop of tile i+1.  Persistent: grid = #SMs, static round-robin tiles.
// Epilogue:  TMEM -> registers (tcgen05.ld 32x32b.x32) -> +bias [-> SiLU | + residual] -> bf16 NHWC,
//            or the fused window compose (src/thor/score.py:76-88,111-141) for the last conv.
#pragma once
#include <cuda_bf16.h>
#include <stdio.h>

#include "c2w_ptx.cuh"

namespace c2w {

enum EpiMode : int {
  EPI_BIAS = 0,       // out = acc + bias                       -> bf16
  EPI_BIAS_SILU = 1,  // out = silu(acc + bias)                 -> bf16
  EPI_BIAS_RES = 2,   // out = acc + bias + res                 -> bf16   (res may alias out)
  EPI_COMPOSE = 3,    // centre-pick / edge-fill compose        -> fp32 eps [L, H, W, 4]
  EPI_F32 = 4,        // out = acc + bias                       -> fp32 [M, ldc]
};

struct ConvParams {
  // main loop
  int taps;        // 9: 3x3 pad 1 stride 1 on NHWC input; 1: A is a plain [M, K] matrix
  int cin_blocks;  // Cin / 64
  int num_m_tiles, num_n_tiles;
  int m_total;  // valid rows
  int tile_h, tile_n, tiles_per_img;
  // epilogue
  int mode;
  int ldc;  // output row pitch (elements)
  const float* bias;
  const __nv_bfloat16* res;
  __nv_bfloat16* out;
  float* out_f32;
  // EPI_COMPOSE
  float* eps;       // [local frames, H, W, 4] fp32
  int hw;           // H*W
  int order_k;      // Markov order k (window = 2k+1)
  int win_first;    // global index of the first window of this launch
  int win_last_global;  // global index of the last window of the trajectory (Nw - 1)
  int frame_base;   // global frame index of eps[0]
  int dbg_skip_loads;  // diagnostics only: after the ring is primed, signal `full` without issuing TMA loads
};

constexpr int kBlockM = 128;
constexpr int kBlockK = 64;
constexpr int kATileBytes = kBlockM * kBlockK * 2;  // 16 KB
constexpr int kConvThreads = 256;
constexpr int kSmemLimit = 232448;  // 227 KB

template <int BN>
struct ConvCfg {
  // A pipeline stage holds kSub K-sub-blocks of 64 (one 128 B swizzle row each).  With BN <= 128 a 64-wide
  // sub-block is only 256 tensor-pipe cycles of work, less than one barrier round trip of the issuing thread,
  // so two sub-blocks share one full/empty barrier pair (measured: issue-bound at kSub = 1, profiles/).
  static constexpr int kSub = (BN <= 128) ? 2 : 1;
  static constexpr int kBTileBytes = BN * kBlockK * 2;
  static constexpr int kSubBytes = kATileBytes + kBTileBytes;
  static constexpr int kStageBytes = kSub * kSubBytes;
  static constexpr int kBarrierBytes = 256;
  static constexpr int kStagesRaw = (kSmemLimit - 1024 - kBarrierBytes) / kStageBytes;
  static constexpr int kStages = kStagesRaw > 8 ? 8 : kStagesRaw;
  static constexpr int kSmemBytes = kStages * kStageBytes + kBarrierBytes + 1024;
  static constexpr int kTmemCols = (2 * BN <= 128) ? 128 : (2 * BN <= 256) ? 256 : 512;
  static_assert(BN % 64 == 0 && BN >= 64 && BN <= 256, "BN must be 64..256, multiple of 64");
  static_assert(kStages >= 3, "pipeline too shallow");
};

// SiLU without the IEEE-division slow path: ex2.approx + rcp.approx, branch-free so the 32 independent
// evaluations of a chunk interleave (the IEEE form serialises into ~100 clk per element).
__device__ __forceinline__ float silu_f(float x) { return __fdividef(x, 1.0f + __expf(-x)); }

// One 32-column chunk of one accumulator row: v = raw fp32 bits from TMEM, col0 = first output channel.
__device__ __forceinline__ void epilogue_chunk(const ConvParams& p, const uint32_t (&v)[32], int col0, int m,
                                               bool valid) {
  float f[32];
  const float4* b4 = reinterpret_cast<const float4*>(p.bias + col0);
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const float4 b = __ldg(b4 + i);
    f[4 * i + 0] = __uint_as_float(v[4 * i + 0]) + b.x;
    f[4 * i + 1] = __uint_as_float(v[4 * i + 1]) + b.y;
    f[4 * i + 2] = __uint_as_float(v[4 * i + 2]) + b.z;
    f[4 * i + 3] = __uint_as_float(v[4 * i + 3]) + b.w;
  }
  if (p.mode == EPI_COMPOSE) {
    if (valid) {
      const int n_img = m / p.hw;
      const int pix = m - n_img * p.hw;
      const int win = p.win_first + n_img;
#pragma unroll
      for (int g = 0; g < 8; ++g) {
        const int tau = (col0 >> 2) + g;  // window slot of channels [4*tau, 4*tau+4)
        const bool take = (tau == p.order_k) || (win == 0 && tau < p.order_k) ||
                          (win == p.win_last_global && tau > p.order_k && tau <= 2 * p.order_k);
        if (take) {
          const long long fl = static_cast<long long>(win + tau - p.frame_base);
          float4 o = make_float4(f[4 * g], f[4 * g + 1], f[4 * g + 2], f[4 * g + 3]);
          *reinterpret_cast<float4*>(p.eps + (fl * p.hw + pix) * 4) = o;
        }
      }
    }
  } else if (p.mode == EPI_F32) {
    if (valid) {
      float4* o4 = reinterpret_cast<float4*>(p.out_f32 + static_cast<size_t>(m) * p.ldc + col0);
#pragma unroll
      for (int i = 0; i < 8; ++i) o4[i] = make_float4(f[4 * i], f[4 * i + 1], f[4 * i + 2], f[4 * i + 3]);
    }
  } else {
    if (p.mode == EPI_BIAS_SILU) {
#pragma unroll
      for (int i = 0; i < 32; ++i) f[i] = silu_f(f[i]);
    }
    if (valid) {
      const size_t off = static_cast<size_t>(m) * p.ldc + col0;
      if (p.mode == EPI_BIAS_RES) {
        const uint4* r4 = reinterpret_cast<const uint4*>(p.res + off);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const uint4 r = r4[i];
          const uint32_t w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            __nv_bfloat162 h = *reinterpret_cast<const __nv_bfloat162*>(&w[j]);
            f[8 * i + 2 * j] += __low2float(h);
            f[8 * i + 2 * j + 1] += __high2float(h);
          }
        }
      }
      uint4* o4 = reinterpret_cast<uint4*>(p.out + off);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        uint32_t w[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          __nv_bfloat162 h = __floats2bfloat162_rn(f[8 * i + 2 * j], f[8 * i + 2 * j + 1]);
          w[j] = *reinterpret_cast<uint32_t*>(&h);
        }
        o4[i] = make_uint4(w[0], w[1], w[2], w[3]);
      }
    }
  }
}

template <int BN>
__global__ void __launch_bounds__(kConvThreads, 1)
conv_gemm_tcgen05_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                         const ConvParams p) {
  using Cfg = ConvCfg<BN>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smA = smem;
  uint8_t* smB = smem + Cfg::kStages * Cfg::kSub * kATileBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::kStages * Cfg::kStageBytes);
  uint64_t* full = bars;
  uint64_t* empty = bars + Cfg::kStages;
  uint64_t* tmem_full = bars + 2 * Cfg::kStages;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp_idx = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int num_tiles = p.num_m_tiles * p.num_n_tiles;
  const int num_kb = p.taps * p.cin_blocks;

  if (warp_idx == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
  }
  if (warp_idx == 1 && lane == 0) {
    for (int s = 0; s < Cfg::kStages; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tmem_full[a], 1);
      mbar_init(&tmem_empty[a], 4);
    }
    fence_mbar_init();
  }
  if (warp_idx == 2) tmem_alloc(tmem_slot, Cfg::kTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // Producer and MMA warps run CONVERGED (all 32 lanes wait on the barriers) and issue under elect_one():
  // operands stay warp-uniform, so ptxas keeps descriptors/coordinates in uniform registers instead of wrapping
  // every tcgen05/TMA instruction in an R2UR waterfall (measured: 41% -> tensor pipe, see profiles/).
  if (warp_idx == 0) {
    // ------------------------------------------------------------ TMA producer
    int stage = 0;
    uint32_t phase = 0;
    int issued = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int nt = tile % p.num_n_tiles;
      const int mt = tile / p.num_n_tiles;
      int b1, b2, b3;
      if (p.taps == 9) {
        b1 = 0;
        b2 = (mt % p.tiles_per_img) * p.tile_h;
        b3 = (mt / p.tiles_per_img) * p.tile_n;
      } else {
        b1 = mt * kBlockM;
        b2 = 0;
        b3 = 0;
      }
      for (int kb = 0; kb < num_kb; kb += Cfg::kSub) {
        const int nsub = (num_kb - kb) < Cfg::kSub ? (num_kb - kb) : Cfg::kSub;
        mbar_wait(&empty[stage], phase ^ 1);
        if (elect_one()) {
          if (p.dbg_skip_loads && issued >= Cfg::kStages) {
            mbar_arrive(&full[stage]);
          } else {
            mbar_arrive_expect_tx(&full[stage], nsub * Cfg::kSubBytes);
            for (int sub = 0; sub < nsub; ++sub) {
              const int kk = kb + sub;
              const int tap = kk / p.cin_blocks;
              const int cb = kk - tap * p.cin_blocks;
              const int dr = (p.taps == 9) ? tap / 3 - 1 : 0;
              const int ds = (p.taps == 9) ? tap % 3 - 1 : 0;
              const int slot = stage * Cfg::kSub + sub;
              tma_load_4d(&tmA, &full[stage], smA + slot * kATileBytes, cb * kBlockK, b1 + ds, b2 + dr, b3);
              tma_load_2d(&tmB, &full[stage], smB + slot * Cfg::kBTileBytes, kk * kBlockK, nt * BN);
            }
          }
        }
        __syncwarp();
        ++issued;
        if (++stage == Cfg::kStages) {
          stage = 0;
          phase ^= 1;
        }
      }
    }
  } else if (warp_idx == 1) {
    // ------------------------------------------------------------ MMA issuer (one elected lane issues)
    constexpr uint32_t idesc = umma_idesc_bf16(kBlockM, BN);
    const uint64_t adesc0 = umma_desc_kmajor_sw128(smem_u32(smA));
    const uint64_t bdesc0 = umma_desc_kmajor_sw128(smem_u32(smB));
    int stage = 0;
    uint32_t phase = 0;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
      tc_fence_after();
      const uint32_t tmem_d = tmem_base + acc * BN;
      for (int kb = 0; kb < num_kb; kb += Cfg::kSub) {
        const int nsub = (num_kb - kb) < Cfg::kSub ? (num_kb - kb) : Cfg::kSub;
        mbar_wait(&full[stage], phase);
        tc_fence_after();
        if (elect_one()) {
          // descriptor start-address field is (addr >> 4): slot stride and the 32 B K-advance are plain adds
          const uint64_t adesc = adesc0 + static_cast<uint64_t>(stage * Cfg::kSub * (kATileBytes >> 4));
          const uint64_t bdesc = bdesc0 + static_cast<uint64_t>(stage * Cfg::kSub * (Cfg::kBTileBytes >> 4));
#pragma unroll
          for (int sub = 0; sub < Cfg::kSub; ++sub) {
            if (sub < nsub) {
#pragma unroll
              for (int k = 0; k < kBlockK / 16; ++k)
                umma_bf16(tmem_d, adesc + sub * (kATileBytes >> 4) + 2 * k, bdesc + sub * (Cfg::kBTileBytes >> 4) + 2 * k,
                          idesc, (kb | sub | k) != 0);
            }
          }
          umma_commit(&empty[stage]);
          if (kb + Cfg::kSub >= num_kb) umma_commit(&tmem_full[acc]);
        }
        __syncwarp();
        if (++stage == Cfg::kStages) {
          stage = 0;
          phase ^= 1;
        }
      }
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
  } else if (warp_idx >= 4) {
    // ------------------------------------------------------------ epilogue (4 warps = 128 rows)
    const int q = warp_idx & 3;  // TMEM lane quarter this warp may access
    const int row = q * 32 + lane;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int nt = tile % p.num_n_tiles;
      const int mt = tile / p.num_n_tiles;
      const int m = mt * kBlockM + row;
      const bool valid = m < p.m_total;
      mbar_wait(&tmem_full[acc], acc_phase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * BN;
      // two register buffers: the tcgen05.ld of chunk i+1 is in flight while chunk i is processed
      uint32_t va[32], vb[32];
      tmem_ld_32x32(taddr, va);
#pragma unroll 1
      for (int c = 0; c < BN; c += 64) {
        tmem_ld_wait();
        tmem_ld_32x32(taddr + c + 32, vb);
        epilogue_chunk(p, va, nt * BN + c, m, valid);
        tmem_ld_wait();
        if (c + 64 < BN) tmem_ld_32x32(taddr + c + 64, va);
        epilogue_chunk(p, vb, nt * BN + c + 32, m, valid);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty[acc]);
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp_idx == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::kTmemCols);
  }
}

// ------------------------------------------------------------------------------------------------
// Host side

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                    CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                    CUtensorMapFloatOOBfill);

inline PFN_encodeTiled get_encode_tiled() {
  static PFN_encodeTiled fn = nullptr;
  if (!fn) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres);
    if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || !ptr) return nullptr;
    fn = reinterpret_cast<PFN_encodeTiled>(ptr);
  }
  return fn;
}

// bf16 tensor map with 128B swizzle and zero OOB fill.  dims/box are innermost-first.
inline bool make_tmap_bf16(CUtensorMap* m, const void* base, int rank, const uint64_t* dims, const uint32_t* box) {
  PFN_encodeTiled enc = get_encode_tiled();
  if (!enc) return false;
  cuuint64_t gdim[5];
  cuuint64_t gstride[4];
  cuuint32_t bdim[5];
  cuuint32_t estr[5];
  uint64_t pitch = 2;
  for (int i = 0; i < rank; ++i) {
    gdim[i] = dims[i];
    bdim[i] = box[i];
    estr[i] = 1;
    pitch *= dims[i];
    if (i < rank - 1) gstride[i] = pitch;
  }
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, rank, const_cast<void*>(base), gdim, gstride, bdim, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS;
}

// One prepared launch of K1: tensor maps + parameters.  Built once per (layer, batch) and replayed.
struct ConvLaunch {
  CUtensorMap tmA, tmB;
  ConvParams p;
  int bn;
  int grid;
};

// Geometry of the activation operand.
//  conv3x3: x is NHWC [n_img, H, W, cin] bf16 (cin % 64 == 0), output pixels = n_img*H*W
//  gemm   : x is [m, k] bf16 row-major (k % 64 == 0)
inline bool conv_launch_init(ConvLaunch* L, bool is_conv3x3, const __nv_bfloat16* x, int n_img, int H, int W, int cin,
                             const __nv_bfloat16* w_packed, int cout_pad, int bn, int num_sms) {
  ConvParams& p = L->p;
  memset(&p, 0, sizeof(p));
  L->bn = bn;
  if (cin % kBlockK != 0 || cout_pad % bn != 0) return false;
  p.cin_blocks = cin / kBlockK;
  p.num_n_tiles = cout_pad / bn;
  const long long m_total = static_cast<long long>(n_img) * H * W;
  p.m_total = static_cast<int>(m_total);
  p.num_m_tiles = static_cast<int>((m_total + kBlockM - 1) / kBlockM);
  if (is_conv3x3) {
    if (W > kBlockM || kBlockM % W != 0) return false;
    int th = kBlockM / W;
    if (th > H) th = H;
    if (H % th != 0) return false;
    const int tn = kBlockM / (W * th);
    p.taps = 9;
    p.tile_h = th;
    p.tile_n = tn;
    p.tiles_per_img = H / th;
    const uint64_t dims[4] = {(uint64_t)cin, (uint64_t)W, (uint64_t)H, (uint64_t)n_img};
    const uint32_t box[4] = {(uint32_t)kBlockK, (uint32_t)W, (uint32_t)th, (uint32_t)tn};
    if (!make_tmap_bf16(&L->tmA, x, 4, dims, box)) return false;
  } else {
    p.taps = 1;
    p.tile_h = 1;
    p.tile_n = 1;
    p.tiles_per_img = 1;
    const uint64_t dims[4] = {(uint64_t)cin, (uint64_t)m_total, 1, 1};
    const uint32_t box[4] = {(uint32_t)kBlockK, (uint32_t)kBlockM, 1, 1};
    if (!make_tmap_bf16(&L->tmA, x, 4, dims, box)) return false;
  }
  {
    const uint64_t dims[2] = {(uint64_t)p.taps * cin, (uint64_t)cout_pad};
    const uint32_t box[2] = {(uint32_t)kBlockK, (uint32_t)bn};
    if (!make_tmap_bf16(&L->tmB, w_packed, 2, dims, box)) return false;
  }
  const int tiles = p.num_m_tiles * p.num_n_tiles;
  L->grid = tiles < num_sms ? tiles : num_sms;
  p.ldc = cout_pad;
  return true;
}

template <int BN>
inline cudaError_t conv_launch_bn(const ConvLaunch& L, cudaStream_t stream) {
  using Cfg = ConvCfg<BN>;
  static bool attr_done = false;
  if (!attr_done) {
    cudaError_t e = cudaFuncSetAttribute(conv_gemm_tcgen05_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         Cfg::kSmemBytes);
    if (e != cudaSuccess) return e;
    attr_done = true;
  }
  conv_gemm_tcgen05_kernel<BN><<<L.grid, kConvThreads, Cfg::kSmemBytes, stream>>>(L.tmA, L.tmB, L.p);
  return cudaGetLastError();
}

inline cudaError_t conv_launch(const ConvLaunch& L, cudaStream_t stream) {
  switch (L.bn) {
    case 64: return conv_launch_bn<64>(L, stream);
    case 128: return conv_launch_bn<128>(L, stream);
    case 192: return conv_launch_bn<192>(L, stream);
    case 256: return conv_launch_bn<256>(L, stream);
    default: return cudaErrorInvalidValue;
  }
}

// N tile for a (padded) output-channel count.
inline int conv_pick_bn(int cout_pad) {
  if (cout_pad % 256 == 0) return 256;
  if (cout_pad % 192 == 0) return 192;
  if (cout_pad % 128 == 0) return 128;
  return 64;
}

}  // namespace c2w
