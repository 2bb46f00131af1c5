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
//            which is exactly padding=1 / padding_mode=zeros (configs/sda_unet.yml:16).  Stride-2 convs use
//            the same box with element strides {1, 2, 2, 1}.
//            B (packed weights [Cout, 9*Cin], K-major) is a 2-D TMA load {64, BN}.
//            Both land in 128B-swizzled K-major tiles that tcgen05.mma consumes directly.
// Pipeline:  warp 0 = TMA producer, warp 1 = MMA issuer (single thread), warp 2 = TMEM allocator,
//            warps 4..11 = epilogue.  STAGES-deep smem ring (full/empty mbarriers) and a
//            double-buffered TMEM accumulator (tmem_full/tmem_empty) so the epilogue of tile i
//            overlaps the main loop of tile i+1.  Persistent: grid = #SMs, static round-robin tiles.
//            Optionally two CTAs (a cluster) share one MMA (cta_group::2, 256 x BN tile).
// Epilogue:  TMEM -> registers (tcgen05.ld 32x32b.x32) -> +bias [-> SiLU | + residual] -> bf16 -> 128B-swizzled
//            staging tile in shared memory -> TMA store (the L1 data pipe is what bounds this kernel: per-thread
//            row stores cost 8x the wavefronts of the staged path, profiles/r01a_ncu_conv_G2.csv).  The residual
//            tile is prefetched into the staging tile by TMA; the fused channel LayerNorm re-reads the staged
//            row, normalises in place and stores a second tensor.  The last conv of the UNet writes fp32 /
//            the fused window compose (src/thor/score.py:76-88,111-141) straight from registers.
#pragma once
#include <cuda_bf16.h>
#include <stdio.h>

#include "c2w_ptx.cuh"

namespace c2w {

enum EpiMode : int {
  EPI_BIAS = 0,       // out = acc + bias                       -> bf16
  EPI_BIAS_SILU = 1,  // out = silu(acc + bias)                 -> bf16
  EPI_BIAS_RES = 2,   // out += acc + bias                      -> bf16   (in place)
  EPI_COMPOSE = 3,    // centre-pick / edge-fill compose        -> fp32 eps [L, H, W, 4]
  EPI_F32 = 4,        // out = acc + bias                       -> fp32 [M, ldc]
  EPI_DSILU = 5,      // out = (acc + bias) * silu'(aux)        -> bf16   (VJP: aux = stashed pre-activation)
};

struct ConvParams {
  // main loop
  int taps;        // 9: 3x3 pad 1 (stride 1 or 2) on NHWC input; 1: A is a plain [M, K] matrix
  int stride;      // taps == 9: input pixel = stride * output pixel + tap - 1
  int cin_blocks;  // Cin / 64
  int num_m_tiles, num_n_tiles;
  int m_total;  // valid rows
  int tile_h, tile_n, tiles_per_img;
  int num_stages;   // depth of the A/B ring
  int num_staging;  // epilogue staging tiles (2 for convs with an auxiliary input tile or a second output)
  int dual_out;     // EPI_BIAS_SILU only: also store the pre-activation (acc + bias) as a second bf16 tensor
  // epilogue
  int mode;
  int ldc;  // output row pitch (elements)
  const float* bias;
  float* out_f32;
  // fused channel LayerNorm of the output row (model/nn.py:154,183; needs one N tile = all channels):
  //   ln_out[pixel] = (v - mean_C v) / sqrt(var_C v + eps),  v = bf16(out) + ln_mod, unbiased variance;
  //   ln_up: every pixel is written to its 2x2 nearest-neighbour block of a [n, 2H, 2W, C] tensor (model/nn.py:184)
  __nv_bfloat16* ln_out;
  float* ln_inv;  // optional: 1 / sqrt(var + eps) per row (stash for the LayerNorm backward)
  const float* ln_mod;
  int ln_up, ln_H, ln_W;  // ln_H x ln_W: output image of this conv (for the upsampled addressing)
  float ln_eps;
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
constexpr int kEpiWarps = 8;                        // two warps per TMEM lane quarter, each half of the columns
constexpr int kEpiThreads = kEpiWarps * 32;
constexpr int kConvThreads = 128 + kEpiThreads;
constexpr int kSmemLimit = 232448;  // 227 KB
constexpr int kEpiBarrier = 1;      // named barrier of the epilogue warps

// BN   : N tile (output channels per tile).
// CG   : CTAs per MMA (tcgen05 cta_group).  CG == 2: a cluster of two CTAs computes a 256 x BN tile, each CTA
//        holding its own 128 A rows and HALF of the B rows in shared memory.
// Shared memory: [A ring][B ring][1 or 2 staging tiles][barriers]; the ring depth is a launch-time number (what is
// left after the staging tiles: residual convs double-buffer them so the residual prefetch runs a tile ahead).
template <int BN, int CG>
struct ConvCfg {
  // A pipeline stage holds kSub K-sub-blocks of 64 (one 128 B swizzle row each).  With BN <= 128 a 64-wide
  // sub-block is only 256 tensor-pipe cycles of work, less than one barrier round trip of the issuing thread,
  // so two sub-blocks share one full/empty barrier pair (measured both ways, profiles/).
  static constexpr int kSub = (BN > 128) ? 1 : 2;
  static constexpr int kBRows = BN / CG;                 // B rows held by one CTA
  static constexpr int kBTileBytes = kBRows * kBlockK * 2;
  static constexpr int kSubBytes = kATileBytes + kBTileBytes;
  static constexpr int kStageBytes = kSub * kSubBytes;
  static constexpr int kStagingBytes = (BN / 64) * kATileBytes;  // bf16 [BN/64 boxes][128 rows][64 ch], swizzled
  static constexpr int kBarrierBytes = 256;
  static constexpr int kMaxStages = 8;
  static constexpr int kTmemCols = (2 * BN <= 128) ? 128 : (2 * BN <= 256) ? 256 : 512;
  static constexpr int stages_for(int num_staging) {
    const int n = (kSmemLimit - 1024 - kBarrierBytes - num_staging * kStagingBytes) / kStageBytes;
    return n > kMaxStages ? kMaxStages : n;
  }
  static constexpr int smem_bytes(int num_staging) {
    return stages_for(num_staging) * kStageBytes + num_staging * kStagingBytes + kBarrierBytes + 1024;
  }
  static_assert(BN % 64 == 0 && BN >= 64 && BN <= 256, "BN must be 64..256, multiple of 64");
  static_assert(CG == 1 || CG == 2, "cta_group is 1 or 2");
  static_assert(stages_for(2) >= 2, "pipeline too shallow");
  static_assert(2 * kMaxStages + 6 <= kBarrierBytes / 8 - 1, "barrier block too small");
};

// SiLU = x * sigmoid(x) = h + h * tanh(h), h = x / 2: one MUFU op per element (tanh.approx, rel. error 2^-11, far
// below the bf16 output rounding) instead of ex2 + rcp — the epilogue's SFU budget is 16 ops/clk/SM.
__device__ __forceinline__ float silu_f(float x) {
  const float h = 0.5f * x;
  return fmaf(h, tanh_approx(h), h);
}

// Direct-to-global epilogue of one 32-column chunk of one accumulator row (final conv only: fp32 / compose).
__device__ __forceinline__ void epilogue_chunk_direct(const ConvParams& p, const uint32_t (&v)[32], int col0, int m,
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
  if (!valid) return;
  if (p.mode == EPI_COMPOSE) {
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
  } else {
    float4* o4 = reinterpret_cast<float4*>(p.out_f32 + static_cast<size_t>(m) * p.ldc + col0);
#pragma unroll
    for (int i = 0; i < 8; ++i) o4[i] = make_float4(f[4 * i], f[4 * i + 1], f[4 * i + 2], f[4 * i + 3]);
  }
}

// silu'(x) = s (1 + x (1 - s)), s = sigmoid(x) = (1 + tanh(x/2)) / 2
__device__ __forceinline__ float dsilu_f(float x) {
  const float sg = fmaf(0.5f, tanh_approx(0.5f * x), 0.5f);
  return sg * fmaf(x, 1.0f - sg, 1.0f);
}

// Staged epilogue of one 32-column chunk: bf16 result into the 128B-swizzled staging tile — in place over the
// TMA-prefetched auxiliary tile (residual in EPI_BIAS_RES, stashed pre-activation in EPI_DSILU).  `stg_row` = shared
// address of this row in box 0 of the tile; `stg2_row` = same for the second output tile (dual_out).
__device__ __forceinline__ void epilogue_chunk_staged(const ConvParams& p, const uint32_t (&v)[32], int gcol, int col,
                                                      uint32_t stg_row, uint32_t stg2_row, int row) {
  float f[32];
  const float4* b4 = reinterpret_cast<const float4*>(p.bias + gcol);
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const float4 b = __ldg(b4 + i);
    f[4 * i + 0] = __uint_as_float(v[4 * i + 0]) + b.x;
    f[4 * i + 1] = __uint_as_float(v[4 * i + 1]) + b.y;
    f[4 * i + 2] = __uint_as_float(v[4 * i + 2]) + b.z;
    f[4 * i + 3] = __uint_as_float(v[4 * i + 3]) + b.w;
  }
  const uint32_t boff = static_cast<uint32_t>(col >> 6) * kATileBytes;
  const int j0 = (col & 63) >> 3;  // first 16 B chunk of this 32-column group inside the 128 B row
  if (p.mode == EPI_BIAS_SILU) {
    if (p.dual_out) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        uint32_t o[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          __nv_bfloat162 h = __floats2bfloat162_rn(f[8 * i + 2 * j], f[8 * i + 2 * j + 1]);
          o[j] = *reinterpret_cast<uint32_t*>(&h);
        }
        st_shared_v4(stg2_row + boff + (static_cast<uint32_t>((j0 + i) ^ (row & 7)) << 4), o[0], o[1], o[2], o[3]);
      }
    }
#pragma unroll
    for (int i = 0; i < 32; ++i) f[i] = silu_f(f[i]);
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const uint32_t addr = stg_row + boff + (static_cast<uint32_t>((j0 + i) ^ (row & 7)) << 4);
    if (p.mode == EPI_BIAS_RES || p.mode == EPI_DSILU) {
      const uint4 r = ld_shared_v4(addr);
      const uint32_t w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        __nv_bfloat162 h = *reinterpret_cast<const __nv_bfloat162*>(&w[j]);
        if (p.mode == EPI_BIAS_RES) {
          f[8 * i + 2 * j] += __low2float(h);
          f[8 * i + 2 * j + 1] += __high2float(h);
        } else {
          f[8 * i + 2 * j] *= dsilu_f(__low2float(h));
          f[8 * i + 2 * j + 1] *= dsilu_f(__high2float(h));
        }
      }
    }
    uint32_t o[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      __nv_bfloat162 h = __floats2bfloat162_rn(f[8 * i + 2 * j], f[8 * i + 2 * j + 1]);
      o[j] = *reinterpret_cast<uint32_t*>(&h);
    }
    st_shared_v4(addr, o[0], o[1], o[2], o[3]);
  }
}

template <int BN, int CG, bool LN>
__global__ void __launch_bounds__(kConvThreads, 1)
conv_gemm_tcgen05_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                         const __grid_constant__ CUtensorMap tmOut, const __grid_constant__ CUtensorMap tmAux,
                         const __grid_constant__ CUtensorMap tmOut2, const ConvParams p) {
  using Cfg = ConvCfg<BN, CG>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int num_stages = p.num_stages;
  uint8_t* smA = smem;
  uint8_t* smB = smem + num_stages * Cfg::kSub * kATileBytes;
  uint8_t* stg0 = smem + num_stages * Cfg::kStageBytes;  // epilogue staging tile(s)
  uint64_t* bars = reinterpret_cast<uint64_t*>(stg0 + p.num_staging * Cfg::kStagingBytes);
  uint64_t* full = bars;
  uint64_t* empty = bars + Cfg::kMaxStages;
  uint64_t* tmem_full = bars + 2 * Cfg::kMaxStages;
  uint64_t* tmem_empty = tmem_full + 2;
  uint64_t* res_full = tmem_empty + 2;  // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(res_full + 2);

  const int warp_idx = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = (CG == 2) ? cluster_ctarank() : 0u;  // 0 = leader: arms the full barriers, issues the MMAs
  const int group_id = blockIdx.x / CG;
  const int num_groups = gridDim.x / CG;
  const int num_tiles = ((p.num_m_tiles + CG - 1) / CG) * p.num_n_tiles;
  const int num_kb = p.taps * p.cin_blocks;
  const bool staged = p.mode != EPI_COMPOSE && p.mode != EPI_F32;
  const bool has_aux = p.mode == EPI_BIAS_RES || p.mode == EPI_DSILU;  // TMA-prefetched input tile, double-buffered

  if (warp_idx == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    if (staged) tma_prefetch_desc(&tmOut);
    if (has_aux) tma_prefetch_desc(&tmAux);
    if (p.dual_out) tma_prefetch_desc(&tmOut2);
  }
  if (warp_idx == 1 && lane == 0) {
    for (int s = 0; s < num_stages; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tmem_full[a], 1);
      mbar_init(&tmem_empty[a], kEpiWarps * CG);
    }
    mbar_init(&res_full[0], 1);
    mbar_init(&res_full[1], 1);
    fence_mbar_init();
  }
  if (warp_idx == 2) {
    if (CG == 2) tmem_alloc_pair(tmem_slot, Cfg::kTmemCols);
    else tmem_alloc(tmem_slot, Cfg::kTmemCols);
  }
  tc_fence_before();
  if (CG == 2) cluster_sync_all();  // the peer's barriers must be initialised before remote arrives / TMA signals
  else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // Producer and MMA warps run CONVERGED (all 32 lanes wait on the barriers) and issue under elect_one():
  // operands stay warp-uniform, so ptxas keeps descriptors/coordinates in uniform registers instead of wrapping
  // every tcgen05/TMA instruction in an R2UR waterfall.
  if (warp_idx == 0) {
    // ------------------------------------------------------------ TMA producer (every CTA loads its own rows)
    int stage = 0;
    uint32_t phase = 0;
    int issued = 0;
    for (int tile = group_id; tile < num_tiles; tile += num_groups) {
      const int nt = tile % p.num_n_tiles;
      const int mt = (tile / p.num_n_tiles) * CG + rank;
      int b1, b2, b3;
      if (p.taps == 9) {
        b1 = 0;
        b2 = (mt % p.tiles_per_img) * p.tile_h * p.stride;
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
          const uint32_t full_bar = (CG == 2) ? mapa_shared(smem_u32(&full[stage]), 0) : smem_u32(&full[stage]);
          if (p.dbg_skip_loads && issued >= num_stages) {
            if (rank == 0) mbar_arrive(&full[stage]);
          } else {
            if (rank == 0) mbar_arrive_expect_tx(&full[stage], CG * nsub * Cfg::kSubBytes);
            for (int sub = 0; sub < nsub; ++sub) {
              const int kk = kb + sub;
              const int tap = kk / p.cin_blocks;
              const int cb = kk - tap * p.cin_blocks;
              const int dr = (p.taps == 9) ? tap / 3 - 1 : 0;
              const int ds = (p.taps == 9) ? tap % 3 - 1 : 0;
              const int slot = stage * Cfg::kSub + sub;
              if (CG == 2) {
                tma_load_4d_pair(&tmA, full_bar, smA + slot * kATileBytes, cb * kBlockK, b1 + ds, b2 + dr, b3);
                tma_load_2d_pair(&tmB, full_bar, smB + slot * Cfg::kBTileBytes, kk * kBlockK,
                                 nt * BN + rank * Cfg::kBRows);
              } else {
                tma_load_4d(&tmA, &full[stage], smA + slot * kATileBytes, cb * kBlockK, b1 + ds, b2 + dr, b3);
                tma_load_2d(&tmB, &full[stage], smB + slot * Cfg::kBTileBytes, kk * kBlockK, nt * BN);
              }
            }
          }
        }
        __syncwarp();
        ++issued;
        if (++stage == num_stages) {
          stage = 0;
          phase ^= 1;
        }
      }
    }
  } else if (warp_idx == 1 && rank == 0) {
    // ------------------------------------------------------------ MMA issuer (leader CTA, one elected lane issues)
    constexpr uint32_t idesc = umma_idesc_bf16(kBlockM * CG, BN);
    const uint64_t adesc0 = umma_desc_kmajor_sw128(smem_u32(smA));
    const uint64_t bdesc0 = umma_desc_kmajor_sw128(smem_u32(smB));
    int stage = 0;
    uint32_t phase = 0;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = group_id; tile < num_tiles; tile += num_groups) {
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
              for (int k = 0; k < kBlockK / 16; ++k) {
                const uint64_t a = adesc + sub * (kATileBytes >> 4) + 2 * k;
                const uint64_t b = bdesc + sub * (Cfg::kBTileBytes >> 4) + 2 * k;
                if (CG == 2) umma_bf16_pair(tmem_d, a, b, idesc, (kb | sub | k) != 0);
                else umma_bf16(tmem_d, a, b, idesc, (kb | sub | k) != 0);
              }
            }
          }
          if (CG == 2) {
            umma_commit_pair(&empty[stage]);
            if (kb + Cfg::kSub >= num_kb) umma_commit_pair(&tmem_full[acc]);
          } else {
            umma_commit(&empty[stage]);
            if (kb + Cfg::kSub >= num_kb) umma_commit(&tmem_full[acc]);
          }
        }
        __syncwarp();
        if (++stage == num_stages) {
          stage = 0;
          phase ^= 1;
        }
      }
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
  } else if (warp_idx >= 4) {
    // ------------------------------------------------------------ epilogue: 8 warps = 128 rows x 2 column halves
    const int ew = warp_idx - 4;
    const int q = ew & 3;       // TMEM lane quarter this warp may access (warp_idx % 4)
    const int half = ew >> 2;   // column half
    const int row = q * 32 + lane;
    const bool leader = threadIdx.x == 128;  // issues the staging tile's TMA loads / stores
    constexpr int kHalfCols = BN / 2;
    constexpr int kChunks = kHalfCols / 32;  // 32-column chunks per warp (BN = 64: 1 ... BN = 256: 4)
    const int col_base = half * kHalfCols;
    int acc = 0;
    uint32_t acc_phase = 0;
    int it_local = 0;  // tiles processed by this CTA: staging tile it_local % num_staging

    // fused LayerNorm: warp `ew` normalises rows [16 ew, 16 ew + 16); a row is spread over kLPR lanes x 16 B
    constexpr int kLPR = BN / 8;                        // lanes per row (8 channels each)
    constexpr int kRPI = (kLPR >= 32) ? 1 : 32 / kLPR;  // rows per warp iteration
    constexpr int kIters = 16 / kRPI;
    static_assert(!LN || (kLPR == 8 || kLPR == 16 || kLPR == 32), "fused LayerNorm needs BN in {64, 128, 256}");
    const int ln_chunk = lane % kLPR;  // 16 B chunk of the row this lane owns (channels 8*ln_chunk .. +8)
    float ln_m[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) ln_m[e] = (LN && p.ln_mod) ? __ldg(p.ln_mod + (ln_chunk * 8 + e) % BN) : 0.f;

    auto tile_coords = [&](int tile, int& nt, int& mt) {
      nt = tile % p.num_n_tiles;
      mt = (tile / p.num_n_tiles) * CG + rank;
    };
    auto prefetch_residual = [&](int tile, int buf) {  // leader only: staging[buf] <- aux[tile]
      int nt, mt;
      tile_coords(tile, nt, mt);
      uint8_t* dst = stg0 + buf * Cfg::kStagingBytes;
      mbar_arrive_expect_tx(&res_full[buf], Cfg::kStagingBytes);
#pragma unroll
      for (int b = 0; b < BN / 64; ++b)
        tma_load_2d(&tmAux, &res_full[buf], dst + b * kATileBytes, nt * BN + b * 64, mt * kBlockM);
    };
    if (has_aux && leader) {  // two staging tiles: the auxiliary tile is prefetched two tiles ahead
      if (group_id < num_tiles) prefetch_residual(group_id, 0);
      if (group_id + num_groups < num_tiles) prefetch_residual(group_id + num_groups, 1);
    }

    for (int tile = group_id; tile < num_tiles; tile += num_groups) {
      int nt, mt;
      tile_coords(tile, nt, mt);
      const int m = mt * kBlockM + row;
      const bool valid = m < p.m_total;
      mbar_wait(&tmem_full[acc], acc_phase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * BN + col_base;
      const int buf = has_aux ? (it_local & 1) : 0;
      uint8_t* stg = stg0 + buf * Cfg::kStagingBytes;
      uint8_t* stg2 = stg0 + Cfg::kStagingBytes;  // second output tile (dual_out; never together with has_aux)
      const uint32_t stg_u32 = smem_u32(stg);
      const uint32_t stg_row = stg_u32 + static_cast<uint32_t>(row) * 128u;
      const uint32_t stg2_row = smem_u32(stg2) + static_cast<uint32_t>(row) * 128u;
      if (has_aux) mbar_wait(&res_full[buf], (it_local >> 1) & 1);
      // two register buffers: the tcgen05.ld of chunk i+1 is in flight while chunk i is processed
      uint32_t va[32], vb[32];
      tmem_ld_32x32(taddr, va);
#pragma unroll
      for (int c = 0; c < kChunks; ++c) {
        tmem_ld_wait();
        if (c + 1 < kChunks) tmem_ld_32x32(taddr + 32 * (c + 1), (c & 1) ? va : vb);
        const int col = col_base + 32 * c;
        if (staged) epilogue_chunk_staged(p, (c & 1) ? vb : va, nt * BN + col, col, stg_row, stg2_row, row);
        else epilogue_chunk_direct(p, (c & 1) ? vb : va, nt * BN + col, m, valid);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (CG == 2 && rank != 0) mbar_arrive_remote(mapa_shared(smem_u32(&tmem_empty[acc]), 0));
        else mbar_arrive(&tmem_empty[acc]);
      }
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
      ++it_local;
      if (!staged) continue;

      // ---- staged tile -> global (TMA store); rows past the tensor end are clipped by the tensor map
      fence_proxy_async();
      named_bar_sync(kEpiBarrier, kEpiThreads);
      if (leader) {
#pragma unroll
        for (int b = 0; b < BN / 64; ++b) tma_store_2d(&tmOut, stg + b * kATileBytes, nt * BN + b * 64, mt * kBlockM);
        if (p.dual_out) {
#pragma unroll
          for (int b = 0; b < BN / 64; ++b)
            tma_store_2d(&tmOut2, stg2 + b * kATileBytes, nt * BN + b * 64, mt * kBlockM);
        }
        bulk_commit();
      }
      if (LN) {
        // ---- channel LayerNorm of the staged rows (two-pass variance in registers) -> global, coalesced: the
        //      kLPR lanes of a row write its C channels as one contiguous segment (x4 when upsampling).
        //      Batches of kBatch rows per lane keep the shuffle reductions of different rows in flight together.
        constexpr int kBatch = kIters < 8 ? kIters : 8;
#pragma unroll 1
        for (int it0 = 0; it0 < kIters; it0 += kBatch) {
          float v[kBatch][8];
          float s[kBatch];
#pragma unroll
          for (int b = 0; b < kBatch; ++b) {
            const int r = ew * 16 + (it0 + b) * kRPI + lane / kLPR;
            const uint32_t addr = stg_u32 + static_cast<uint32_t>(ln_chunk >> 3) * kATileBytes +
                                  static_cast<uint32_t>(r) * 128u +
                                  (static_cast<uint32_t>((ln_chunk & 7) ^ (r & 7)) << 4);
            const uint4 xr = ld_shared_v4(addr);
            const uint32_t w[4] = {xr.x, xr.y, xr.z, xr.w};
            s[b] = 0.f;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              __nv_bfloat162 h = *reinterpret_cast<const __nv_bfloat162*>(&w[j]);
              v[b][2 * j] = __low2float(h) + ln_m[2 * j];
              v[b][2 * j + 1] = __high2float(h) + ln_m[2 * j + 1];
              s[b] += v[b][2 * j] + v[b][2 * j + 1];
            }
          }
#pragma unroll
          for (int o = kLPR / 2; o > 0; o >>= 1)
#pragma unroll
            for (int b = 0; b < kBatch; ++b) s[b] += __shfl_xor_sync(0xffffffffu, s[b], o);
#pragma unroll
          for (int b = 0; b < kBatch; ++b) {
            const float mean = s[b] * (1.0f / BN);
            float ss = 0.f;
#pragma unroll
            for (int e = 0; e < 8; ++e) {
              v[b][e] -= mean;
              ss += v[b][e] * v[b][e];
            }
            s[b] = ss;
          }
#pragma unroll
          for (int o = kLPR / 2; o > 0; o >>= 1)
#pragma unroll
            for (int b = 0; b < kBatch; ++b) s[b] += __shfl_xor_sync(0xffffffffu, s[b], o);
#pragma unroll
          for (int b = 0; b < kBatch; ++b) {
            const float inv = rsqrtf(s[b] * (1.0f / (BN - 1)) + p.ln_eps);
            uint32_t o4[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              __nv_bfloat162 h = __floats2bfloat162_rn(v[b][2 * j] * inv, v[b][2 * j + 1] * inv);
              o4[j] = *reinterpret_cast<uint32_t*>(&h);
            }
            const uint4 yv = make_uint4(o4[0], o4[1], o4[2], o4[3]);
            const int r = ew * 16 + (it0 + b) * kRPI + lane / kLPR;
            const int mr = mt * kBlockM + r;
            if (mr < p.m_total) {
              if (p.ln_inv != nullptr && ln_chunk == 0) p.ln_inv[mr] = inv;
              if (p.ln_up) {
                const int w0 = mr % p.ln_W;
                const int t = mr / p.ln_W;
                const int h0 = t % p.ln_H;
                const long long n = t / p.ln_H;
                const long long o00 = (n * 2 * p.ln_H + 2 * h0) * 2 * p.ln_W + 2 * w0;
                __nv_bfloat16* dst = p.ln_out + o00 * BN + ln_chunk * 8;
                *reinterpret_cast<uint4*>(dst) = yv;
                *reinterpret_cast<uint4*>(dst + BN) = yv;
                *reinterpret_cast<uint4*>(dst + 2ll * p.ln_W * BN) = yv;
                *reinterpret_cast<uint4*>(dst + (2ll * p.ln_W + 1) * BN) = yv;
              } else {
                *reinterpret_cast<uint4*>(p.ln_out + static_cast<long long>(mr) * BN + ln_chunk * 8) = yv;
              }
            }
          }
        }
      }
      // ---- the staging tile is reused by the next tile once these stores have read it
      if (leader) bulk_wait_read_all();
      named_bar_sync(kEpiBarrier, kEpiThreads);  // store has read the tile AND every warp is done reading it (LN)
      if (leader && has_aux) {
        const int next = tile + 2 * num_groups;  // this staging tile's next user
        if (next < num_tiles) prefetch_residual(next, buf);
      }
    }
    if (leader) bulk_wait_all();  // global writes of the last stores are complete before the CTA retires
  }

  tc_fence_before();
  if (CG == 2) cluster_sync_all();  // the leader's MMAs read the peer's shared memory until the last commit
  else __syncthreads();
  if (warp_idx == 2) {
    tc_fence_after();
    if (CG == 2) tmem_dealloc_pair(tmem_base, Cfg::kTmemCols);
    else tmem_dealloc(tmem_base, Cfg::kTmemCols);
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

inline char* tmap_error_slot() {
  static thread_local char buf[256] = {0};
  return buf;
}

// bf16 tensor map with 128B swizzle and zero OOB fill.  dims/box/estride are innermost-first; with an element
// stride e > 1 in a dimension the box TRAVERSES box[i] elements and loads every e-th one (ceil(box/e) elements).
// strides_bytes (rank - 1 entries, for dims 1..rank-1) defaults to the packed layout.
inline bool make_tmap_bf16(CUtensorMap* m, const void* base, int rank, const uint64_t* dims, const uint32_t* box,
                           const uint32_t* estride = nullptr, const uint64_t* strides_bytes = nullptr) {
  PFN_encodeTiled enc = get_encode_tiled();
  if (!enc) return false;
  // The driver entry point needs the primary context current on THIS thread; a thread that has made no runtime call
  // yet (e.g. PyTorch's autograd worker) has none.  cudaFree(0) is the canonical no-op that binds it.
  static thread_local bool ctx_bound = false;
  if (!ctx_bound) {
    cudaFree(0);
    ctx_bound = true;
  }
  cuuint64_t gdim[5];
  cuuint64_t gstride[4];
  cuuint32_t bdim[5];
  cuuint32_t estr[5];
  uint64_t pitch = 2;
  for (int i = 0; i < rank; ++i) {
    gdim[i] = dims[i];
    bdim[i] = box[i];
    estr[i] = estride ? estride[i] : 1;
    pitch *= dims[i];
    if (i < rank - 1) gstride[i] = strides_bytes ? strides_bytes[i] : pitch;
  }
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, rank, const_cast<void*>(base), gdim, gstride, bdim, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    char* e = tmap_error_slot();
    int o = snprintf(e, 256, "cuTensorMapEncodeTiled -> %d: base %p rank %d dims", (int)r, base, rank);
    for (int i = 0; i < rank && o < 230; ++i) o += snprintf(e + o, 256 - o, " %llu/%u", (unsigned long long)gdim[i], bdim[i]);
  }
  return r == CUDA_SUCCESS;
}

// One prepared launch of K1: tensor maps + parameters.  Built once per (layer, batch) and replayed.
struct ConvLaunch {
  CUtensorMap tmA, tmB, tmOut, tmAux, tmOut2;
  ConvParams p;
  int bn;
  int cg;     // CTAs per MMA (1 or 2)
  int ln;     // fused LayerNorm output
  int grid;   // CTAs (a multiple of cg)
  int n_img, Ho, Wo, cout_pad;
  const void* out_ptr;  // bf16 output bound by conv_launch_set_out (LayerNorm fusion matches on it)
};

// Kernel variant of a launch: a pair of CTAs per MMA whenever there are at least two M tiles.
inline int conv_pick_cg(int num_m_tiles) { return num_m_tiles >= 2 ? 2 : 1; }

inline void conv_set_grid(ConvLaunch* L, int num_sms) {
  const int groups = ((L->p.num_m_tiles + L->cg - 1) / L->cg) * L->p.num_n_tiles;
  const int max_groups = num_sms / L->cg;
  L->grid = (groups < max_groups ? groups : max_groups) * L->cg;
  if (L->grid < L->cg) L->grid = L->cg;
}

// Geometry of the activation operand.
//  conv3x3: x is NHWC [n_img, H, W, cin] bf16 (cin % 64 == 0); stride 1: output [n_img, H, W], stride 2: output
//           [n_img, H/2, W/2] (pad 1 either way)
//  gemm   : x is [m, k] bf16 row-major (k % 64 == 0) with m = n_img * H * W
// variant: -1 = pick; else bit 0 = CTA pair (cta_group::2)
inline bool conv_launch_init(ConvLaunch* L, bool is_conv3x3, const __nv_bfloat16* x, int n_img, int H, int W, int cin,
                             const __nv_bfloat16* w_packed, int cout_pad, int bn, int num_sms, int stride = 1,
                             int variant = -1) {
  ConvParams& p = L->p;
  memset(&p, 0, sizeof(p));
  memset(&L->tmOut, 0, sizeof(CUtensorMap));
  memset(&L->tmAux, 0, sizeof(CUtensorMap));
  memset(&L->tmOut2, 0, sizeof(CUtensorMap));
  L->bn = bn;
  L->ln = 0;
  L->out_ptr = nullptr;
  if (cin % kBlockK != 0 || cout_pad % bn != 0) return false;
  if (stride != 1 && !(stride == 2 && is_conv3x3 && H % 2 == 0 && W % 2 == 0)) return false;
  p.cin_blocks = cin / kBlockK;
  p.num_n_tiles = cout_pad / bn;
  p.stride = stride;
  p.ln_eps = 1e-5f;
  const int Ho = H / stride, Wo = W / stride;
  L->n_img = n_img;
  L->Ho = Ho;
  L->Wo = Wo;
  L->cout_pad = cout_pad;
  const long long m_total = static_cast<long long>(n_img) * Ho * Wo;
  p.m_total = static_cast<int>(m_total);
  p.num_m_tiles = static_cast<int>((m_total + kBlockM - 1) / kBlockM);
  if (is_conv3x3) {
    if (Wo > kBlockM || kBlockM % Wo != 0) return false;
    int th = kBlockM / Wo;
    if (th > Ho) th = Ho;
    if (Ho % th != 0) return false;
    const int tn = kBlockM / (Wo * th);
    p.taps = 9;
    p.tile_h = th;
    p.tile_n = tn;
    p.tiles_per_img = Ho / th;
    const uint64_t dims[4] = {(uint64_t)cin, (uint64_t)W, (uint64_t)H, (uint64_t)n_img};
    const uint32_t box[4] = {(uint32_t)kBlockK, (uint32_t)(Wo * stride), (uint32_t)(th * stride), (uint32_t)tn};
    const uint32_t est[4] = {1u, (uint32_t)stride, (uint32_t)stride, 1u};
    if (!make_tmap_bf16(&L->tmA, x, 4, dims, box, est)) return false;
  } else {
    p.taps = 1;
    p.tile_h = 1;
    p.tile_n = 1;
    p.tiles_per_img = 1;
    const uint64_t dims[4] = {(uint64_t)cin, (uint64_t)m_total, 1, 1};
    const uint32_t box[4] = {(uint32_t)kBlockK, (uint32_t)kBlockM, 1, 1};
    if (!make_tmap_bf16(&L->tmA, x, 4, dims, box)) return false;
  }
  L->cg = variant < 0 ? conv_pick_cg(p.num_m_tiles) : ((variant & 1) ? 2 : 1);
  {
    const uint64_t dims[2] = {(uint64_t)p.taps * cin, (uint64_t)cout_pad};
    const uint32_t box[2] = {(uint32_t)kBlockK, (uint32_t)(bn / L->cg)};
    if (!make_tmap_bf16(&L->tmB, w_packed, 2, dims, box)) return false;
  }
  conv_set_grid(L, num_sms);
  p.ldc = cout_pad;
  return true;
}

// bf16 output [M, cout_pad] (modes 0..2, 5).  Mode 2 accumulates in place: `out` is also the auxiliary (residual) tile.
inline bool conv_launch_set_out(ConvLaunch* L, __nv_bfloat16* out) {
  const uint64_t dims[2] = {(uint64_t)L->cout_pad, (uint64_t)L->p.m_total};
  const uint32_t box[2] = {64u, (uint32_t)kBlockM};
  L->out_ptr = out;
  if (!make_tmap_bf16(&L->tmOut, out, 2, dims, box)) return false;
  if (L->p.mode == EPI_BIAS_RES) L->tmAux = L->tmOut;
  return true;
}
// Auxiliary input tile of mode 5 (the stashed pre-activation), same shape as the output.
inline bool conv_launch_set_aux(ConvLaunch* L, const __nv_bfloat16* aux) {
  const uint64_t dims[2] = {(uint64_t)L->cout_pad, (uint64_t)L->p.m_total};
  const uint32_t box[2] = {64u, (uint32_t)kBlockM};
  return make_tmap_bf16(&L->tmAux, aux, 2, dims, box);
}
// Second output of mode 1: the pre-activation acc + bias (stash for the VJP).
inline bool conv_launch_set_out2(ConvLaunch* L, __nv_bfloat16* out2) {
  const uint64_t dims[2] = {(uint64_t)L->cout_pad, (uint64_t)L->p.m_total};
  const uint32_t box[2] = {64u, (uint32_t)kBlockM};
  if (L->p.mode != EPI_BIAS_SILU) return false;
  L->p.dual_out = 1;
  return make_tmap_bf16(&L->tmOut2, out2, 2, dims, box);
}

// Whether this launch can also emit the channel LayerNorm of its output (one N tile = all channels of a pixel).
inline bool conv_launch_can_ln(const ConvLaunch* L, int upsample) {
  const ConvParams& p = L->p;
  if (L->ln || p.num_n_tiles != 1 || !(L->bn == 64 || L->bn == 128 || L->bn == 256)) return false;
  if (!(p.mode == EPI_BIAS || p.mode == EPI_BIAS_RES)) return false;
  if (upsample && p.taps != 9) return false;
  return true;
}

// Fused LayerNorm output: ln_out is [M, C] or, upsampled, [n_img, 2Ho, 2Wo, C]
inline bool conv_launch_set_ln(ConvLaunch* L, __nv_bfloat16* ln_out, const float* ln_mod, int upsample) {
  ConvParams& p = L->p;
  if (!conv_launch_can_ln(L, upsample)) return false;
  p.ln_out = ln_out;
  p.ln_mod = ln_mod;
  p.ln_up = upsample;
  p.ln_H = L->Ho;
  p.ln_W = L->Wo;
  L->ln = 1;
  return true;
}

template <int BN, int CG, bool LN>
inline cudaError_t conv_launch_variant(const ConvLaunch& L, cudaStream_t stream) {
  using Cfg = ConvCfg<BN, CG>;
  static bool attr_done = false;
  if (!attr_done) {
    cudaError_t e = cudaFuncSetAttribute(conv_gemm_tcgen05_kernel<BN, CG, LN>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemLimit);
    if (e != cudaSuccess) return e;
    attr_done = true;
  }
  ConvParams p = L.p;
  p.num_staging = (p.mode == EPI_BIAS_RES || p.mode == EPI_DSILU || p.dual_out) ? 2 : 1;
  p.num_stages = Cfg::stages_for(p.num_staging);
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(L.grid);
  cfg.blockDim = dim3(kConvThreads);
  cfg.dynamicSmemBytes = Cfg::smem_bytes(p.num_staging);
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CG;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, conv_gemm_tcgen05_kernel<BN, CG, LN>, L.tmA, L.tmB, L.tmOut, L.tmAux, L.tmOut2, p);
}

template <int BN, bool LN>
inline cudaError_t conv_launch_bn_ln(const ConvLaunch& L, cudaStream_t stream) {
  if (L.cg == 2) return conv_launch_variant<BN, 2, LN>(L, stream);
  return conv_launch_variant<BN, 1, LN>(L, stream);
}

template <int BN>
inline cudaError_t conv_launch_bn(const ConvLaunch& L, cudaStream_t stream) {
  if constexpr (BN != 192) {
    if (L.ln) return conv_launch_bn_ln<BN, true>(L, stream);
  }
  return conv_launch_bn_ln<BN, false>(L, stream);
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
