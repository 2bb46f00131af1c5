// EXPERIMENTAL — weight-gradient GEMM of a 3x3 stride-1 conv (the wgrad half of SURVEY 8(f) N2).
// Written at the end of round 1, compiled for sm_100a, NOT WORKING YET: its first (and, for lack of GPU budget, only)
// hardware runs ended in cudaErrorIllegalInstruction after ~4 s in every case (transpose_bf16_kernel passes its test).
// That is mbar_wait's watchdog (BPT.TRAP after 4e9 cycles), i.e. the kernel DEADLOCKS: by the SASS the first wait to
// time out is the MMA warp's on full[stage] — the stage's TMA bytes never complete.  Launch style (plain vs 1-CTA
// cluster) makes no difference.  Next step: compute-sanitizer + a single-K-block run dumping the barrier state.  No product path calls it, no parity claim covers it, and
// its GPU test is opt-in (C2W_EXPERIMENTAL=1).  Ground truth for it: tests/golden/train_step.npz.
//
//   dW[co, (r*3+s)*Cin + ci] += sum over pixels  dY[pix, co] * X[pix + (r-1, s-1), ci]        (zero padding)
//
// As a GEMM per tap: M = Cout, N = Cin, K = pixels.  The reduction dimension must be contiguous for the K-major
// shared-memory descriptors K1 uses, so the operands are the TRANSPOSED activations dY^T [Cout][pix] and
// X^T [Cin][n][H][W] (transpose_bf16_kernel; one extra pass per tensor).  A K step is a block of 64 pixels of one image
// (bw x bh, bw = min(W, 64)): A = a 2-D TMA box of dY^T, B = the same block of X^T shifted by the tap through the 4-D
// box coordinates, out-of-range pixels zero-filled by the TMA unit.  Only 9 * (Cout/128) * (Cin/BN) output tiles
// exist, so the pixel range is split over the grid (split-K): one CTA per (tap, M tile, N tile, split), fp32 partial
// sums added to dW with red.global.add.f32.  Warp 0: TMA producer, warp 1: tcgen05.mma issuer, warps 2-5: epilogue.
#pragma once
#include "conv_tcgen05.cuh"

namespace c2w {

constexpr int kWgStages = 5;
constexpr int kWgThreads = 192;

struct WgradParams {
  int n_img, H, W;
  int cin, cout;        // padded channel counts (multiples of 64)
  int bw, bh;           // pixel block of one K step (bw * bh == 64)
  int blocks_per_img;   // (H / bh) * (W / bw)
  int k_blocks;         // n_img * blocks_per_img
  int splits, m_tiles, n_tiles;
  float* dw;            // fp32 [cout][9 * cin], accumulated
  int ldw;              // 9 * cin
};

template <int BN>
__global__ void __launch_bounds__(kWgThreads, 1)
wgrad_tcgen05_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const WgradParams p) {
  constexpr int kABytes = kBlockM * kBlockK * 2;  // [128 co][64 px] bf16
  constexpr int kBBytes = BN * kBlockK * 2;       // [BN ci][64 px] bf16
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* smA = smem;
  uint8_t* smB = smem + kWgStages * kABytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smB + kWgStages * kBBytes);
  uint64_t* full = bars;
  uint64_t* empty = bars + kWgStages;
  uint64_t* acc_full = bars + 2 * kWgStages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_full + 1);

  const int warp_idx = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // work item
  int idx = blockIdx.x;
  const int split = idx % p.splits;
  idx /= p.splits;
  const int nt = idx % p.n_tiles;
  idx /= p.n_tiles;
  const int mt = idx % p.m_tiles;
  const int tap = idx / p.m_tiles;
  const int r = tap / 3, s = tap - 3 * r;
  const int kb0 = static_cast<int>(static_cast<long long>(p.k_blocks) * split / p.splits);
  const int kb1 = static_cast<int>(static_cast<long long>(p.k_blocks) * (split + 1) / p.splits);

  if (warp_idx == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int i = 0; i < kWgStages; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 1);
    }
    mbar_init(acc_full, 1);
    fence_mbar_init();
  }
  if (warp_idx == 2) tmem_alloc(tmem_slot, BN);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_d = *tmem_slot;

  if (warp_idx == 0) {
    // ---------------------------------------------------------------- TMA producer
    const int blocks_w = p.W / p.bw;
    int stage = 0;
    uint32_t phase = 0;
    for (int kb = kb0; kb < kb1; ++kb) {
      mbar_wait(&empty[stage], phase ^ 1);
      if (elect_one()) {
        const int img = kb / p.blocks_per_img, t = kb - img * p.blocks_per_img;
        const int h0 = (t / blocks_w) * p.bh, w0 = (t - (t / blocks_w) * blocks_w) * p.bw;
        const int pix0 = (img * p.H + h0) * p.W + w0;  // the block's pixels are contiguous (bw == W or bh == 1)
        mbar_arrive_expect_tx(&full[stage], kABytes + kBBytes);
        tma_load_2d(&tmA, &full[stage], smA + stage * kABytes, pix0, mt * kBlockM);
        tma_load_4d(&tmB, &full[stage], smB + stage * kBBytes, w0 + s - 1, h0 + r - 1, img, nt * BN);
      }
      __syncwarp();
      if (++stage == kWgStages) {
        stage = 0;
        phase ^= 1;
      }
    }
  } else if (warp_idx == 1) {
    // ---------------------------------------------------------------- MMA issuer
    constexpr uint32_t idesc = umma_idesc_bf16(kBlockM, BN);
    const uint64_t adesc0 = umma_desc_kmajor_sw128(smem_u32(smA));
    const uint64_t bdesc0 = umma_desc_kmajor_sw128(smem_u32(smB));
    int stage = 0;
    uint32_t phase = 0;
    for (int kb = kb0; kb < kb1; ++kb) {
      mbar_wait(&full[stage], phase);
      tc_fence_after();
      if (elect_one()) {
        const uint64_t adesc = adesc0 + static_cast<uint64_t>(stage * (kABytes >> 4));
        const uint64_t bdesc = bdesc0 + static_cast<uint64_t>(stage * (kBBytes >> 4));
#pragma unroll
        for (int k = 0; k < kBlockK / 16; ++k) umma_bf16(tmem_d, adesc + 2 * k, bdesc + 2 * k, idesc, (kb > kb0) || k != 0);
        umma_commit(&empty[stage]);
        if (kb + 1 == kb1) umma_commit(acc_full);
      }
      __syncwarp();
      if (++stage == kWgStages) {
        stage = 0;
        phase ^= 1;
      }
    }
  } else if (kb1 > kb0) {
    // ---------------------------------------------------------------- epilogue: TMEM -> red.global.add.f32
    const int q = warp_idx & 3;  // TMEM lane quadrant this warp may read
    const int row = q * 32 + lane;
    const int co = mt * kBlockM + row;
    mbar_wait(acc_full, 0);
    tc_fence_after();
    float* dst = p.dw + static_cast<long long>(co) * p.ldw + tap * p.cin + nt * BN;
#pragma unroll 1
    for (int c = 0; c < BN / 32; ++c) {
      uint32_t v[32];
      tmem_ld_32x32(tmem_d + (static_cast<uint32_t>(q * 32) << 16) + c * 32, v);
      tmem_ld_wait();
      if (co < p.cout) {
#pragma unroll
        for (int j = 0; j < 32; ++j) atomicAdd(dst + c * 32 + j, __uint_as_float(v[j]));
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp_idx == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_d, BN);
  }
}

// bf16 [rows][cols] -> [cols][rows] (activations [pix][C] -> [C][pix]); 64 x 64 tiles through shared memory
__global__ void transpose_bf16_kernel(const __nv_bfloat16* __restrict__ in, __nv_bfloat16* __restrict__ out,
                                      long long rows, int cols) {
  __shared__ __nv_bfloat16 tile[64][66];
  const long long r0 = static_cast<long long>(blockIdx.x) * 64;
  const int c0 = blockIdx.y * 64;
  for (int i = threadIdx.y; i < 64; i += blockDim.y) {
    const long long rr = r0 + i;
    const int cc = c0 + threadIdx.x;
    for (int h = 0; h < 2; ++h) {
      const int c = cc + 32 * h;
      tile[i][threadIdx.x + 32 * h] = (rr < rows && c < cols) ? in[rr * cols + c] : __float2bfloat16(0.f);
    }
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 64; i += blockDim.y) {
    const int c = c0 + i;
    for (int h = 0; h < 2; ++h) {
      const long long rr = r0 + threadIdx.x + 32 * h;
      if (c < cols && rr < rows) out[static_cast<long long>(c) * rows + rr] = tile[threadIdx.x + 32 * h][i];
    }
  }
}

// Host side: xt = X^T [cin][n*H*W], dyt = dY^T [cout][n*H*W] (bf16), dw fp32 [cout][9*cin] (+=)
template <int BN>
inline cudaError_t wgrad_launch_bn(const CUtensorMap& tmA, const CUtensorMap& tmB, const WgradParams& p, int grid,
                                   cudaStream_t stream) {
  constexpr int smem = kWgStages * (kBlockM * kBlockK * 2 + BN * kBlockK * 2) + 256 + 1024;
  static bool attr_done = false;
  if (!attr_done) {
    cudaError_t e = cudaFuncSetAttribute(wgrad_tcgen05_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return e;
    attr_done = true;
  }
  // launched like K1 (explicit 1-CTA cluster): the tcgen05.commit / TMA forms address shared::cluster
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(kWgThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 1;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, wgrad_tcgen05_kernel<BN>, tmA, tmB, p);
}

inline int wgrad_launch(const __nv_bfloat16* xt, const __nv_bfloat16* dyt, int n_img, int H, int W, int cin, int cout,
                        float* dw, int num_sms, cudaStream_t stream, char* err, int err_len) {
  auto bad = [&](const char* m) {
    snprintf(err, err_len, "%s", m);
    return -1;
  };
  if (cin % 64 != 0 || cout % 64 != 0) return bad("wgrad: channel counts must be multiples of 64");
  if (W < 8 || W > 128 || (W & (W - 1)) != 0) return bad("wgrad: W must be a power of two in 8..128");
  WgradParams p;
  memset(&p, 0, sizeof(p));
  p.n_img = n_img;
  p.H = H;
  p.W = W;
  p.cin = cin;
  p.cout = cout;
  p.bw = W < 64 ? W : 64;
  p.bh = 64 / p.bw;
  if (H % p.bh != 0) return bad("wgrad: H must be a multiple of 64 / min(W, 64)");
  p.blocks_per_img = (H / p.bh) * (W / p.bw);
  p.k_blocks = n_img * p.blocks_per_img;
  const int bn = (cin % 128 == 0) ? 128 : 64;
  p.m_tiles = (cout + kBlockM - 1) / kBlockM;
  p.n_tiles = cin / bn;
  const int tiles = 9 * p.m_tiles * p.n_tiles;
  int splits = num_sms / tiles;
  if (splits < 1) splits = 1;
  if (splits > p.k_blocks) splits = p.k_blocks;
  p.splits = splits;
  p.dw = dw;
  p.ldw = 9 * cin;
  const long long pix = static_cast<long long>(n_img) * H * W;
  CUtensorMap tmA, tmB;
  {
    const uint64_t dims[2] = {(uint64_t)pix, (uint64_t)cout};
    const uint32_t box[2] = {(uint32_t)kBlockK, (uint32_t)kBlockM};
    if (!make_tmap_bf16(&tmA, dyt, 2, dims, box)) return bad(tmap_error_slot());
  }
  {
    const uint64_t dims[4] = {(uint64_t)W, (uint64_t)H, (uint64_t)n_img, (uint64_t)cin};
    const uint32_t box[4] = {(uint32_t)p.bw, (uint32_t)p.bh, 1u, (uint32_t)bn};
    if (!make_tmap_bf16(&tmB, xt, 4, dims, box)) return bad(tmap_error_slot());
  }
  const int grid = tiles * splits;
  cudaError_t e = bn == 128 ? wgrad_launch_bn<128>(tmA, tmB, p, grid, stream) : wgrad_launch_bn<64>(tmA, tmB, p, grid, stream);
  if (e != cudaSuccess) return bad(cudaGetErrorString(e));
  return 0;
}

}  // namespace c2w
