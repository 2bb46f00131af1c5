// K10 — weight gradient of the UNet convolutions on tcgen05 tensor cores (the wgrad half of the training step,
// training_loop.py:372-378: `fabric.backward(loss)` through every Conv2d / Conv1d of model/nn.py).
//
//   dW[co, ci, r, s] = sum over output pixels p   dY[p, co] * X[stride * p + (r - 1, s - 1), ci]        (zero padding)
//
// A GEMM per filter tap with M = Cout, N = Cin and K = PIXELS.  Both operands are stored pixel-major (NHWC bf16), i.e.
// MN-major for this product: a TMA box {64 channels, 8 x 8 pixels} lands in shared memory as 64 rows (one per pixel =
// one K index) of 128 B (64 channels) in the 128B-swizzle pattern, which is exactly tcgen05's canonical MN-major
// SWIZZLE_128B operand (cute: Layout_MN_SW128_Atom; descriptor: LBO = distance between 64-channel groups, SBO = 1024 B
// between 8-pixel groups; instruction descriptor with a_major = b_major = MN).  No transposed copies of the activations.
//
// K blocks are 8 x 8 spatial blocks of one image.  For stride-1 3x3 convs the three taps of one filter COLUMN s share
// one load of X: the block with a one-row halo above and below ([10 x 8 px][64 ch], shifted by s - 1 columns through
// the box coordinates, out-of-image pixels zero-filled by the TMA unit = the conv's zero padding); filter row r is the
// sub-view starting r * 8 pixels = r KB into it (swizzle phase preserved).  Stride-2 heads and the 1x1 GEMMs load one
// box per tap.  A CTA owns (filter column | tap group, 128 output channels, BN input channels, one slice of the pixel
// range): only (Cout/128) * (Cin/BN) * 3 output tiles exist per layer, so the pixel range is split over the grid
// (split-K) and every CTA writes its three fp32 accumulators to a scratch slab; wgrad_reduce_kernel sums the slabs in
// a fixed order (deterministic) into the fp32 OIHW gradient tensor.
//
// The conv's BIAS gradient, sum over pixels of dY[p, co], rides along on the tensor core: it is the same product with a
// column of ones for X.  The CTAs of N tile 0 issue, for every `groups`-th K block (so the groups of a split share the
// work), four extra 128 x 16 x 16 MMAs of the dY tile against a constant all-ones operand in shared memory (any layout
// of ones is ones, so the tile needs no swizzle care) into 16 spare TMEM columns; the epilogue writes column 0.  No
// second pass over dY and no LSU traffic (a first version summed the staged tiles with the idle epilogue warps: their
// shared-memory reads held every stage and made the bias CTAs 1.4x slower than the rest of the grid).
//
// Roles: warp 0 TMA producer, warp 1 tcgen05.mma issuer (one elected lane), warp 2 TMEM allocator, warps 4-7 epilogue.
#pragma once
#include "conv_tcgen05.cuh"

namespace c2w {

constexpr int kWgThreads = 256;
constexpr int kWgPix = 64;  // pixels (K) per block: 8 x 8

struct WgradParams {
  int taps;            // 9 (3x3, pad 1) or 1 (GEMM)
  int stride;          // 1 or 2 (3x3 only)
  int share;           // 1: halo'd unit shared by the 3 filter rows (stride-1 3x3)
  int blocks_w, blocks_per_img, k_blocks;  // 8x8 output-pixel blocks
  int splits, m_tiles, n_tiles;
  int groups;          // work items per (m, n, split): 3 (filter columns, or filter rows when !share) or 1 (GEMM)
  int cout_slab, cin_slab;  // slab dims: m_tiles * 128, n_tiles * BN
  float* partial;      // [splits][taps][cout_slab][cin_slab]
  float* bias_partial; // [splits * groups][cout_slab]: per-CTA column sums of dY (the bias gradient), or null
  int num_stages;
};

// Shared-memory descriptor of an MN-major, 128B-swizzled operand: rows of 128 B (64 channels of one pixel), 8-pixel
// groups 1024 B apart (SBO), the next 64-channel group `lbo_bytes` further on (LBO).
__device__ __forceinline__ uint64_t umma_desc_mnmajor_sw128(uint32_t smem_addr, uint32_t lbo_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr >> 4) & 0x3FFF);
  d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= static_cast<uint64_t>(1024 >> 4) << 32;
  d |= static_cast<uint64_t>(1) << 46;  // descriptor version (sm_100)
  d |= static_cast<uint64_t>(2) << 61;  // SWIZZLE_128B
  return d;
}
// bf16 x bf16 -> fp32, A and B both MN-major (bits 15 / 16)
__host__ __device__ constexpr uint32_t umma_idesc_bf16_mn(int M, int N) {
  return (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | (static_cast<uint32_t>(N >> 3) << 17) |
         (static_cast<uint32_t>(M >> 4) << 24);
}

template <int BN>
struct WgradCfg {
  static constexpr int kABytes = 2 * kWgPix * 128;                 // dY: two 64-channel groups of 64 pixels
  static constexpr int kBGroupShare = 80 * 128;                    // X unit: 10 x 8 pixels per 64-channel group
  static constexpr int kBGroupTap = kWgPix * 128;                  // X box of one tap per 64-channel group
  static constexpr int kBBytesShare = (BN / 64) * kBGroupShare;
  static constexpr int kBBytesTaps = 3 * (BN / 64) * kBGroupTap;
  static constexpr int kTmemCols = (3 * BN + 16 <= 256) ? 256 : 512;  // three accumulators + 16 bias columns
  static constexpr int kOnesBytes = 2048;                          // all-ones B operand of the bias MMAs (16 pixels x 128 B)
  static constexpr int stage_bytes(bool share) {
    // stages stay 1024-B aligned (the swizzle pattern is a function of the absolute address)
    return ((kABytes + (share ? kBBytesShare : kBBytesTaps)) + 1023) / 1024 * 1024;
  }
  static constexpr int kBarBytes = 512;
  static constexpr int stages(bool share) {
    const int n = (kSmemLimit - 1024 - kBarBytes - kOnesBytes) / stage_bytes(share);
    return n > 8 ? 8 : n;
  }
  static constexpr int smem_bytes(bool share) { return stages(share) * stage_bytes(share) + kOnesBytes + kBarBytes + 1024; }
};

template <int BN>
__global__ void __launch_bounds__(kWgThreads, 1)
wgrad_tcgen05_kernel(const __grid_constant__ CUtensorMap tmDY, const __grid_constant__ CUtensorMap tmX, const WgradParams p) {
  using Cfg = WgradCfg<BN>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const bool share = p.share != 0;
  const int stage_bytes = Cfg::stage_bytes(share);
  const int num_stages = p.num_stages;
  uint8_t* ones = smem + num_stages * stage_bytes;  // 1024-B aligned
  uint64_t* bars = reinterpret_cast<uint64_t*>(ones + Cfg::kOnesBytes);
  uint64_t* full = bars;
  uint64_t* empty = bars + 8;
  uint64_t* acc_full = bars + 16;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 17);

  const int warp_idx = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // work item: (group g, m tile, n tile, split)
  int idx = blockIdx.x;
  const int split = idx % p.splits;
  idx /= p.splits;
  const int nt = idx % p.n_tiles;
  idx /= p.n_tiles;
  const int mt = idx % p.m_tiles;
  const int g = idx / p.m_tiles;
  const int ntap = p.taps == 1 ? 1 : 3;  // accumulators of this CTA
  const int kb0 = static_cast<int>(static_cast<long long>(p.k_blocks) * split / p.splits);
  const int kb1 = static_cast<int>(static_cast<long long>(p.k_blocks) * (split + 1) / p.splits);
  // bias gradient: this CTA takes the K blocks kb with kb % groups == g of its split
  const bool do_bias = p.bias_partial != nullptr && nt == 0;
  const int kb_bias0 = kb0 + ((g - kb0 % p.groups) + p.groups) % p.groups;  // first of them

  if (warp_idx == 0 && lane == 0) {
    tma_prefetch_desc(&tmDY);
    tma_prefetch_desc(&tmX);
  }
  if (warp_idx == 1 && lane == 0) {
    for (int i = 0; i < num_stages; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 1);
    }
    mbar_init(acc_full, 1);
    fence_mbar_init();
  }
  if (warp_idx == 2) tmem_alloc(tmem_slot, Cfg::kTmemCols);
  if (do_bias && warp_idx >= 4) {  // 128 threads fill the ones tile; the async proxy (tcgen05) reads it
    uint4* o = reinterpret_cast<uint4*>(ones);
    const uint4 one8 = make_uint4(0x3F803F80u, 0x3F803F80u, 0x3F803F80u, 0x3F803F80u);
    o[threadIdx.x - 128] = one8;
    fence_proxy_async();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_d = *tmem_slot;

  if (warp_idx == 0) {
    // ---------------------------------------------------------------- TMA producer
    const int tx_bytes = Cfg::kABytes + (share ? Cfg::kBBytesShare : ntap * (BN / 64) * Cfg::kBGroupTap);
    int stage = 0;
    uint32_t phase = 0;
    for (int kb = kb0; kb < kb1; ++kb) {
      mbar_wait(&empty[stage], phase ^ 1);
      if (elect_one()) {
        const int img = kb / p.blocks_per_img, t = kb - img * p.blocks_per_img;
        const int h0 = (t / p.blocks_w) * 8, w0 = (t % p.blocks_w) * 8;  // output-pixel block
        uint8_t* sA = smem + stage * stage_bytes;
        uint8_t* sB = sA + Cfg::kABytes;
        mbar_arrive_expect_tx(&full[stage], tx_bytes);
        for (int b = 0; b < 2; ++b)  // channels beyond the tensor (Cout = 64) are zero-filled by the TMA unit
          tma_load_4d(&tmDY, &full[stage], sA + b * (kWgPix * 128), mt * 128 + b * 64, w0, h0, img);
        if (share) {  // g = filter column s: one halo'd unit per 64-channel group, rows h0-1 .. h0+8
          for (int b = 0; b < BN / 64; ++b)
            tma_load_4d(&tmX, &full[stage], sB + b * Cfg::kBGroupShare, nt * BN + b * 64, w0 + g - 1, h0 - 1, img);
        } else if (p.taps == 9) {  // g = filter row r: one box per filter column s (stride 1 or 2)
          for (int s = 0; s < 3; ++s)
            for (int b = 0; b < BN / 64; ++b)
              tma_load_4d(&tmX, &full[stage], sB + (s * (BN / 64) + b) * Cfg::kBGroupTap, nt * BN + b * 64,
                          p.stride * w0 + s - 1, p.stride * h0 + g - 1, img);
        } else {  // GEMM: the same 64 rows of X
          for (int b = 0; b < BN / 64; ++b)
            tma_load_4d(&tmX, &full[stage], sB + b * Cfg::kBGroupTap, nt * BN + b * 64, w0, h0, img);
        }
      }
      __syncwarp();
      if (++stage == num_stages) {
        stage = 0;
        phase ^= 1;
      }
    }
  } else if (warp_idx == 1) {
    // ---------------------------------------------------------------- MMA issuer
    constexpr uint32_t idesc = umma_idesc_bf16_mn(128, BN);
    constexpr uint32_t idesc_bias = umma_idesc_bf16_mn(128, 16);
    const uint64_t odesc = umma_desc_mnmajor_sw128(smem_u32(ones), 1024);
    int stage = 0;
    uint32_t phase = 0;
    for (int kb = kb0; kb < kb1; ++kb) {
      mbar_wait(&full[stage], phase);
      tc_fence_after();
      if (elect_one()) {
        const uint32_t sA = smem_u32(smem + stage * stage_bytes);
        const uint32_t sB = sA + Cfg::kABytes;
        const uint64_t adesc = umma_desc_mnmajor_sw128(sA, kWgPix * 128);
        for (int a = 0; a < ntap; ++a) {
          // accumulator a: share -> filter row r = a (sub-view a * 8 pixels into the unit); else filter column s = a
          const uint64_t bdesc = share ? umma_desc_mnmajor_sw128(sB + a * 1024, Cfg::kBGroupShare)
                                       : umma_desc_mnmajor_sw128(sB + a * (BN / 64) * Cfg::kBGroupTap, Cfg::kBGroupTap);
#pragma unroll
          for (int k = 0; k < kWgPix / 16; ++k)  // 16 pixels = two 8-pixel groups = 2048 B further on
            umma_bf16(tmem_d + a * BN, adesc + k * (2048 >> 4), bdesc + k * (2048 >> 4), idesc, (kb > kb0) || k != 0);
        }
        if (do_bias && (kb - kb_bias0) % p.groups == 0 && kb >= kb_bias0) {
#pragma unroll
          for (int k = 0; k < kWgPix / 16; ++k)
            umma_bf16(tmem_d + ntap * BN, adesc + k * (2048 >> 4), odesc, idesc_bias, (kb > kb_bias0) || k != 0);
        }
        umma_commit(&empty[stage]);
        if (kb + 1 == kb1) umma_commit(acc_full);
      }
      __syncwarp();
      if (++stage == num_stages) {
        stage = 0;
        phase ^= 1;
      }
    }
  } else if (warp_idx >= 4) {
    // ---------------------------------------------------------------- epilogue: TMEM -> fp32 slab
    const int q = warp_idx & 3;  // TMEM lane quarter this warp may read
    const int row = q * 32 + lane;
    const bool any = kb1 > kb0;
    if (any) {
      mbar_wait(acc_full, 0);
      tc_fence_after();
    }
    if (do_bias) {
      float bsum = 0.f;
      if (kb_bias0 < kb1) {  // else: no K block of this split was ours
        uint32_t v;
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];"
                     : "=r"(v)
                     : "r"(tmem_d + (static_cast<uint32_t>(q * 32) << 16) + ntap * BN)
                     : "memory");
        tmem_ld_wait();
        bsum = __uint_as_float(v);
      }
      p.bias_partial[(static_cast<size_t>(split) * p.groups + g) * p.cout_slab + mt * 128 + row] = bsum;
    }
    for (int a = 0; a < ntap; ++a) {
      const int tap = p.taps == 1 ? 0 : (share ? a * 3 + g : g * 3 + a);  // tap = r * 3 + s
      float* dst = p.partial + ((static_cast<size_t>(split) * p.taps + tap) * p.cout_slab + mt * 128 + row) * p.cin_slab +
                   nt * BN;
#pragma unroll 1
      for (int c = 0; c < BN / 32; ++c) {
        uint32_t v[32];
        if (any) {
          tmem_ld_32x32(tmem_d + (static_cast<uint32_t>(q * 32) << 16) + a * BN + c * 32, v);
          tmem_ld_wait();
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = 0u;
        }
        float4* d4 = reinterpret_cast<float4*>(dst + c * 32);
#pragma unroll
        for (int j = 0; j < 8; ++j)
          d4[j] = make_float4(__uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1]), __uint_as_float(v[4 * j + 2]),
                              __uint_as_float(v[4 * j + 3]));
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp_idx == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_d, Cfg::kTmemCols);
  }
}

// dw[co][ci][tap] (fp32 OIHW / OI1, real Cout x Cin) (+)= sum over splits of partial[split][tap][co][ci]
// One thread per (tap, co, ci): consecutive threads read consecutive ci of one slab row (coalesced), the 36-byte-strided
// writes into the OIHW tensor are the small side.  Threads of tap 0 with ci == 0 also finish the bias gradient.
__global__ void wgrad_reduce_kernel(const float* __restrict__ partial, float* __restrict__ dw, int splits, int taps,
                                    int cout, int cin, int cout_slab, int cin_slab, int accumulate, float scale,
                                    const float* __restrict__ bias_partial, float* __restrict__ db, int bias_parts) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const long long per_tap = static_cast<long long>(cout) * cin;
  if (i >= per_tap * taps) return;
  const int t = static_cast<int>(i / per_tap);
  const long long r = i - t * per_tap;
  const int co = static_cast<int>(r / cin), ci = static_cast<int>(r - static_cast<long long>(co) * cin);
  const size_t slab = static_cast<size_t>(cout_slab) * cin_slab;
  const float* src = partial + static_cast<size_t>(t) * slab + static_cast<size_t>(co) * cin_slab + ci;
  // four independent partial sums (fixed order: deterministic) keep four loads in flight per thread
  float a4[4] = {0.f, 0.f, 0.f, 0.f};
  const size_t sstep = static_cast<size_t>(taps) * slab;
  int s = 0;
  for (; s + 4 <= splits; s += 4) {
#pragma unroll
    for (int j = 0; j < 4; ++j) a4[j] += src[static_cast<size_t>(s + j) * sstep];
  }
  for (; s < splits; ++s) a4[0] += src[static_cast<size_t>(s) * sstep];
  const float acc = (a4[0] + a4[1]) + (a4[2] + a4[3]);
  float* d = dw + (static_cast<size_t>(co) * cin + ci) * taps + t;
  *d = accumulate ? *d + acc * scale : acc * scale;
  if (db != nullptr && t == 0 && ci == 0) {  // bias gradient: sum of the per-split column sums, fixed order
    float b = 0.f;
    for (int s = 0; s < bias_parts; ++s) b += bias_partial[static_cast<size_t>(s) * cout_slab + co];
    db[co] = accumulate ? db[co] + b * scale : b * scale;
  }
}

// Column sums of a bf16 [rows, C] matrix per group of `rows_per_group` consecutive rows (bias gradients: one group;
// per-sample modulation gradients: one group per image):  out[group][c] (+)= scale * sum_rows x[row][c]
// grid = (row slabs, C / 64 * ... ) — each CTA sums kColsumRows rows for 256 channel lanes and adds atomically.
constexpr int kColsumRows = 128;
__global__ void colsum_bf16_kernel(const __nv_bfloat16* __restrict__ x, float* __restrict__ out, long long rows, int C,
                                   long long rows_per_group, int out_stride, float scale, int c_valid) {
  // thread -> (channel pair cp, row lane rl): blockIdx.y selects a group of up to 256 channel pairs, the remaining
  // threads of the CTA take rows in parallel
  const int pairs = C / 2;
  const int pairs_cta = pairs < static_cast<int>(blockDim.x) ? pairs : static_cast<int>(blockDim.x);
  const int rl_n = blockDim.x / pairs_cta;
  const int cp = blockIdx.y * pairs_cta + threadIdx.x % pairs_cta, rl = threadIdx.x / pairs_cta;
  if (rl >= rl_n || cp >= pairs) return;
  const long long r0 = static_cast<long long>(blockIdx.x) * kColsumRows;
  const long long r1 = (r0 + kColsumRows < rows) ? r0 + kColsumRows : rows;
  float a0 = 0.f, a1 = 0.f;
  long long grp = r0 / rows_per_group;
  for (long long r = r0 + rl; r < r1; r += rl_n) {
    const long long gnow = r / rows_per_group;
    if (gnow != grp) {  // a slab may straddle a group boundary
      if (2 * cp < c_valid) atomicAdd(out + grp * out_stride + 2 * cp, a0 * scale);
      if (2 * cp + 1 < c_valid) atomicAdd(out + grp * out_stride + 2 * cp + 1, a1 * scale);
      a0 = a1 = 0.f;
      grp = gnow;
    }
    const __nv_bfloat162 v = *reinterpret_cast<const __nv_bfloat162*>(x + r * C + 2 * cp);
    a0 += __low2float(v);
    a1 += __high2float(v);
  }
  if (2 * cp < c_valid) atomicAdd(out + grp * out_stride + 2 * cp, a0 * scale);  // channels >= c_valid are padding
  if (2 * cp + 1 < c_valid) atomicAdd(out + grp * out_stride + 2 * cp + 1, a1 * scale);
}

// ------------------------------------------------------------------------------------------------ host side
struct WgradLaunch {
  CUtensorMap tmDY, tmX;
  WgradParams p;
  int bn, grid;
  size_t scratch_floats;
};

// bf16 NHWC tensor map with a per-dimension box / element stride, 128B swizzle (rank 4: C, W, H, N)
inline bool wgrad_make_map(CUtensorMap* m, const void* base, int C, int W, int H, int N, int bw, int bh, int estride) {
  const uint64_t dims[4] = {(uint64_t)C, (uint64_t)W, (uint64_t)H, (uint64_t)N};
  const uint32_t box[4] = {64u, (uint32_t)bw, (uint32_t)bh, 1u};
  const uint32_t est[4] = {1u, (uint32_t)estride, (uint32_t)estride, 1u};
  return make_tmap_bf16(m, base, 4, dims, box, est);
}

// x: bf16 NHWC [n, H, W, cin_pad] (conv input; GEMM: H = 1, W = rows); dy: bf16 NHWC [n, H/stride, W/stride, cout_pad].
// Output-pixel blocks are 8 x 8 (GEMM: 64 consecutive rows).
inline bool wgrad_launch_init(WgradLaunch* L, bool conv3x3, const __nv_bfloat16* x, const __nv_bfloat16* dy, int n, int H,
                              int W, int cin_pad, int cout_pad, int stride, int num_sms, float* scratch,
                              size_t scratch_floats) {
  WgradParams& p = L->p;
  memset(&p, 0, sizeof(p));
  if (cin_pad % 64 != 0 || cout_pad % 64 != 0) return false;
  const int bn = (cin_pad % 128 == 0) ? 128 : 64;
  L->bn = bn;
  p.taps = conv3x3 ? 9 : 1;
  p.stride = conv3x3 ? stride : 1;
  p.m_tiles = (cout_pad + 127) / 128;
  p.n_tiles = cin_pad / bn;
  p.cout_slab = p.m_tiles * 128;
  p.cin_slab = cin_pad;
  long long k_blocks;
  if (conv3x3) {
    const int Ho = H / stride, Wo = W / stride;
    if (Ho % 8 != 0 || Wo % 8 != 0 || (stride != 1 && stride != 2)) return false;
    p.share = stride == 1;
    p.groups = 3;
    p.blocks_w = Wo / 8;
    p.blocks_per_img = (Ho / 8) * p.blocks_w;
    k_blocks = static_cast<long long>(n) * p.blocks_per_img;
    if (!wgrad_make_map(&L->tmDY, dy, cout_pad, Wo, Ho, n, 8, 8, 1)) return false;
    if (p.share) {
      if (!wgrad_make_map(&L->tmX, x, cin_pad, W, H, n, 8, 10, 1)) return false;
    } else {
      if (!wgrad_make_map(&L->tmX, x, cin_pad, W, H, n, 8 * stride, 8 * stride, stride)) return false;
    }
  } else {
    const long long rows = static_cast<long long>(n) * H * W;
    if (rows % 64 != 0) return false;
    p.share = 0;
    p.groups = 1;
    p.blocks_w = 1;
    p.blocks_per_img = static_cast<int>(rows / 64);
    k_blocks = rows / 64;
    // rows as the H dimension of a [rows / 8][8] image: a block of 64 rows is an 8 x 8 box
    if (!wgrad_make_map(&L->tmDY, dy, cout_pad, 8, static_cast<int>(rows / 8), 1, 8, 8, 1)) return false;
    if (!wgrad_make_map(&L->tmX, x, cin_pad, 8, static_cast<int>(rows / 8), 1, 8, 8, 1)) return false;
  }
  p.k_blocks = static_cast<int>(k_blocks);
  const int tiles = p.groups * p.m_tiles * p.n_tiles;
  long long splits = num_sms / tiles;
  if (splits < 1) splits = 1;
  if (splits > k_blocks) splits = k_blocks;
  const size_t per_split = static_cast<size_t>(p.taps) * p.cout_slab * p.cin_slab;
  const size_t bias_floats = static_cast<size_t>(p.cout_slab) * p.groups;
  while (splits > 1 && (per_split + bias_floats) * splits > scratch_floats) --splits;
  if ((per_split + bias_floats) * splits > scratch_floats) return false;
  p.splits = static_cast<int>(splits);
  p.partial = scratch;
  p.bias_partial = nullptr;  // wgrad_run binds it behind the slabs when a bias gradient is asked for
  L->scratch_floats = (per_split + bias_floats) * splits;
  L->grid = tiles * p.splits;
  p.num_stages = bn == 128 ? WgradCfg<128>::stages(p.share != 0) : WgradCfg<64>::stages(p.share != 0);
  return true;
}

template <int BN>
inline cudaError_t wgrad_launch_bn(const WgradLaunch& L, cudaStream_t stream) {
  using Cfg = WgradCfg<BN>;
  static unsigned long long attr_done = 0;
  int dev = 0;
  cudaGetDevice(&dev);
  if (!((attr_done >> (dev & 63)) & 1ull)) {
    cudaError_t e = cudaFuncSetAttribute(wgrad_tcgen05_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemLimit);
    if (e != cudaSuccess) return e;
    attr_done |= 1ull << (dev & 63);
  }
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(L.grid);
  cfg.blockDim = dim3(kWgThreads);
  cfg.dynamicSmemBytes = Cfg::smem_bytes(L.p.share != 0);
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;  // explicit 1-CTA cluster: the TMA / commit forms address shared::cluster
  attr[0].val.clusterDim.x = 1;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, wgrad_tcgen05_kernel<BN>, L.tmDY, L.tmX, L.p);
}

// GEMM + slab reduction into dw (fp32 [cout][cin][taps], real channel counts) and, if db != null, the bias gradient
// db fp32 [cout] from the column sums the GEMM collected on the way.
inline cudaError_t wgrad_run(const WgradLaunch& L0, float* dw, int cout, int cin, int accumulate, float scale,
                             cudaStream_t stream, float* db = nullptr) {
  WgradLaunch L = L0;
  const size_t slab_floats = static_cast<size_t>(L.p.splits) * L.p.taps * L.p.cout_slab * L.p.cin_slab;
  L.p.bias_partial = db ? L.p.partial + slab_floats : nullptr;
  cudaError_t e = L.bn == 128 ? wgrad_launch_bn<128>(L, stream) : wgrad_launch_bn<64>(L, stream);
  if (e != cudaSuccess) return e;
  const long long items = static_cast<long long>(cout) * cin * L.p.taps;
  wgrad_reduce_kernel<<<static_cast<int>((items + 255) / 256), 256, 0, stream>>>(
      L.p.partial, dw, L.p.splits, L.p.taps, cout, cin, L.p.cout_slab, L.p.cin_slab, accumulate, scale, L.p.bias_partial, db,
      L.p.splits * L.p.groups);
  return cudaGetLastError();
}

inline cudaError_t colsum_run(const __nv_bfloat16* x, float* out, long long rows, int C, long long rows_per_group,
                              int out_stride, float scale, cudaStream_t stream, int c_valid = -1) {
  const int threads = 256;
  const long long blocks = (rows + kColsumRows - 1) / kColsumRows;
  const int pairs = C / 2;
  const dim3 grid(static_cast<unsigned>(blocks), static_cast<unsigned>((pairs + threads - 1) / threads));
  colsum_bf16_kernel<<<grid, threads, 0, stream>>>(x, out, rows, C, rows_per_group, out_stride, scale,
                                                                        c_valid < 0 ? C : c_valid);
  return cudaGetLastError();
}

}  // namespace c2w
