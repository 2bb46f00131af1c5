// K8 — time-axis halo exchange over NVLink peer memory (new in this build: the reference never shards a
// trajectory).  Frame i's score needs frames i-k .. i+k (src/thor/score.py:68-93), so after every state update each rank
// hands its k boundary frames to each neighbour.  One process per GPU: every rank cudaMalloc's a MAILBOX
//   [2 parities][2 sides][k frames]  +  one step counter per side
// exports it as a CUDA IPC handle (the host exchanges the 64-byte handles by any means — torch.distributed here) and maps
// its neighbours' mailboxes.  c2w_halo_exchange then is two small kernels on the caller's stream, no NCCL launch, no host
// synchronisation:
//   push: my first / last k owned frames -> the left / right neighbour's mailbox (plain 16-byte stores through NVLink),
//         __threadfence_system, and the last CTA publishes step+1 in the neighbour's counter (st.release.sys);
//   pull: wait until both of MY counters reached step+1 (ld.acquire.sys, with a watchdog), copy my mailbox into the halo
//         frames of x.
// Mailboxes are double-buffered by step parity: a neighbour can only be one exchange ahead (its next push needs my
// current one), so the slot it writes is never the one I am still reading.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "c2w_b200.h"
#include "common.cuh"

using namespace c2w;

void c2w_count_launches(int n);  // engine.cu: bench.py's gpu_launches

struct c2w_halo {
  int64_t halo_bytes = 0;     // k frames
  uint8_t* mailbox = nullptr; // [2][2][halo_bytes] then 2 x uint32 counters (256-byte aligned block)
  uint32_t* flags = nullptr;  // my counters: [0] written by the left neighbour, [1] by the right one
  uint32_t* done = nullptr;   // push-kernel CTA counter (local)
  uint8_t* peer_box[2] = {nullptr, nullptr};   // left / right neighbour's mailbox (mapped)
  uint32_t* peer_flags[2] = {nullptr, nullptr};
  void* peer_base[2] = {nullptr, nullptr};
  uint32_t step = 0;
  int64_t box_bytes() const { return 4 * halo_bytes; }
};

namespace {

__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

// src_l / src_r: my first / last k owned frames; dst_l / dst_r: slots in the left / right neighbour's mailbox (or null)
__global__ void halo_push_kernel(const float4* __restrict__ src_l, float4* dst_l, uint32_t* flag_l,
                                 const float4* __restrict__ src_r, float4* dst_r, uint32_t* flag_r, long long n16,
                                 uint32_t publish, uint32_t* done) {
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n16; i += stride) {
    if (dst_l) dst_l[i] = src_l[i];
    if (dst_r) dst_r[i] = src_r[i];
  }
  __threadfence_system();  // this thread's peer stores are visible system-wide before the CTA counts itself done
  __syncthreads();
  if (threadIdx.x == 0) {
    const uint32_t prev = atomicAdd(done, 1u);
    if (prev == gridDim.x - 1) {  // last CTA: every CTA's stores are fenced -> publish the step to the neighbours
      *done = 0;
      __threadfence_system();
      if (flag_l) st_release_sys(flag_l, publish);
      if (flag_r) st_release_sys(flag_r, publish);
    }
  }
}

__global__ void halo_pull_kernel(const float4* __restrict__ box_l, float4* halo_l, const uint32_t* flag_l,
                                 const float4* __restrict__ box_r, float4* halo_r, const uint32_t* flag_r, long long n16,
                                 uint32_t expect) {
  if (threadIdx.x == 0) {
    const long long t0 = clock64();
    while ((flag_l && static_cast<int32_t>(ld_acquire_sys(flag_l) - expect) < 0) ||
           (flag_r && static_cast<int32_t>(ld_acquire_sys(flag_r) - expect) < 0)) {
      if (clock64() - t0 > 20000000000ll) {  // ~10 s: a neighbour died — fail the launch instead of hanging the GPU
        printf("c2w: halo exchange watchdog (step %u)\n", expect);
        __trap();
      }
    }
  }
  __syncthreads();
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n16; i += stride) {
    if (halo_l) halo_l[i] = box_l[i];
    if (halo_r) halo_r[i] = box_r[i];
  }
}

}  // namespace

extern "C" {

int c2w_halo_create(int64_t halo_bytes, c2w_halo** out) {
  C2W_REQUIRE(out && halo_bytes >= 16 && halo_bytes % 16 == 0, "c2w_halo_create: halo_bytes must be a positive multiple of 16");
  c2w_halo* h = new c2w_halo();
  h->halo_bytes = halo_bytes;
  const size_t total = static_cast<size_t>(h->box_bytes()) + 256;
  if (cudaMalloc(reinterpret_cast<void**>(&h->mailbox), total) != cudaSuccess ||
      cudaMalloc(reinterpret_cast<void**>(&h->done), 256) != cudaSuccess) {
    delete h;
    return fail(C2W_ERR_CUDA, "c2w_halo_create: cudaMalloc of %zu bytes failed", total);
  }
  cudaMemset(h->mailbox, 0, total);
  cudaMemset(h->done, 0, 256);
  h->flags = reinterpret_cast<uint32_t*>(h->mailbox + h->box_bytes());
  *out = h;
  return C2W_OK;
}

void c2w_halo_destroy(c2w_halo* h) {
  if (!h) return;
  for (int s = 0; s < 2; ++s)
    if (h->peer_base[s]) cudaIpcCloseMemHandle(h->peer_base[s]);
  cudaFree(h->mailbox);
  cudaFree(h->done);
  delete h;
}

int c2w_halo_handle(c2w_halo* h, void* out64) {
  C2W_REQUIRE(h && out64, "c2w_halo_handle: bad argument");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "CUDA IPC handles are 64 bytes");
  cudaIpcMemHandle_t hd;
  C2W_CUDA(cudaIpcGetMemHandle(&hd, h->mailbox));
  memcpy(out64, &hd, 64);
  return C2W_OK;
}

int c2w_halo_connect(c2w_halo* h, const void* left_handle64, const void* right_handle64) {
  C2W_REQUIRE(h, "c2w_halo_connect: null handle");
  const void* hs[2] = {left_handle64, right_handle64};
  for (int s = 0; s < 2; ++s) {
    if (!hs[s]) continue;
    cudaIpcMemHandle_t hd;
    memcpy(&hd, hs[s], 64);
    void* base = nullptr;
    cudaError_t e = cudaIpcOpenMemHandle(&base, hd, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess)
      return fail(C2W_ERR_CUDA, "c2w_halo_connect: cudaIpcOpenMemHandle(%s neighbour) -> %s (no peer access between the "
                  "two GPUs?)", s == 0 ? "left" : "right", cudaGetErrorString(e));
    h->peer_base[s] = base;
    h->peer_box[s] = static_cast<uint8_t*>(base);
    h->peer_flags[s] = reinterpret_cast<uint32_t*>(h->peer_box[s] + h->box_bytes());
  }
  return C2W_OK;
}

// Targets of a push FUSED into another kernel (c2w_guided_step with g->halo set): the neighbours' mailbox slots of the
// current parity, their step counters, the CTA counter and the value to publish.  The caller then completes the exchange
// with c2w_halo_pull, which also advances the step.
int c2w_halo_push_targets(c2w_halo* h, void** slot_left, void** slot_right, void** flag_left, void** flag_right, void** done,
                          uint32_t* publish) {
  C2W_REQUIRE(h && slot_left && slot_right && flag_left && flag_right && done && publish, "c2w_halo_push_targets: bad argument");
  const uint32_t par = h->step & 1u;
  const bool has_l = h->peer_box[0] != nullptr, has_r = h->peer_box[1] != nullptr;
  *slot_left = has_l ? h->peer_box[0] + (par * 2 + 1) * h->halo_bytes : nullptr;   // the left rank's "from the right" slot
  *slot_right = has_r ? h->peer_box[1] + (par * 2 + 0) * h->halo_bytes : nullptr;  // the right rank's "from the left" slot
  *flag_left = has_l ? h->peer_flags[0] + 1 : nullptr;
  *flag_right = has_r ? h->peer_flags[1] + 0 : nullptr;
  *done = h->done;
  *publish = h->step + 1;
  return C2W_OK;
}

// Second half of an exchange whose push was fused into the producing kernel: wait for the neighbours' counters, copy
// the mailbox into the halo frames of x_local, advance the step.
int c2w_halo_pull(c2w_halo* h, float* x_local, int64_t n_local_frames, int64_t frame_floats, int32_t k, void* stream) {
  C2W_REQUIRE(h && x_local && k >= 1 && n_local_frames >= 3 * k, "c2w_halo_pull: bad argument");
  const int64_t hb = static_cast<int64_t>(k) * frame_floats * 4;
  C2W_REQUIRE(hb == h->halo_bytes, "c2w_halo_pull: %lld halo bytes, the mailbox was created for %lld", (long long)hb,
              (long long)h->halo_bytes);
  const bool has_l = h->peer_box[0] != nullptr, has_r = h->peer_box[1] != nullptr;
  if (!has_l && !has_r) return C2W_OK;
  const uint32_t par = h->step & 1u;
  auto slot = [&](int side) { return reinterpret_cast<const float4*>(h->mailbox + (par * 2 + side) * hb); };
  halo_pull_kernel<<<64, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      slot(0), has_l ? reinterpret_cast<float4*>(x_local) : nullptr, has_l ? h->flags + 0 : nullptr, slot(1),
      has_r ? reinterpret_cast<float4*>(x_local + (n_local_frames - k) * frame_floats) : nullptr, has_r ? h->flags + 1 : nullptr,
      hb / 16, h->step + 1);
  C2W_CUDA(cudaGetLastError());
  c2w_count_launches(1);
  h->step += 1;
  return C2W_OK;
}

// x_local: [n_local_frames][frame_floats] fp32 with k halo frames on every side that has a neighbour.
int c2w_halo_exchange(c2w_halo* h, float* x_local, int64_t n_local_frames, int64_t frame_floats, int32_t k, void* stream) {
  C2W_REQUIRE(h && x_local && k >= 1 && n_local_frames >= 3 * k, "c2w_halo_exchange: bad argument");
  const int64_t hb = static_cast<int64_t>(k) * frame_floats * 4;
  C2W_REQUIRE(hb == h->halo_bytes, "c2w_halo_exchange: %lld halo bytes, the mailbox was created for %lld", (long long)hb,
              (long long)h->halo_bytes);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const bool has_l = h->peer_box[0] != nullptr, has_r = h->peer_box[1] != nullptr;
  if (!has_l && !has_r) return C2W_OK;
  const uint32_t par = h->step & 1u;
  const long long n16 = hb / 16;
  const int64_t kf = static_cast<int64_t>(k) * frame_floats;
  // mailbox slot [parity][side]: side 0 = data FROM the left neighbour, side 1 = FROM the right neighbour
  auto slot = [&](uint8_t* box, int side) { return reinterpret_cast<float4*>(box + (par * 2 + side) * hb); };
  const float4* src_l = reinterpret_cast<const float4*>(x_local + kf);                            // first k owned frames
  const float4* src_r = reinterpret_cast<const float4*>(x_local + (n_local_frames - 2 * k) * frame_floats);  // last k owned
  // a rank without a left neighbour owns its first frames outright (no halo there): its first owned frame is frame 0
  if (!has_l) src_l = nullptr;
  const int grid = 64;
  halo_push_kernel<<<grid, 256, 0, st>>>(src_l, has_l ? slot(h->peer_box[0], 1) : nullptr, has_l ? h->peer_flags[0] + 1 : nullptr,
                                         src_r, has_r ? slot(h->peer_box[1], 0) : nullptr, has_r ? h->peer_flags[1] + 0 : nullptr,
                                         n16, h->step + 1, h->done);
  C2W_CUDA(cudaGetLastError());
  halo_pull_kernel<<<grid, 256, 0, st>>>(slot(h->mailbox, 0), has_l ? reinterpret_cast<float4*>(x_local) : nullptr,
                                         has_l ? h->flags + 0 : nullptr, slot(h->mailbox, 1),
                                         has_r ? reinterpret_cast<float4*>(x_local + (n_local_frames - k) * frame_floats) : nullptr,
                                         has_r ? h->flags + 1 : nullptr, n16, h->step + 1);
  C2W_CUDA(cudaGetLastError());
  c2w_count_launches(2);
  h->step += 1;
  return C2W_OK;
}

}  // extern "C"
