// All-reduce of the gradient arena over NVLink / NVSwitch peer memory: the one exchange step of the
// data-parallel path - DistributedDataParallel's gradient average (nerfstudio/pipelines/base_pipeline.py:305-307) - as ONE
// kernel per rank, fused with the x 1/world average.
//
// Every rank's arena lives in symmetric memory: buffers[p] is rank p's arena mapped into this process (peer pointer),
// signals[p] rank p's signal pad.  Two-shot in place:
//   barrier A   every rank's gradients are complete (this kernel is stream-ordered after the backward on its rank)
//   phase 1     rank r owns slice r: it reads that slice from all W arenas (W - 1 remote reads over NVLink), adds them in
//               rank order, scales, and writes the result into slice r of all W arenas (W - 1 remote writes)
//   barrier B   all slices have landed everywhere
// Per rank (W - 1) / W of the bytes cross NVLink in each direction - half of what a ring all-reduce moves - and there are
// two synchronisations instead of 2 (W - 1) ring steps.  Barriers are epoch-stamped flags in the signal pads (release
// store / acquire load at system scope), so nothing is reset between calls and the kernel can be replayed from a CUDA graph.
#include "common.cuh"

namespace nrb {

constexpr int kPeerMax = 8;  // ranks of one NVSwitch domain (one box)

struct PeerArgs {
  float* buffers[kPeerMax];
  uint32_t* signals[kPeerMax];  // flag area of each rank: [slot][2 barriers][kPeerMax] + epoch / counter words
  float* multicast;             // the arenas as ONE multicast object (NVLS), or null
  int rank, world;
};

__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ float4 ld_peer(const float4* p) {  // never from a stale L1 line: system-scope relaxed load
  float4 v;
  asm volatile("ld.relaxed.sys.global.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
  return v;
}

// In-switch reduction (NVLS): one load returns the sum of the W replicas of a multicast address, one store writes all W
// replicas.  Per rank S / W bytes of reduced data come in and S / W go out (the switch fans them out), instead of
// (W - 1) / W x S in each direction for the loads and again for the stores of the unicast version.
__device__ __forceinline__ float4 multimem_ld_reduce(const float4* p) {
  float4 v;
  asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0, %1, %2, %3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void multimem_st(float4* p, float4 v) {
  asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

// flag words of one slot inside a pad (uint32 indices): A flags [0, 16), B flags [16, 32), epoch 32, CTA counter 33
constexpr int kSlotWords = 64;

__global__ void __launch_bounds__(512) peer_all_reduce_kernel(const __grid_constant__ PeerArgs a, int slot, int64_t offset,
                                                              int64_t n, float scale) {
  __shared__ uint32_t s_epoch;
  __shared__ bool s_last;
  uint32_t* mine = a.signals[a.rank] + slot * kSlotWords;
  if (threadIdx.x == 0) s_epoch = ld_acquire_sys(mine + 32) + 1u;  // (the last CTA of the previous call advanced it)
  __syncthreads();
  const uint32_t epoch = s_epoch;
  // ---- barrier A: tell every peer that this rank's arena is complete, wait until all of them said so
  if (blockIdx.x == 0 && threadIdx.x < a.world) st_release_sys(a.signals[threadIdx.x] + slot * kSlotWords + a.rank, epoch);
  if (threadIdx.x < a.world) {
    while (static_cast<int32_t>(ld_acquire_sys(mine + threadIdx.x) - epoch) < 0) __nanosleep(40);
  }
  __syncthreads();
  // ---- phase 1: reduce my slice from every arena, scatter the result to every arena
  const int64_t n4 = n >> 2;                               // (n and offset are multiples of 4 floats)
  const int64_t per = (n4 + a.world - 1) / a.world;
  const int64_t lo = min(per * a.rank, n4), hi = min(lo + per, n4);
  const int64_t base4 = offset >> 2;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  if (a.multicast != nullptr) {
    float4* mc = reinterpret_cast<float4*>(a.multicast) + base4;
    constexpr int kU = 4;  // independent elements in flight per thread
    for (int64_t i0 = lo + static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i0 < hi; i0 += kU * stride) {
      float4 v[kU];
#pragma unroll
      for (int u = 0; u < kU; ++u)
        if (i0 + u * stride < hi) v[u] = multimem_ld_reduce(mc + i0 + u * stride);
#pragma unroll
      for (int u = 0; u < kU; ++u)
        if (i0 + u * stride < hi) {
          v[u].x *= scale, v[u].y *= scale, v[u].z *= scale, v[u].w *= scale;
          multimem_st(mc + i0 + u * stride, v[u]);
        }
    }
  } else {
    for (int64_t i = lo + static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < hi; i += stride) {
      // all W loads of an element are issued before the first is used (a remote load takes ~1.8 K cycles)
      float4 v[kPeerMax];
#pragma unroll
      for (int p = 0; p < kPeerMax; ++p)
        if (p < a.world) v[p] = ld_peer(reinterpret_cast<const float4*>(a.buffers[p]) + base4 + i);
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int p = 0; p < kPeerMax; ++p)                   // rank order: every rank would compute the same bits
        if (p < a.world) acc.x += v[p].x, acc.y += v[p].y, acc.z += v[p].z, acc.w += v[p].w;
      acc.x *= scale, acc.y *= scale, acc.z *= scale, acc.w *= scale;
#pragma unroll
      for (int p = 0; p < kPeerMax; ++p)
        if (p < a.world) reinterpret_cast<float4*>(a.buffers[p])[base4 + i] = acc;
    }
  }
  // ---- barrier B: the last CTA of this rank to finish announces it and waits for the other ranks
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) s_last = atomicAdd(mine + 33, 1u) == gridDim.x - 1;
  __syncthreads();
  if (!s_last) return;
  if (threadIdx.x < a.world) st_release_sys(a.signals[threadIdx.x] + slot * kSlotWords + 16 + a.rank, epoch);
  if (threadIdx.x < a.world) {
    while (static_cast<int32_t>(ld_acquire_sys(mine + 16 + threadIdx.x) - epoch) < 0) __nanosleep(40);
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    mine[33] = 0u;
    st_release_sys(mine + 32, epoch);  // next call's epoch
  }
}

}  // namespace nrb

using namespace nrb;

extern "C" int nrb_peer_all_reduce(const uint64_t* buffer_ptrs, const uint64_t* signal_flag_ptrs, uint64_t multicast_ptr,
                                   int32_t rank, int32_t world, int32_t slot, int64_t offset, int64_t n, float scale, int32_t max_ctas,
                                   nrb_stream_t stream) {
  NRB_REQUIRE(buffer_ptrs && signal_flag_ptrs, NRB_ERR_BAD_ARG, "nrb_peer_all_reduce: null pointer table");
  NRB_REQUIRE(world >= 1 && world <= kPeerMax && rank >= 0 && rank < world, NRB_ERR_BAD_ARG, "nrb_peer_all_reduce: bad rank / world");
  NRB_REQUIRE(slot >= 0 && slot < 4, NRB_ERR_BAD_ARG, "nrb_peer_all_reduce: slot must be 0..3");
  NRB_REQUIRE(offset >= 0 && n >= 0 && (offset & 3) == 0 && (n & 3) == 0, NRB_ERR_ALIGNMENT,
              "nrb_peer_all_reduce: offset and n must be multiples of 4 floats");
  if (n == 0) return NRB_OK;
  PeerArgs a{};
  for (int p = 0; p < world; ++p) {
    NRB_REQUIRE(buffer_ptrs[p] != 0 && signal_flag_ptrs[p] != 0 && (buffer_ptrs[p] & 15) == 0, NRB_ERR_BAD_ARG,
                "nrb_peer_all_reduce: peer %d pointer null or unaligned", p);
    a.buffers[p] = reinterpret_cast<float*>(buffer_ptrs[p]);
    a.signals[p] = reinterpret_cast<uint32_t*>(signal_flag_ptrs[p]);
  }
  a.multicast = reinterpret_cast<float*>(multicast_ptr);
  a.rank = rank, a.world = world;
  const int64_t per4 = ((n >> 2) + world - 1) / world;
  int ctas = max_ctas > 0 ? max_ctas : 2 * sm_count();
  ctas = static_cast<int>(std::max<int64_t>(1, std::min<int64_t>(ctas, (per4 + 511) / 512)));
  peer_all_reduce_kernel<<<ctas, 512, 0, static_cast<cudaStream_t>(stream)>>>(a, slot, offset, n, scale);
  return finish_launch("nrb_peer_all_reduce");
}
