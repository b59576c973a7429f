// Device code only (no launch syntax): included by exchange.cu for the GPU build and, with SAEB_CPU_EMU defined, by the
// CPU emulation harness under tests/emu (tests/test_kernel_emu.py runs R emulated ranks as processes over shared memory).
#pragma once
#include "common.cuh"

namespace saeb {

constexpr int PUSH_MAX_RANKS = 32;
constexpr int PUSH_THREADS = 256;
constexpr unsigned long long PUSH_TIMEOUT_NS = 20ull * 1000ull * 1000ull * 1000ull;   // 20 s: a peer that never arrives

#if defined(SAEB_CPU_EMU)
// host equivalents: the emulated ranks are processes sharing memory, so C++ atomics give the same release / acquire
inline void st_release_sys_u32(uint32_t* p, uint32_t v) { __atomic_store_n(p, v, __ATOMIC_RELEASE); }
inline uint32_t ld_acquire_sys_u32(const uint32_t* p) { return __atomic_load_n(p, __ATOMIC_ACQUIRE); }
inline unsigned long long global_timer_ns() {
  timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return (unsigned long long)ts.tv_sec * 1000000000ull + (unsigned long long)ts.tv_nsec;
}
inline void multimem_st_u4(uint4* mc_ptr, const uint4& v) { *mc_ptr = v; }   // no multicast mapping on the host
#else
__device__ __forceinline__ void st_release_sys_u32(uint32_t* p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys_u32(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned long long global_timer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)::"memory");
  return t;
}
// one 16-byte store replicated by the switch into every device of the multicast group
__device__ __forceinline__ void multimem_st_u4(uint4* mc_ptr, const uint4& v) {
  asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(mc_ptr), "r"(v.x), "r"(v.y),
               "r"(v.z), "r"(v.w)
               : "memory");
}
#endif

// src [n_vec] uint4 (this rank's slab) -> (peer_bases[p] + slab_off) for every p (or mc_dst once);
// flags of channel c live at (base + flags_off) as uint32 [channels][R], indexed by SOURCE rank.
__global__ void __launch_bounds__(PUSH_THREADS)
push_gather_kernel(const uint4* __restrict__ src, size_t n_vec, void* const* __restrict__ peer_bases, int R, int self,
                   size_t slab_off, uint4* mc_dst, size_t flags_off, int channel, uint32_t seq, int* counter) {
  __shared__ uint8_t* s_base[PUSH_MAX_RANKS];
  __shared__ int s_last;
  if ((int)threadIdx.x < R) s_base[threadIdx.x] = reinterpret_cast<uint8_t*>(peer_bases[threadIdx.x]);
  __syncthreads();
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  if (mc_dst != nullptr) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_vec; i += stride)
      multimem_st_u4(mc_dst + i, src[i]);
  } else {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_vec; i += stride) {
      const uint4 v = src[i];
      // start with the own copy and walk the peers in rank order from there, so that at any moment the R ranks aim
      // at R different destinations
      for (int q = 0; q < R; ++q) {
        int p = self + q;
        if (p >= R) p -= R;
        reinterpret_cast<uint4*>(s_base[p] + slab_off)[i] = v;
      }
    }
  }
  // "last block" pattern: every thread's stores are fenced at system scope before the CTA is counted
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) s_last = (atomicAdd(counter, 1) == (int)gridDim.x - 1) ? 1 : 0;
  __syncthreads();
  if (!s_last) return;
  __threadfence_system();
  if (threadIdx.x == 0) *counter = 0;   // every CTA has been counted: ready for the next launch on this channel
  if ((int)threadIdx.x < R) {
    // publish: "rank `self` has delivered exchange `seq` of this channel" on every peer (and locally)
    uint32_t* remote = reinterpret_cast<uint32_t*>(s_base[threadIdx.x] + flags_off) + (size_t)channel * R + self;
    st_release_sys_u32(remote, seq);
    // wait for peer `threadIdx.x`'s slab in the local buffer
    const uint32_t* mine = reinterpret_cast<const uint32_t*>(s_base[self] + flags_off) + (size_t)channel * R + threadIdx.x;
    const unsigned long long t0 = global_timer_ns();
    while ((int32_t)(ld_acquire_sys_u32(mine) - seq) < 0) {
      if (global_timer_ns() - t0 > PUSH_TIMEOUT_NS) {
        printf("saeb200: push_gather timed out waiting for rank %d (channel %d, seq %u, have %u)\n", (int)threadIdx.x,
               channel, seq, ld_acquire_sys_u32(mine));
        __trap();
      }
    }
  }
}

}  // namespace saeb
