// Peer-memory all-gather of small per-token lists for the feature-sharded scan.
//
// Under feature sharding every token chunk needs two exchanges of [Tc, m] fp32 lists per rank (lower bounds, then exact
// local TopK values).  NCCL's all-gather kernels cannot co-reside with the persistent GEMM CTAs of the next chunk
// (registers), so on the second stream they wait for a GEMM launch boundary and then displace GEMM CTAs.  This kernel
// is small enough to run beside them: every rank PUSHES its slab straight into the gathered buffer of every peer
// (plain 16-byte stores to mapped peer pointers over NVLink, or ONE multimem.st per vector when the buffer has an
// NVSwitch multicast mapping), then the last CTA to finish publishes a per-(channel, source rank) sequence number on
// every peer with a system-scope release store and spins (acquire loads) until all peers' numbers have arrived.  When
// the kernel ends, the local gathered buffer [R][Tc][m] is complete and the next kernel in the stream may read it.
//
// The buffers are symmetric allocations (same offsets on every rank; PyTorch's symmetric-memory rendezvous provides
// the peer / multicast mappings -- plumbing only).  Sequence numbers only grow, so the flags are never reset.
// Re-use of a gathered region is safe without an extra barrier because the two exchanges of a chunk alternate: a rank
// can only push exchange e of chunk c+1 after it has seen every peer's push of the OTHER exchange, which those peers
// issue (stream order) after their last read of region e for chunk c.
#include "common.cuh"

namespace saeb {

constexpr int PUSH_MAX_RANKS = 32;
constexpr int PUSH_THREADS = 256;
constexpr unsigned long long PUSH_TIMEOUT_NS = 20ull * 1000ull * 1000ull * 1000ull;   // 20 s: a peer that never arrives

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

int push_gather_launch(const void* src, size_t bytes, void* const* peer_bases_dev, int R, int self_rank,
                       size_t region_offset, void* multicast_base, size_t flags_offset, int channel, unsigned int seq,
                       int* counter, cudaStream_t stream) {
  SAEB_REQUIRE(R >= 1 && R <= PUSH_MAX_RANKS && self_rank >= 0 && self_rank < R, "push_gather: bad rank %d of %d",
               self_rank, R);
  SAEB_REQUIRE(bytes % 16 == 0 && region_offset % 16 == 0 && (reinterpret_cast<uintptr_t>(src) & 15) == 0,
               "push_gather: slab size, region offset and source must be 16-byte aligned");
  SAEB_REQUIRE(flags_offset % 4 == 0 && channel >= 0, "push_gather: bad flags offset / channel");
  const size_t n_vec = bytes / 16;
  const size_t slab_off = region_offset + (size_t)self_rank * bytes;
  uint4* mc = multicast_base
                  ? reinterpret_cast<uint4*>(reinterpret_cast<uint8_t*>(multicast_base) + slab_off)
                  : nullptr;
  size_t blocks = (n_vec + PUSH_THREADS * 4 - 1) / (PUSH_THREADS * 4);
  if (blocks < 1) blocks = 1;
  if (blocks > 64) blocks = 64;
  push_gather_kernel<<<(unsigned)blocks, PUSH_THREADS, 0, stream>>>(reinterpret_cast<const uint4*>(src), n_vec,
                                                                    peer_bases_dev, R, self_rank, slab_off, mc,
                                                                    flags_offset, channel, seq, counter);
  SAEB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace saeb
