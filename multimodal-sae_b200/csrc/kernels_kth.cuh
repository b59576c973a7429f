// Device code only (no launch syntax): included by coo_scan.cu for the GPU build and, with SAEB_CPU_EMU defined, by the CPU
// emulation harness under tests/emu, which runs these kernels thread by thread on the host (tests/test_kernel_emu.py).
#pragma once
#include "common.cuh"

namespace saeb {

template <int VPL>
__global__ void __launch_bounds__(256)
kth_gathered_reg_kernel(const float* __restrict__ g, int R, long long T, int m, int kth, float* __restrict__ tok_thr) {
  const long long t = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (t >= T) return;   // t is per warp: the whole warp leaves together
  const int M = R * m;
  uint32_t key[VPL];
#pragma unroll
  for (int s = 0; s < VPL; ++s) {
    const int i = s * 32 + lane;
    uint32_t b = 0;
    if (i < M) {
      const int r = i / m, j = i - r * m;
      const float v = __ldg(g + ((long long)r * T + t) * m + j);
      if (v > 0.f) b = __float_as_uint(v);   // positive floats order like their bit patterns; <= 0 and NaN count as 0
    }
    key[s] = b;
  }
  uint32_t prefix = 0;
  for (int bit = 30; bit >= 0; --bit) {
    const uint32_t trial = prefix | (1u << bit);
    int c = 0;
#pragma unroll
    for (int s = 0; s < VPL; ++s) c += (key[s] >= trial) ? 1 : 0;
    c = __reduce_add_sync(0xffffffffu, c);
    if (c >= kth) prefix = trial;
  }
  if (lane == 0) tok_thr[t] = __uint_as_float(prefix);
}

// any R*m: the values stay in memory (L1/L2) and are re-read in every search step
__global__ void kth_gathered_mem_kernel(const float* __restrict__ g, int R, long long T, int m, int kth,
                                        float* __restrict__ tok_thr) {
  const long long t = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (t >= T) return;
  const int M = R * m;
  uint32_t prefix = 0;
  for (int bit = 30; bit >= 0; --bit) {
    const uint32_t trial = prefix | (1u << bit);
    int c = 0;
    for (int i = lane; i < M; i += 32) {
      const int r = i / m, j = i - r * m;
      const float v = g[((long long)r * T + t) * m + j];
      c += (v > 0.f && __float_as_uint(v) >= trial) ? 1 : 0;
    }
    c = __reduce_add_sync(0xffffffffu, c);
    if (c >= kth) prefix = trial;
  }
  if (lane == 0) tok_thr[t] = __uint_as_float(prefix);
}

}  // namespace saeb
