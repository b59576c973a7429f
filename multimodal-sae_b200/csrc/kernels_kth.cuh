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
  // Member values of the sharded refinement (value_mode 2) are 3e38 ("certainly in the TopK") for all but a handful of
  // entries: with c of them, the kth largest is 3e38 if c >= kth, else the (kth - c)-th largest of the rest -- a few
  // max-extractions instead of the 31-step search.  Generic inputs (no such values) fall through to the search.
  const uint32_t SURE = __float_as_uint(3.0e38f);
  int c_sure = 0, n_pos = 0;
#pragma unroll
  for (int s = 0; s < VPL; ++s) {
    c_sure += (key[s] == SURE) ? 1 : 0;
    n_pos += (key[s] != 0u) ? 1 : 0;
  }
  const int both = __reduce_add_sync(0xffffffffu, c_sure | (n_pos << 16));
  c_sure = both & 0xffff;
  n_pos = both >> 16;
  uint32_t prefix = 0;
  if (c_sure >= kth) {
    prefix = SURE;
  } else if (n_pos < kth) {
    prefix = 0;
  } else if (c_sure > 0 && kth - c_sure <= 8) {
    const int r = kth - c_sure;
    for (int it = 0; it < r; ++it) {
      uint32_t mx = 0;
#pragma unroll
      for (int s = 0; s < VPL; ++s) mx = (key[s] != SURE && key[s] > mx) ? key[s] : mx;
      const uint32_t wm = __reduce_max_sync(0xffffffffu, mx);
      prefix = wm;
      // drop ONE instance of the maximum (equal values count separately)
      const uint32_t holders = __ballot_sync(0xffffffffu, mx == wm);
      if (lane == __ffs((int)holders) - 1) {
        bool removed = false;
#pragma unroll
        for (int s = 0; s < VPL; ++s) {
          if (!removed && key[s] == wm) {
            key[s] = 0u;
            removed = true;
          }
        }
      }
    }
  } else {
    for (int bit = 30; bit >= 0; --bit) {
      const uint32_t trial = prefix | (1u << bit);
      int c = 0;
#pragma unroll
      for (int s = 0; s < VPL; ++s) c += (key[s] >= trial) ? 1 : 0;
      c = __reduce_add_sync(0xffffffffu, c);
      if (c >= kth) prefix = trial;
    }
  }
  if (lane == 0) tok_thr[t] = __uint_as_float(prefix);
}

// Exchange 1 of the feature-sharded scan in one pass: g [R][T][2*m1] holds, per shard and token, the m1 largest lower
// bounds followed by the m1 largest upper bounds (both descending).  One warp per token:
//   ext_L[t] = k-th largest of the R*m1 lower bounds                      (0 if there are fewer than k)
//   ext_U[t] = max((k+1)-th largest of the R*m1 upper bounds (0 if fewer), max over shards of the SMALLEST upper
//              bound the shard sent)
// -- what a shard did not send is no larger than the last column it did send, so ext_U bounds the token's (k+1)-th
// largest upper bound over ALL latents from above.  Replaces two strided copies, two kth launches, an amax and a
// maximum of saeb200.dist; needs no shared memory, so it runs beside a resident GEMM CTA.
template <int VPL>
__global__ void __launch_bounds__(128)
gathered_bounds_kernel(const float* __restrict__ g, int R, long long T, int m1, int k, float* __restrict__ ext_L,
                       float* __restrict__ ext_U) {
  const long long t = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (t >= T) return;
  const int M = R * m1;
  uint32_t keyL[VPL], keyU[VPL];
  uint32_t tail = 0;
#pragma unroll
  for (int s = 0; s < VPL; ++s) {
    const int i = s * 32 + lane;
    uint32_t bl = 0, bu = 0;
    if (i < M) {
      const int r = i / m1, j = i - r * m1;
      const float* row = g + ((long long)r * T + t) * (2 * m1);
      const float vl = __ldg(row + j), vu = __ldg(row + m1 + j);
      if (vl > 0.f) bl = __float_as_uint(vl);
      if (vu > 0.f) bu = __float_as_uint(vu);
    }
    keyL[s] = bl;
    keyU[s] = bu;
  }
  // per shard the SMALLEST upper bound it sent (0 if its list is not full: then nothing was cut off); lists need not
  // be sorted
  for (int r = 0; r < R; ++r) {
    const float* row = g + ((long long)r * T + t) * (2 * m1) + m1;
    uint32_t mn = 0xffffffffu;
    for (int j = lane; j < m1; j += 32) {
      const float vu = __ldg(row + j);
      const uint32_t b = vu > 0.f ? __float_as_uint(vu) : 0u;
      mn = b < mn ? b : mn;
    }
    mn = __reduce_min_sync(0xffffffffu, mn);
    tail = mn > tail ? mn : tail;
  }
  uint32_t pl = 0, pu = 0;
  for (int bit = 30; bit >= 0; --bit) {
    const uint32_t tl = pl | (1u << bit), tu = pu | (1u << bit);
    int cl = 0, cu = 0;
#pragma unroll
    for (int s = 0; s < VPL; ++s) {
      cl += (keyL[s] >= tl) ? 1 : 0;
      cu += (keyU[s] >= tu) ? 1 : 0;
    }
    // one reduction for both counters (each is at most 32 * VPL <= 2048)
    const int both = __reduce_add_sync(0xffffffffu, cl | (cu << 16));
    if ((both & 0xffff) >= k) pl = tl;
    if ((both >> 16) >= k + 1) pu = tu;
  }
  if (k > M) pl = 0;
  if (k + 1 > M) pu = 0;
  if (lane == 0) {
    ext_L[t] = __uint_as_float(pl);
    ext_U[t] = __uint_as_float(pu > tail ? pu : tail);
  }
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
