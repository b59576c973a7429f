// Device code only (no launch syntax): included by decode_bwd.cu for the GPU build and, with SAEB_CPU_EMU defined, by the CPU
// emulation harness under tests/emu, which runs these kernels thread by thread on the host (tests/test_kernel_emu.py).
#pragma once
#include "common.cuh"

namespace saeb {

constexpr int DBW_THREADS = 256;

__global__ void __launch_bounds__(DBW_THREADS)
decode_bwd_acts_kernel(const float* __restrict__ g, long long ld_g, const long long* __restrict__ idx, int k,
                       const float* __restrict__ W, long long d, long long N, float* __restrict__ d_vals,
                       int* __restrict__ err_flag) {
  extern __shared__ float gsm[];   // [d4] gradient row (zero padded to a multiple of 4)
  const long long t = blockIdx.x;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
  const int d4 = (int)((d + 3) & ~3ll);
  for (int i = tid; i < d4; i += blockDim.x) gsm[i] = (i < d) ? g[t * ld_g + i] : 0.f;
  __syncthreads();
  const bool vec = (d & 3) == 0 && ((reinterpret_cast<uintptr_t>(W) & 15) == 0);
  for (int j = warp; j < k; j += nwarps) {
    const long long r = idx[t * k + j];
    float acc = 0.f;
    if (r < 0 || r >= N) {
      if (lane == 0 && err_flag) atomicExch(err_flag, 1);   // tl.device_assert(i < N), sae/kernels.py:380
    } else {
      const float* wr = W + r * d;
      float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
      if (vec) {
        const float4* w4 = reinterpret_cast<const float4*>(wr);
        const float4* g4 = reinterpret_cast<const float4*>(gsm);
        const int n4 = (int)(d >> 2);
        for (int c = lane; c < n4; c += 32) {
          const float4 wv = ldg_nc_f4(w4 + c);
          const float4 gv = g4[c];
          a0 = fmaf(wv.x, gv.x, a0);
          a1 = fmaf(wv.y, gv.y, a1);
          a2 = fmaf(wv.z, gv.z, a2);
          a3 = fmaf(wv.w, gv.w, a3);
        }
      } else {
        for (long long i = lane; i < d; i += 32) a0 = fmaf(wr[i], gsm[i], a0);
      }
      acc = (a0 + a1) + (a2 + a3);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) d_vals[t * k + j] = acc;
  }
}

__global__ void __launch_bounds__(DBW_THREADS)
decode_bwd_weight_kernel(const float* __restrict__ g, long long ld_g, const long long* __restrict__ idx,
                         const float* __restrict__ vals, int k, long long d, long long N, float* __restrict__ dW,
                         int* __restrict__ err_flag) {
  extern __shared__ float gsm[];   // [d4]
  const long long t = blockIdx.x;
  const int tid = threadIdx.x;
  const int d4 = (int)((d + 3) & ~3ll);
  for (int i = tid; i < d4; i += blockDim.x) gsm[i] = (i < d) ? g[t * ld_g + i] : 0.f;
  __syncthreads();
  const bool vec = (d & 3) == 0 && ((reinterpret_cast<uintptr_t>(dW) & 15) == 0);
  for (int j = 0; j < k; ++j) {
    const float v = vals[t * k + j];   // block-uniform
    const long long r = idx[t * k + j];
    if (v == 0.f) continue;            // zero values contribute nothing (sae/kernels.py:166)
    if (r < 0 || r >= N) {
      if (tid == 0 && err_flag) atomicExch(err_flag, 1);
      continue;
    }
    float* row = dW + r * d;
    if (vec) {
      const int n4 = (int)(d >> 2);
      for (int c = tid; c < n4; c += blockDim.x) {
        const float4 gv = reinterpret_cast<const float4*>(gsm)[c];
        atomicAdd(reinterpret_cast<float4*>(row) + c, make_float4(v * gv.x, v * gv.y, v * gv.z, v * gv.w));
      }
    } else {
      for (long long i = tid; i < d; i += blockDim.x) atomicAdd(row + i, v * gsm[i]);
    }
  }
}

}  // namespace saeb
