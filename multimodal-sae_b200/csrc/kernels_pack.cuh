// Device code only (no launch syntax): included by pack.cu for the GPU build and, with SAEB_CPU_EMU defined, by the CPU
// emulation harness under tests/emu, which runs these kernels thread by thread on the host (tests/test_kernel_emu.py).
#pragma once
#include "common.cuh"

namespace saeb {

// mode 3 statistics, one warp per feature row: folded bias, ||w_j|| (rounded up), max |W| and max ||w_j|| (bit patterns)
__global__ void w_stats_kernel(const float* __restrict__ W, const float* __restrict__ b_enc,
                               const float* __restrict__ b_dec, long long N, long long d, float* __restrict__ bias,
                               float* __restrict__ wnorm, unsigned int* __restrict__ absmax_bits,
                               unsigned int* __restrict__ wnorm_max_bits) {
  const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= N) return;
  double dot = 0.0, sq = 0.0;
  float amax = 0.f;
  for (long long i = lane; i < d; i += 32) {
    const float w = W[row * d + i];
    dot += (double)w * (double)b_dec[i];
    sq += (double)w * (double)w;
    amax = fmaxf(amax, fabsf(w));
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    dot += __shfl_xor_sync(0xffffffffu, dot, o);
    sq += __shfl_xor_sync(0xffffffffu, sq, o);
    amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, o));
  }
  if (lane == 0) {
    bias[row] = (float)((double)b_enc[row] - dot);
    const float nrm = (float)sqrt(sq) * (1.0f + 1e-6f);   // rounded up: it is used in an upper bound
    wnorm[row] = nrm;
    atomicMax(absmax_bits, __float_as_uint(amax));        // non-negative floats order like unsigned ints
    atomicMax(wnorm_max_bits, __float_as_uint(nrm));
  }
}

// one warp per feature row: scaled fp16 plane + the exact norm of its rounding error
__global__ void pack_w_f16_kernel(const float* __restrict__ W, long long N, long long d, long long d_pad,
                                  const unsigned int* __restrict__ absmax_bits, __half* __restrict__ out,
                                  float* __restrict__ dnorm, float* __restrict__ trailer) {
  const float amax = __uint_as_float(*absmax_bits);
  int e = 0;
  if (amax > 0.f) frexpf(amax, &e);            // amax = m * 2^e, m in [0.5, 1)
  const float scale = ldexpf(1.0f, 14 - e);    // largest |W| lands in [2^13, 2^14): far from fp16 overflow/underflow
  const float unscale = ldexpf(1.0f, e - 14);
  if (blockIdx.x == 0 && threadIdx.x == 0) trailer[0] = unscale;
  const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= N) return;
  double sq = 0.0;
  for (long long c = lane; c < d_pad; c += 32) {
    const float ws = (c < d) ? W[row * d + c] * scale : 0.f;   // exact: power-of-two scale
    const __half h = __float2half_rn(ws);
    out[row * d_pad + c] = h;
    const double dw = (double)ws - (double)__half2float(h);
    sq += dw * dw;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
  if (lane == 0) {
    const float dn = (float)(sqrt(sq) * (double)unscale) * (1.0f + 1e-6f);   // rounded up: used in an upper bound
    dnorm[row] = dn;
    atomicMax(reinterpret_cast<unsigned int*>(trailer + 3), __float_as_uint(dn));
  }
}

// mode 4: residual plane lo = fp16((W * scale - fp16(W * scale)) * 2^11), scale = 1 / trailer[0] (see pack.cu)
__global__ void pack_w_lo_kernel(const float* __restrict__ W, long long N, long long d, long long d_pad,
                                 const float* __restrict__ trailer, __half* __restrict__ lo) {
  const float scale = 1.0f / trailer[0];   // exact: trailer[0] is a power of two
  const long long total = N * d_pad;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    const long long r = i / d_pad, c = i - r * d_pad;
    const float ws = (c < d) ? W[r * d + c] * scale : 0.f;
    const float hi = __half2float(__float2half_rn(ws));
    lo[i] = __float2half_rn((ws - hi) * 2048.0f);
  }
}

// activations -> power-of-two row-scaled fp16 plane + row scale + ||x|| + ||x - fp16(x)|| (see pack.cu)
template <typename Tin>
__global__ void prep_x_f16_kernel(const Tin* __restrict__ x, long long T, long long d, long long ld_x, long long d_pad,
                                  __half* __restrict__ out, float* __restrict__ row_scale, float* __restrict__ xnorm,
                                  float* __restrict__ xdnorm) {
  const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= T) return;
  const Tin* xr = x + row * ld_x;
  float amax = 0.f, sq = 0.f;
  for (long long i = lane; i < d; i += 32) {
    const float v = (float)xr[i];
    amax = fmaxf(amax, fabsf(v));
    sq = fmaf(v, v, sq);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, o));
    sq += __shfl_xor_sync(0xffffffffu, sq, o);
  }
  int e = 0;
  if (amax > 0.f && amax < 3.0e38f) frexpf(amax, &e);
  const float scale = ldexpf(1.0f, 14 - e);
  if (lane == 0) {
    row_scale[row] = ldexpf(1.0f, e - 14);
    xnorm[row] = sqrtf(sq) * (1.0f + 1e-5f);
  }
  __half* o = out + row * d_pad;
  float dsq = 0.f;   // squared norm of the rounding error (0 for bf16 / fp16 inputs away from underflow)
  for (long long i = lane; i < d_pad; i += 32) {
    const float xs = (i < d) ? (float)xr[i] * scale : 0.f;
    const __half h = __float2half_rn(xs);
    o[i] = h;
    const float dx = xs - __half2float(h);
    dsq = fmaf(dx, dx, dsq);
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) dsq += __shfl_xor_sync(0xffffffffu, dsq, off);
  if (lane == 0) xdnorm[row] = sqrtf(dsq) * ldexpf(1.0f, e - 14) * (1.0f + 1e-5f);
}

}  // namespace saeb
