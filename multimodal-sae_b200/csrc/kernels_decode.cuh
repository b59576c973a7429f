// Device code only (no launch syntax): included by decode.cu for the GPU build and, with SAEB_CPU_EMU defined, by the CPU
// emulation harness under tests/emu, which runs these kernels thread by thread on the host (tests/test_kernel_emu.py).
#pragma once
#include "common.cuh"

namespace saeb {

constexpr int DEC_THREADS = 256;
constexpr int DEC_KMAX = 1024;
constexpr int DEC_UNROLL = 8;

template <typename WT>
__device__ __forceinline__ float4 load_w4(const WT* row, int col4);
template <>
__device__ __forceinline__ float4 load_w4<float>(const float* row, int col4) {
  return ldg_nc_f4(reinterpret_cast<const float4*>(row) + col4);
}
template <>
__device__ __forceinline__ float4 load_w4<__nv_bfloat16>(const __nv_bfloat16* row, int col4) {
  uint2 u;
  const uint2* p = reinterpret_cast<const uint2*>(row) + col4;
#if defined(SAEB_CPU_EMU)
  u = *p;
#else
  asm volatile("ld.global.nc.L1::no_allocate.v2.u32 {%0, %1}, [%2];" : "=r"(u.x), "=r"(u.y) : "l"(p));
#endif
  float4 f;
  f.x = __uint_as_float(u.x << 16);
  f.y = __uint_as_float(u.x & 0xffff0000u);
  f.z = __uint_as_float(u.y << 16);
  f.w = __uint_as_float(u.y & 0xffff0000u);
  return f;
}

template <>
__device__ __forceinline__ float4 load_w4<__half>(const __half* row, int col4) {
  uint2 u;
  const uint2* p = reinterpret_cast<const uint2*>(row) + col4;
#if defined(SAEB_CPU_EMU)
  u = *p;
#else
  asm volatile("ld.global.nc.L1::no_allocate.v2.u32 {%0, %1}, [%2];" : "=r"(u.x), "=r"(u.y) : "l"(p));
#endif
  const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&u.x));
  const float2 b = __half22float2(*reinterpret_cast<const __half2*>(&u.y));
  return make_float4(a.x, a.y, b.x, b.y);
}

template <typename XT>
__device__ __forceinline__ float4 load_x4(const XT* row, int col4);
template <>
__device__ __forceinline__ float4 load_x4<float>(const float* row, int col4) {
  return reinterpret_cast<const float4*>(row)[col4];
}
template <>
__device__ __forceinline__ float4 load_x4<__nv_bfloat16>(const __nv_bfloat16* row, int col4) {
  const uint2 u = reinterpret_cast<const uint2*>(row)[col4];
  return make_float4(__uint_as_float(u.x << 16), __uint_as_float(u.x & 0xffff0000u), __uint_as_float(u.y << 16),
                     __uint_as_float(u.y & 0xffff0000u));
}
template <>
__device__ __forceinline__ float4 load_x4<__half>(const __half* row, int col4) {
  const uint2 u = reinterpret_cast<const uint2*>(row)[col4];
  const __half2 a = *reinterpret_cast<const __half2*>(&u.x), b = *reinterpret_cast<const __half2*>(&u.y);
  const float2 fa = __half22float2(a), fb = __half22float2(b);
  return make_float4(fa.x, fa.y, fb.x, fb.y);
}

template <typename OT>
__device__ __forceinline__ void store_o4(OT* row, int col4, float4 v);
template <>
__device__ __forceinline__ void store_o4<float>(float* row, int col4, float4 v) {
  reinterpret_cast<float4*>(row)[col4] = v;
}
template <>
__device__ __forceinline__ void store_o4<__half>(__half* row, int col4, float4 v) {
  __half2 a = __floats2half2_rn(v.x, v.y), b = __floats2half2_rn(v.z, v.w);
  uint2 u;
  u.x = *reinterpret_cast<uint32_t*>(&a);
  u.y = *reinterpret_cast<uint32_t*>(&b);
  reinterpret_cast<uint2*>(row)[col4] = u;
}
template <>
__device__ __forceinline__ void store_o4<__nv_bfloat16>(__nv_bfloat16* row, int col4, float4 v) {
  __nv_bfloat162 a = __floats2bfloat162_rn(v.x, v.y), b = __floats2bfloat162_rn(v.z, v.w);
  uint2 u;
  u.x = *reinterpret_cast<uint32_t*>(&a);
  u.y = *reinterpret_cast<uint32_t*>(&b);
  reinterpret_cast<uint2*>(row)[col4] = u;
}

template <typename WT, typename OT, typename XT>
__global__ void __launch_bounds__(DEC_THREADS)
decode_kernel(const long long* __restrict__ idx, const float* __restrict__ vals, int k, const WT* __restrict__ W,
              long long d, long long N, const float* __restrict__ b_dec, OT* __restrict__ out, long long ld_out,
              const XT* __restrict__ x, long long ld_x, double* __restrict__ sq_err, int* __restrict__ err_flag,
              long long T) {
  __shared__ long long s_row[DEC_KMAX];
  __shared__ float s_val[DEC_KMAX];
  __shared__ int s_n;
  __shared__ float s_red[DEC_THREADS / 32];
  const int tid = threadIdx.x;
  float local_sq = 0.f;
  // one token per CTA when gridDim.x >= T; a smaller (persistent) grid walks the tokens with stride gridDim.x
  for (long long t = blockIdx.x; t < T; t += gridDim.x) {
  if (tid == 0) s_n = 0;
  __syncthreads();
  // compact away zero activations (sae/kernels.py:277) keeping j order; validate indices (kernels.py:276)
  if (tid < 32) {
    int n = 0;
    for (int j0 = 0; j0 < k; j0 += 32) {
      const int j = j0 + tid;
      float v = 0.f;
      long long r = 0;
      if (j < k) {
        v = vals[t * k + j];
        r = idx[t * k + j];
        if (r < 0 || r >= N) {
          if (err_flag) atomicExch(err_flag, 1);
          v = 0.f;
        }
      }
      const bool keep = v != 0.f;
      const uint32_t m = __ballot_sync(0xffffffffu, keep);
      if (keep) {
        const int p = n + __popc(m & ((1u << tid) - 1u));
        s_row[p] = r * d;
        s_val[p] = v;
      }
      n += __popc(m);
    }
    if (tid == 0) s_n = n;
  }
  __syncthreads();
  const int n = s_n;
  const int ncol4 = (int)(d >> 2);
  for (int c4 = tid; c4 < ncol4; c4 += DEC_THREADS) {
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    int j = 0;
    for (; j + DEC_UNROLL <= n; j += DEC_UNROLL) {
      float4 w[DEC_UNROLL];
#pragma unroll
      for (int u = 0; u < DEC_UNROLL; ++u) w[u] = load_w4<WT>(W + s_row[j + u], c4);
#pragma unroll
      for (int u = 0; u < DEC_UNROLL; ++u) {
        const float v = s_val[j + u];
        acc.x = fmaf(v, w[u].x, acc.x);
        acc.y = fmaf(v, w[u].y, acc.y);
        acc.z = fmaf(v, w[u].z, acc.z);
        acc.w = fmaf(v, w[u].w, acc.w);
      }
    }
    for (; j < n; ++j) {
      const float4 w = load_w4<WT>(W + s_row[j], c4);
      const float v = s_val[j];
      acc.x = fmaf(v, w.x, acc.x);
      acc.y = fmaf(v, w.y, acc.y);
      acc.z = fmaf(v, w.z, acc.z);
      acc.w = fmaf(v, w.w, acc.w);
    }
    if (b_dec != nullptr) {
      const float4 b = reinterpret_cast<const float4*>(b_dec)[c4];
      acc.x += b.x; acc.y += b.y; acc.z += b.z; acc.w += b.w;
    }
    store_o4<OT>(out + t * ld_out, c4, acc);
    if (sq_err != nullptr) {
      const float4 xv = load_x4<XT>(x + t * ld_x, c4);
      const float e0 = acc.x - xv.x, e1 = acc.y - xv.y, e2 = acc.z - xv.z, e3 = acc.w - xv.w;
      local_sq += e0 * e0 + e1 * e1 + e2 * e2 + e3 * e3;
    }
  }
  if (sq_err != nullptr) {
    // per token: warp sums -> one double add (same rounding whether the grid is one CTA per token or persistent)
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) local_sq += __shfl_xor_sync(0xffffffffu, local_sq, o);
    if ((tid & 31) == 0) s_red[tid >> 5] = local_sq;
    __syncthreads();
    if (tid == 0) {
      double s = 0.0;
      for (int w = 0; w < DEC_THREADS / 32; ++w) s += (double)s_red[w];
      atomicAdd(sq_err, s);
    }
    local_sq = 0.f;
  }
  __syncthreads();   // s_row / s_val / s_red are reused by the next token
  }
}

// generic fallback for shapes the vector kernel cannot take (d, ld_out or ld_x not a multiple of 4): one column per
// thread, scalar loads.  Same arithmetic (fp32 fma in j order, zeros skipped).
template <typename WT, typename OT, typename XT>
__global__ void __launch_bounds__(DEC_THREADS)
decode_scalar_kernel(const long long* __restrict__ idx, const float* __restrict__ vals, int k,
                     const WT* __restrict__ W, long long d, long long N, const float* __restrict__ b_dec,
                     OT* __restrict__ out, long long ld_out, const XT* __restrict__ x, long long ld_x,
                     double* __restrict__ sq_err, int* __restrict__ err_flag, long long T) {
  __shared__ float s_red[DEC_THREADS / 32];
  const int tid = threadIdx.x;
  for (long long t = blockIdx.x; t < T; t += gridDim.x) {
  float local_sq = 0.f;
  for (long long c = tid; c < d; c += DEC_THREADS) {
    float acc = 0.f;
    for (int j = 0; j < k; ++j) {
      const float v = vals[t * k + j];
      const long long r = idx[t * k + j];
      if (r < 0 || r >= N) {
        if (err_flag) atomicExch(err_flag, 1);
        continue;
      }
      if (v != 0.f) acc = fmaf(v, (float)W[r * d + c], acc);
    }
    if (b_dec != nullptr) acc += b_dec[c];
    out[t * ld_out + c] = (OT)acc;
    if (sq_err != nullptr) {
      const float e = acc - (float)x[t * ld_x + c];
      local_sq += e * e;
    }
  }
  if (sq_err != nullptr) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) local_sq += __shfl_xor_sync(0xffffffffu, local_sq, o);
    if ((tid & 31) == 0) s_red[tid >> 5] = local_sq;
    __syncthreads();
    if (tid == 0) {
      double s2 = 0.0;
      for (int w = 0; w < DEC_THREADS / 32; ++w) s2 += (double)s_red[w];
      atomicAdd(sq_err, s2);
    }
    __syncthreads();
  }
  }
}

// column statistics for the FVU denominator  sum((x - mean_0(x))^2)  (sae/sae.py:204), fp64 accumulation
template <typename XT>
__global__ void colstats_kernel(const XT* __restrict__ x, long long T, long long d, long long ld_x, int rows_per_block,
                                double* __restrict__ colsum, double* __restrict__ colsq) {
  const long long r0 = (long long)blockIdx.x * rows_per_block;
  const long long r1 = (r0 + rows_per_block < T) ? r0 + rows_per_block : T;
  for (long long c = threadIdx.x; c < d; c += blockDim.x) {
    double s = 0.0, q = 0.0;
    for (long long r = r0; r < r1; ++r) {
      const double v = (double)(float)x[r * ld_x + c];
      s += v;
      q += v * v;
    }
    atomicAdd(colsum + c, s);
    atomicAdd(colsq + c, q);
  }
}
__global__ void totvar_kernel(const double* __restrict__ colsum, const double* __restrict__ colsq, long long T,
                              long long d, double* __restrict__ out) {
  __shared__ double red[256];
  double acc = 0.0;
  for (long long c = threadIdx.x; c < d; c += blockDim.x) acc += colsq[c] - colsum[c] * colsum[c] / (double)T;
  red[threadIdx.x] = acc;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if ((int)threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) *out = red[0];
}

}  // namespace saeb
