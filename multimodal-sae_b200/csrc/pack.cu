// One-time weight repack and activation plane split.
//
// reference parameters (sae/sae.py:59-66): encoder.weight [N,d] fp32, encoder.bias [N], W_dec [N,d], b_dec [d].
//   W planes : W = bf16(W) + bf16(W - bf16(W)) + O(2^-17 |W|)          -> [BP][N][d] bf16
//   bias     : W (x - b_dec) + b_enc = W x + (b_enc - W b_dec)         -> [N] fp32 (dot product in fp64)
// fp16 / fp32 activations are split into two bf16 planes the same way (fp16 splits exactly: 11 = 8 + 3 bits);
// bf16 activations (the cache path, launch/cache/cache_image.py:36-39) are consumed in place.
#include "kernels_pack.cuh"

namespace saeb {

__global__ void pack_w_kernel(const float* __restrict__ W, long long N, long long d, long long d_pad, int planes,
                              __nv_bfloat16* __restrict__ out) {
  const long long total = N * d_pad;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    const long long r = i / d_pad, c = i - r * d_pad;
    const float f = (c < d) ? W[r * d + c] : 0.f;
    const __nv_bfloat16 hi = __float2bfloat16_rn(f);
    out[i] = hi;
    if (planes > 1) out[total + i] = __float2bfloat16_rn(f - __bfloat162float(hi));
  }
}

// one warp per feature row: bias_f[n] = b_enc[n] - sum_i W[n,i] * b_dec[i]  (fp64 accumulate)
__global__ void fold_bias_kernel(const float* __restrict__ W, const float* __restrict__ b_enc,
                                 const float* __restrict__ b_dec, long long N, long long d,
                                 float* __restrict__ out) {
  const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= N) return;
  double acc = 0.0;
  for (long long i = lane; i < d; i += 32) acc += (double)W[row * d + i] * (double)b_dec[i];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if (lane == 0) out[row] = (float)((double)b_enc[row] - acc);
}

// activations -> `planes` bf16 planes [planes][T][d_pad] (zero padded columns); planes == 1 is a padded copy
template <typename Tin>
__global__ void split_x_kernel(const Tin* __restrict__ x, long long T, long long d, long long ld_x, long long d_pad,
                               int planes, __nv_bfloat16* __restrict__ out) {
  const long long total = T * d_pad;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    const long long r = i / d_pad, c = i - r * d_pad;
    const float f = (c < d) ? (float)x[r * ld_x + c] : 0.f;
    const __nv_bfloat16 hi = __float2bfloat16_rn(f);
    out[i] = hi;
    if (planes > 1) out[total + i] = __float2bfloat16_rn(f - __bfloat162float(hi));
  }
}

int pack_weights_launch(const float* W_enc, const float* b_enc, const float* b_dec, long long N, long long d,
                        long long d_pad, int planes, void* w_planes, float* bias, cudaStream_t stream) {
  SAEB_REQUIRE(N > 0 && d > 0, "pack: need N>0 and d>0 (got N=%lld d=%lld)", N, d);
  SAEB_REQUIRE(planes == 1 || planes == 2, "pack: planes must be 1 or 2");
  const long long n = N * d_pad;
  int blocks = (int)((n + 255) / 256);
  if (blocks > 148 * 32) blocks = 148 * 32;
  pack_w_kernel<<<blocks, 256, 0, stream>>>(W_enc, N, d, d_pad, planes, reinterpret_cast<__nv_bfloat16*>(w_planes));
  SAEB_CHECK_CUDA(cudaGetLastError());
  fold_bias_kernel<<<(int)((N + 7) / 8), 256, 0, stream>>>(W_enc, b_enc, b_dec, N, d, bias);
  SAEB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int split_x_launch(const void* x, int x_dtype, long long T, long long d, long long ld_x, long long d_pad, int planes,
                   void* out, cudaStream_t stream) {
  const long long total = T * d_pad;
  int blocks = (int)((total + 255) / 256);
  if (blocks > 148 * 32) blocks = 148 * 32;
  __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(out);
  if (x_dtype == DT_F32)
    split_x_kernel<float><<<blocks, 256, 0, stream>>>(reinterpret_cast<const float*>(x), T, d, ld_x, d_pad, planes, o);
  else if (x_dtype == DT_F16)
    split_x_kernel<__half><<<blocks, 256, 0, stream>>>(reinterpret_cast<const __half*>(x), T, d, ld_x, d_pad, planes, o);
  else if (x_dtype == DT_BF16)
    split_x_kernel<__nv_bfloat16><<<blocks, 256, 0, stream>>>(reinterpret_cast<const __nv_bfloat16*>(x), T, d, ld_x,
                                                              d_pad, planes, o);
  else {
    set_error("split_x: unsupported dtype %d", x_dtype);
    return -1;
  }
  SAEB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

// ---------------------------------------------------------------------------------------------
// mode 3 ("fp16 + refine"): one fp16 plane of W scaled by a power of two, per-feature norms, folded bias.
// trailer = { w_unscale = 2^-s, wnorm_max }.  No host synchronisation: the scale lives in device memory.
// ---------------------------------------------------------------------------------------------
int pack_weights_f16_launch(const float* W_enc, const float* b_enc, const float* b_dec, long long N, long long d,
                            long long d_pad, void* w_plane, float* bias, float* wnorm, float* dnorm, float* trailer,
                            cudaStream_t stream) {
  SAEB_REQUIRE(N > 0 && d > 0, "pack: need N>0 and d>0 (got N=%lld d=%lld)", N, d);
  // trailer[0] = w_unscale, trailer[1] = max ||w_j||, trailer[2] = |W|max (scratch bits), trailer[3] = max dnorm_j
  SAEB_CHECK_CUDA(cudaMemsetAsync(trailer, 0, 16, stream));
  w_stats_kernel<<<(int)((N + 7) / 8), 256, 0, stream>>>(W_enc, b_enc, b_dec, N, d, bias, wnorm,
                                                         reinterpret_cast<unsigned int*>(trailer + 2),
                                                         reinterpret_cast<unsigned int*>(trailer + 1));
  SAEB_CHECK_CUDA(cudaGetLastError());
  pack_w_f16_kernel<<<(int)((N + 7) / 8), 256, 0, stream>>>(W_enc, N, d, d_pad,
                                                           reinterpret_cast<unsigned int*>(trailer + 2),
                                                           reinterpret_cast<__half*>(w_plane), dnorm, trailer);
  SAEB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

// mode 4 = mode 3 + the residual plane: lo = fp16((W * scale - hi) * 2^11) with the hi plane's power-of-two scale (read
// back from trailer[0], written by pack_w_f16_kernel earlier in the stream); hi + lo / 2^11 reproduces W * scale to
// 2^-22 relative.  The fp16 rounding error of a value below 2^14 is at most 4, so the scaled residual stays below 2^13.
int pack_weights_lo_launch(const float* W_enc, long long N, long long d, long long d_pad, const float* trailer,
                           void* lo_plane, cudaStream_t stream) {
  const long long total = N * d_pad;
  long long blocks = (total + 255) / 256;
  if (blocks > 148 * 32) blocks = 148 * 32;
  pack_w_lo_kernel<<<(int)blocks, 256, 0, stream>>>(W_enc, N, d, d_pad, trailer, reinterpret_cast<__half*>(lo_plane));
  SAEB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

// activations -> one fp16 plane [T][d_pad], each row scaled by a power of two so that its largest element lands in
// [2^13, 2^14); row_scale[t] undoes it; xnorm[t] >= ||x_t||_2 (of the original activations).  bf16 / fp16 inputs are
// represented exactly (up to fp16 underflow 2^-28 below the row maximum); fp32 inputs are rounded to 11 bits.
int prep_x_f16_launch(const void* x, int x_dtype, long long T, long long d, long long ld_x, long long d_pad, void* out,
                      float* row_scale, float* xnorm, float* xdnorm, cudaStream_t stream) {
  const int wpb = 8;
  const unsigned blocks = (unsigned)((T + wpb - 1) / wpb);
  __half* o = reinterpret_cast<__half*>(out);
  if (x_dtype == DT_F32)
    SAEB_CARVEOUT(prep_x_f16_kernel<float>), prep_x_f16_kernel<float><<<blocks, wpb * 32, 0, stream>>>(reinterpret_cast<const float*>(x), T, d, ld_x, d_pad, o,
                                                             row_scale, xnorm, xdnorm);
  else if (x_dtype == DT_F16)
    SAEB_CARVEOUT(prep_x_f16_kernel<__half>), prep_x_f16_kernel<__half><<<blocks, wpb * 32, 0, stream>>>(reinterpret_cast<const __half*>(x), T, d, ld_x, d_pad, o,
                                                              row_scale, xnorm, xdnorm);
  else if (x_dtype == DT_BF16)
    SAEB_CARVEOUT(prep_x_f16_kernel<__nv_bfloat16>), prep_x_f16_kernel<__nv_bfloat16><<<blocks, wpb * 32, 0, stream>>>(reinterpret_cast<const __nv_bfloat16*>(x), T, d,
                                                                     ld_x, d_pad, o, row_scale, xnorm, xdnorm);
  else {
    set_error("prep_x: unsupported dtype %d", x_dtype);
    return -1;
  }
  SAEB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace saeb
