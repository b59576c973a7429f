// Device code only (no launch syntax), like the other kernels_*.cuh files.
//
// "Which features fire on this input?" queries of the reference's probing tool (tools/probe_activations.py:109-126):
//   latents = sae.pre_acts(hidden)                       dense relu latents [T, N]
//   top     = latents.mean(dim=0).topk(k).indices        features with the largest MEAN activation over the tokens
//   maps    = latents[:, top]                            per-token activation of those features
// The dense latents are produced chunk by chunk by the fused GEMM (dense store) and reduced at once by
// column_sums_kernel; the maps of the few selected features are re-evaluated exactly in fp32 by feature_maps_kernel,
// so the [T, N] tensor never has to exist as a whole.
#pragma once
#include "common.cuh"

namespace saeb {

constexpr int CS_THREADS = 256;
constexpr int CS_ROWS = 64;

// colsum[n] += sum over this block's rows of dense[t][n]   (fp64 accumulation; blockIdx.x = column block, .y = row block)
__global__ void __launch_bounds__(CS_THREADS)
column_sums_kernel(const float* __restrict__ dense, long long T, long long ld, long long N,
                   double* __restrict__ colsum) {
  const long long n = (long long)blockIdx.x * CS_THREADS + threadIdx.x;
  if (n >= N) return;
  const long long t0 = (long long)blockIdx.y * CS_ROWS;
  const long long t1 = (t0 + CS_ROWS < T) ? t0 + CS_ROWS : T;
  double acc = 0.0;
  for (long long t = t0; t < t1; ++t) acc += (double)dense[t * ld + n];
  atomicAdd(colsum + n, acc);
}

constexpr int FM_THREADS = 256;
constexpr int FM_TOKENS = 64;   // tokens per block

// out[j][t] = relu((x_t - b_dec) . W[sel_j] + b_enc[sel_j])   exact fp32, the reference's operation order
// (sae/sae.py:174-177: subtract b_dec first, then the linear layer).  blockIdx.x = selected feature, .y = token block.
template <typename XT>
__global__ void __launch_bounds__(FM_THREADS)
feature_maps_kernel(const XT* __restrict__ x, long long T, long long ld_x, const float* __restrict__ W,
                    const float* __restrict__ b_enc, const float* __restrict__ b_dec, long long d, long long N,
                    const long long* __restrict__ sel, float* __restrict__ out, int* __restrict__ err_flag) {
  extern __shared__ float fsm[];   // [d] weight row | [d] b_dec
  float* ws = fsm;
  float* bs = fsm + d;
  const long long f = sel[blockIdx.x];
  if (f < 0 || f >= N) {
    if (err_flag != nullptr && threadIdx.x == 0) atomicExch(err_flag, 1);
    return;
  }
  for (long long i = threadIdx.x; i < d; i += blockDim.x) {
    ws[i] = W[f * d + i];
    bs[i] = b_dec[i];
  }
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  const long long t0 = (long long)blockIdx.y * FM_TOKENS;
  const long long t1 = (t0 + FM_TOKENS < T) ? t0 + FM_TOKENS : T;
  const float be = b_enc[f];
  for (long long t = t0 + warp; t < t1; t += nw) {
    const XT* xr = x + t * ld_x;
    float acc = 0.f;
    for (long long i = lane; i < d; i += 32) acc = fmaf((float)xr[i] - bs[i], ws[i], acc);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) out[(long long)blockIdx.x * T + t] = fmaxf(acc + be, 0.f);
  }
}

}  // namespace saeb
