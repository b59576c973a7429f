// Launchers of the probing queries (reference tools/probe_activations.py:109-126): column sums of dense latents and
// exact activation maps of a few selected features.  Device code: kernels_probe.cuh.
#include "kernels_probe.cuh"

namespace saeb {

int column_sums_launch(const float* dense, long long T, long long ld, long long N, double* colsum, cudaStream_t stream) {
  SAEB_REQUIRE(T >= 0 && N >= 1 && ld >= N, "column_sums: bad shape");
  if (T == 0) return 0;
  dim3 grid((unsigned)((N + CS_THREADS - 1) / CS_THREADS), (unsigned)((T + CS_ROWS - 1) / CS_ROWS));
  column_sums_kernel<<<grid, CS_THREADS, 0, stream>>>(dense, T, ld, N, colsum);
  SAEB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int feature_maps_launch(const void* x, int x_dtype, long long T, long long ld_x, const float* W, const float* b_enc,
                        const float* b_dec, long long d, long long N, const long long* sel, int n_sel, float* out,
                        int* err_flag, cudaStream_t stream) {
  SAEB_REQUIRE(T >= 0 && d >= 1 && N >= 1 && n_sel >= 0, "feature_maps: bad shape");
  if (T == 0 || n_sel == 0) return 0;
  const size_t smem = (size_t)2 * d * sizeof(float);
  SAEB_REQUIRE(smem <= 200 * 1024, "feature_maps: d=%lld too large for the shared-memory row buffer", d);
  dim3 grid((unsigned)n_sel, (unsigned)((T + FM_TOKENS - 1) / FM_TOKENS));
#define SAEB_FM(XT)                                                                                              \
  do {                                                                                                           \
    auto kern = feature_maps_kernel<XT>;                                                                         \
    SAEB_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));         \
    kern<<<grid, FM_THREADS, smem, stream>>>(reinterpret_cast<const XT*>(x), T, ld_x, W, b_enc, b_dec, d, N, sel, \
                                             out, err_flag);                                                     \
  } while (0)
  if (x_dtype == DT_F32) SAEB_FM(float);
  else if (x_dtype == DT_BF16) SAEB_FM(__nv_bfloat16);
  else if (x_dtype == DT_F16) SAEB_FM(__half);
  else {
    set_error("feature_maps: unsupported x dtype %d", x_dtype);
    return -1;
  }
#undef SAEB_FM
  SAEB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

// ---- diagnostics: a synthetic co-resident load (what slows a GEMM launch that shares its SMs with small kernels?) ----
// mode 0: pure ALU (dependent FMA chains, no memory); 1: streaming 16-byte reads of `buf` (HBM / L2 pollution);
// 2: reads confined to the first 32 MB of `buf` (L2 hits: L2 bandwidth without DRAM traffic).  `ctas` CTAs of 128
// threads, no shared memory, `iters` inner iterations per thread.
__global__ void __launch_bounds__(128)
coload_kernel(int mode, long long iters, const uint4* __restrict__ buf, long long n_vec, float* __restrict__ sink) {
  const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long nthr = (long long)gridDim.x * blockDim.x;
  // bits 4..7 of `mode`: only warps whose (warp id & 3) is in that mask work (0 = all): which SM sub-partition hosts
  // the load?  bit 8: use the hardware warp slot %warpid instead of the CTA-local warp id
  const int wmask = (mode >> 4) & 15;
  mode &= 15 | 256;
  if (wmask) {
    unsigned wid = threadIdx.x >> 5;
    if (mode & 256) asm volatile("mov.u32 %0, %%warpid;" : "=r"(wid));
    if (!((wmask >> (wid & 3)) & 1)) return;
  }
  mode &= 15;
  float acc0 = (float)tid, acc1 = 1.f, acc2 = 2.f, acc3 = 3.f;
  if (mode == 0) {
    for (long long i = 0; i < iters; ++i) {
      acc0 = fmaf(acc0, 1.0001f, 0.5f);
      acc1 = fmaf(acc1, 0.9999f, 0.25f);
      acc2 = fmaf(acc2, 1.0002f, 0.125f);
      acc3 = fmaf(acc3, 0.9998f, 0.0625f);
    }
  } else {
    const long long span = (mode == 2 && n_vec > (32ll << 20) / 16) ? (32ll << 20) / 16 : n_vec;
    long long j = tid % span;
    for (long long i = 0; i < iters; ++i) {
      const uint4 v = __ldcg(buf + j);
      acc0 += __uint_as_float(v.x & 0x3fffffffu);
      acc1 += __uint_as_float(v.y & 0x3fffffffu);
      j += nthr;
      if (j >= span) j -= span;
    }
  }
  if (acc0 + acc1 + acc2 + acc3 == 123.456f) sink[0] = acc0;
}

int coload_launch(int mode, int ctas, long long iters, const void* buf, size_t bytes, float* sink, cudaStream_t stream) {
  SAEB_REQUIRE(mode >= 0 && (mode & 15) <= 2 && ctas >= 1 && iters >= 0 && sink != nullptr, "coload: bad arguments");
  SAEB_REQUIRE((mode & 15) == 0 || (buf != nullptr && bytes >= 16 * 128), "coload: modes 1 / 2 need a buffer");
  SAEB_CARVEOUT(coload_kernel);
  coload_kernel<<<ctas, 128, 0, stream>>>(mode, iters, reinterpret_cast<const uint4*>(buf), (long long)(bytes / 16), sink);
  SAEB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace saeb
