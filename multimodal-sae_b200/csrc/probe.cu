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

}  // namespace saeb
