// Backward of the sparse decode (the autograd contract of the reference's decoder seam, TritonDecoder.backward,
// sae/kernels.py:403-429):
//   d_vals[t, j] = grad_out[t, :] . W_dec[idx[t, j], :]                       replaces triton_dense_dense_sparseout_matmul
//                                                                              (sae/kernels.py:287-400)
//   dW_dec[n, :] += sum_{(t, j): idx[t, j] == n} vals[t, j] * grad_out[t, :]   replaces triton_sparse_transpose_dense_matmul
//                                                                              (sae/kernels.py:10-175; zero values skipped)
// Both are row gathers / scatters of k fp32 rows per token, i.e. HBM / L2-atomic bound like the forward decode.
// One CTA per token; the token's gradient row is staged in shared memory as fp32.
#include "kernels_decode_bwd.cuh"

namespace saeb {

int decode_bwd_acts_launch(const float* grad_out, long long ld_g, const long long* idx, long long T, int k,
                           const float* W_dec, long long d, long long N, float* d_vals, int* err_flag,
                           cudaStream_t stream) {
  SAEB_REQUIRE(T >= 0 && k >= 1 && d >= 1 && N >= 1 && ld_g >= d, "decode_backward_acts: bad arguments");
  if (T == 0) return 0;
  const size_t smem = (size_t)((d + 3) & ~3ll) * sizeof(float);
  SAEB_REQUIRE(smem <= 200 * 1024, "decode_backward_acts: d=%lld too large for the shared-memory row buffer", d);
  SAEB_CHECK_CUDA(cudaFuncSetAttribute(decode_bwd_acts_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  decode_bwd_acts_kernel<<<(unsigned)T, DBW_THREADS, smem, stream>>>(grad_out, ld_g, idx, k, W_dec, d, N, d_vals,
                                                                    err_flag);
  SAEB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int decode_bwd_weight_launch(const float* grad_out, long long ld_g, const long long* idx, const float* vals,
                             long long T, int k, long long d, long long N, float* dW, int* err_flag,
                             cudaStream_t stream) {
  SAEB_REQUIRE(T >= 0 && k >= 1 && d >= 1 && N >= 1 && ld_g >= d, "decode_backward_weight: bad arguments");
  if (T == 0) return 0;
  const size_t smem = (size_t)((d + 3) & ~3ll) * sizeof(float);
  SAEB_REQUIRE(smem <= 200 * 1024, "decode_backward_weight: d=%lld too large for the shared-memory row buffer", d);
  SAEB_CHECK_CUDA(cudaFuncSetAttribute(decode_bwd_weight_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  decode_bwd_weight_kernel<<<(unsigned)T, DBW_THREADS, smem, stream>>>(grad_out, ld_g, idx, vals, k, d, N, dW,
                                                                      err_flag);
  SAEB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace saeb
