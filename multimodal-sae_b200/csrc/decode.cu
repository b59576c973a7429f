// Sparse decode: out[t,:] = sum_j vals[t,j] * W_dec[idx[t,j], :] + b_dec   (HBM-bound row gather)
//
// Replaces decoder_impl (sae/utils.py:107-129): the Triton kernel triton_sparse_dense_matmul_kernel
// (sae/kernels.py:222-284; fp32 accumulate in j order, zero values skipped) and the eager fallback
// (sae/utils.py:108-111, a second dense [T,N]x[N,d] GEMM), plus the `+ b_dec` of Sae.decode (sae/sae.py:191),
// the output cast of the steering hook (features/steering.py:116-118) and the residual sum of squares of
// Sae.forward (sae/sae.py:201,229).
//
// One CTA per token.  The k selected rows (16 KB each at d=4096 fp32) are streamed with 16-byte read-only loads,
// 8 rows in flight per thread; algorithmic traffic per token = k * d * sizeof(W) + d * sizeof(out) + 12 k.
#include "kernels_decode.cuh"

namespace saeb {

template <typename WT, typename OT>
static int decode_dispatch_x(const long long* idx, const float* vals, long long T, int k, const WT* W, long long d,
                             long long N, const float* b_dec, OT* out, long long ld_out, const void* x, int x_dtype,
                             long long ld_x, double* sq_err, int* err_flag, int max_ctas, cudaStream_t stream) {
  // max_ctas > 0: persistent grid (tokens walked with stride gridDim.x) so that a fixed number of CTAs per SM rides
  // beside a resident GEMM grid; 0: one CTA per token
  dim3 grid((unsigned)((max_ctas > 0 && max_ctas < T) ? max_ctas : T)), block(DEC_THREADS);
  const bool with_x = sq_err != nullptr && x != nullptr;
  const bool vec = d % 4 == 0 && ld_out % 4 == 0 && (!with_x || ld_x % 4 == 0) &&
                   (reinterpret_cast<uintptr_t>(W) & 15) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0 &&
                   (!with_x || (reinterpret_cast<uintptr_t>(x) & 15) == 0) &&
                   (b_dec == nullptr || (reinterpret_cast<uintptr_t>(b_dec) & 15) == 0);
#define SAEB_DEC_LAUNCH(XT, XP, LDX, SQ)                                                                            \
  do {                                                                                                              \
    if (vec) {                                                                                                      \
      SAEB_CARVEOUT((decode_kernel<WT, OT, XT>));                                                                    \
      decode_kernel<WT, OT, XT><<<grid, block, 0, stream>>>(idx, vals, k, W, d, N, b_dec, out, ld_out, XP, LDX, SQ,  \
                                                            err_flag, T);                                           \
    } else                                                                                                          \
      decode_scalar_kernel<WT, OT, XT><<<grid, block, 0, stream>>>(idx, vals, k, W, d, N, b_dec, out, ld_out, XP,    \
                                                                   LDX, SQ, err_flag, T);                           \
  } while (0)
  if (!with_x)
    SAEB_DEC_LAUNCH(float, (const float*)nullptr, 0, (double*)nullptr);
  else if (x_dtype == DT_F32)
    SAEB_DEC_LAUNCH(float, reinterpret_cast<const float*>(x), ld_x, sq_err);
  else if (x_dtype == DT_BF16)
    SAEB_DEC_LAUNCH(__nv_bfloat16, reinterpret_cast<const __nv_bfloat16*>(x), ld_x, sq_err);
  else if (x_dtype == DT_F16)
    SAEB_DEC_LAUNCH(__half, reinterpret_cast<const __half*>(x), ld_x, sq_err);
  else {
    set_error("decode: unsupported x dtype %d", x_dtype);
    return -1;
  }
  SAEB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

template <typename WT>
static int decode_dispatch_o(const long long* idx, const float* vals, long long T, int k, const WT* W, long long d,
                             long long N, const float* b_dec, void* out, int out_dtype, long long ld_out,
                             const void* x, int x_dtype, long long ld_x, double* sq_err, int* err_flag, int max_ctas,
                             cudaStream_t stream) {
  if (out_dtype == DT_F32)
    return decode_dispatch_x<WT, float>(idx, vals, T, k, W, d, N, b_dec, reinterpret_cast<float*>(out), ld_out, x,
                                        x_dtype, ld_x, sq_err, err_flag, max_ctas, stream);
  if (out_dtype == DT_F16)
    return decode_dispatch_x<WT, __half>(idx, vals, T, k, W, d, N, b_dec, reinterpret_cast<__half*>(out), ld_out, x,
                                         x_dtype, ld_x, sq_err, err_flag, max_ctas, stream);
  if (out_dtype == DT_BF16)
    return decode_dispatch_x<WT, __nv_bfloat16>(idx, vals, T, k, W, d, N, b_dec, reinterpret_cast<__nv_bfloat16*>(out),
                                                ld_out, x, x_dtype, ld_x, sq_err, err_flag, max_ctas, stream);
  set_error("decode: unsupported output dtype %d", out_dtype);
  return -1;
}

int decode_launch(const long long* idx, const float* vals, long long T, int k, const void* W_dec, int w_dtype,
                  long long d, long long N, const float* b_dec, void* out, int out_dtype, long long ld_out,
                  const void* x, int x_dtype, long long ld_x, double* sq_err, int* err_flag, int max_ctas,
                  cudaStream_t stream) {
  if (T == 0) return 0;
  SAEB_REQUIRE(k >= 1 && k <= DEC_KMAX, "decode: k=%d out of range (1..%d)", k, DEC_KMAX);
  SAEB_REQUIRE(d >= 1 && ld_out >= d, "decode: bad d / ld_out");
  if (w_dtype == DT_F32)
    return decode_dispatch_o<float>(idx, vals, T, k, reinterpret_cast<const float*>(W_dec), d, N, b_dec, out,
                                    out_dtype, ld_out, x, x_dtype, ld_x, sq_err, err_flag, max_ctas, stream);
  if (w_dtype == DT_BF16)
    return decode_dispatch_o<__nv_bfloat16>(idx, vals, T, k, reinterpret_cast<const __nv_bfloat16*>(W_dec), d, N,
                                            b_dec, out, out_dtype, ld_out, x, x_dtype, ld_x, sq_err, err_flag, max_ctas, stream);
  if (w_dtype == DT_F16)
    return decode_dispatch_o<__half>(idx, vals, T, k, reinterpret_cast<const __half*>(W_dec), d, N, b_dec, out,
                                     out_dtype, ld_out, x, x_dtype, ld_x, sq_err, err_flag, max_ctas, stream);
  set_error("decode: unsupported W_dec dtype %d", w_dtype);
  return -1;
}

int total_variance_launch(const void* x, int x_dtype, long long T, long long d, long long ld_x, double* scratch,
                          double* out, cudaStream_t stream) {
  SAEB_REQUIRE(T > 0 && d > 0, "total_variance: empty input");
  SAEB_CHECK_CUDA(cudaMemsetAsync(scratch, 0, sizeof(double) * 2 * d, stream));
  const int rpb = 64;
  const int blocks = (int)((T + rpb - 1) / rpb);
  if (x_dtype == DT_F32)
    colstats_kernel<float><<<blocks, 256, 0, stream>>>(reinterpret_cast<const float*>(x), T, d, ld_x, rpb, scratch,
                                                      scratch + d);
  else if (x_dtype == DT_BF16)
    colstats_kernel<__nv_bfloat16><<<blocks, 256, 0, stream>>>(reinterpret_cast<const __nv_bfloat16*>(x), T, d, ld_x,
                                                              rpb, scratch, scratch + d);
  else if (x_dtype == DT_F16)
    colstats_kernel<__half><<<blocks, 256, 0, stream>>>(reinterpret_cast<const __half*>(x), T, d, ld_x, rpb, scratch,
                                                       scratch + d);
  else {
    set_error("total_variance: unsupported dtype %d", x_dtype);
    return -1;
  }
  SAEB_CHECK_CUDA(cudaGetLastError());
  totvar_kernel<<<1, 256, 0, stream>>>(scratch, scratch + d, T, d, out);
  SAEB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace saeb
