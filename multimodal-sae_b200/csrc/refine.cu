// Exact refinement of an approximate TopK ("fp16 + refine" mode) and the exact dense fallback.
//
// The fused GEMM runs ONE tensor-core pass with W_enc rounded to fp16 (11-bit significand) and returns, per row, the
// K2 = k + margin best candidates by APPROXIMATE value a_j.  With x = xh + dx, w_j = wh_j + dw_j (h = what the tensor
// cores saw) the error of candidate j is  xh.dw_j + dx.wh_j + dx.dw_j  plus fp32 accumulation error, hence bounded by
//   eps_j = 1.001 * (||x|| * ||dw_j|| + ||dx|| * ||w_j||) + slack * ||x|| * ||w_j||        (Cauchy-Schwarz)
// where ||dw_j|| is the EXACT norm of the feature's fp16 rounding error (computed once at pack time, ~0.4 * 2^-11
// ||w_j||), ||dx|| the exact norm of the activation row's rounding error (0 for bf16 / fp16 inputs) and slack = 2^-14
// covers the accumulation (256 block additions of <= 2^-23 relative error each, doubled).  Hence
//   * the true k-th value is at least L = k-th largest of (a_j - eps_j);
//   * only candidates with a_j + eps_j >= L can belong to the true TopK; they are re-evaluated EXACTLY here
//     (fp32 dot product with the fp32 W_enc row + folded bias), and the final TopK is taken over the exact values.
//     Two stages: first the k best by a_j; the smallest of their exact values replaces L (it is a lower bound of
//     the k-th value that does not sit a full eps below it), which roughly halves the rest of the work;
//   * a non-candidate has a <= a_last (the smallest kept approximation); if a_last + c_eps*||x||*max_j||w_j|| >= L
//     the candidate list might be too short: the row is FLAGGED and recomputed by the exact dense kernels below.
// Output values are fp32-exact like the reference's (sae/sae.py:172-181), the index set is the reference's up to
// fp32 summation noise; the gather is HBM-bound (about (k + 20) fp32 rows per token).
#include <type_traits>

#include "kernels_refine.cuh"

namespace saeb {

unsigned long long* stats_ptr();   // encode_topk.cu: device counters of option "stats" (nullptr when off)

int candidate_bounds_launch(const float* cand_vals, const long long* cand_idx, long long T, int K2, int k,
                            const float* wnorm, const float* dnorm, const float* xnorm, const float* xdnorm,
                            float c_eps, long long clamp_feature, float* lb_out, float* ub_out, cudaStream_t stream) {
  if (T == 0) return 0;
  SAEB_CARVEOUT(candidate_bounds_kernel);
  candidate_bounds_kernel<<<(unsigned)T, 128, (size_t)2 * K2 * sizeof(float), stream>>>(
      cand_vals, cand_idx, K2, k, wnorm, dnorm, xnorm, xdnorm, c_eps, clamp_feature, lb_out, ub_out);
  SAEB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

// ---------------------------------------------------------------------------------------------
// host launchers
// ---------------------------------------------------------------------------------------------
// threads per refinement CTA in feature-sharded calls (ext_lower given): a shard evaluates only ~k/R + a few candidates
// per token, so smaller blocks keep more tokens in flight when the kernel has the GPU to itself (128, the measured
// default); beside a persistent GEMM grid only one CTA fits per SM and 256 threads keep more loads in flight
static thread_local int g_refine_threads_sharded = 128;
int set_refine_threads(int v) {
  if (v < 32 || v > RF_THREADS || (v & 31)) {
    set_error("refine_threads must be a multiple of 32 in 32..%d", RF_THREADS);
    return -1;
  }
  g_refine_threads_sharded = v;
  return 0;
}

// option "scan_warp" (default 1): feature-sharded scan calls of the refinement use the warp-per-token kernel
static thread_local int g_scan_warp = 1;
int set_scan_warp(int v) {
  g_scan_warp = v != 0;
  return 0;
}
int scan_warp_enabled() { return g_scan_warp; }

size_t refine_fallback_bytes(long long N) { return (size_t)RF_MAX_FLAG * (size_t)N * sizeof(float) + 1024; }

template <typename XT>
static int refine_launch_t(const XT* x, long long T, long long ld_x, const float* W, long long d, long long N,
                           const float* bias, const float* wnorm, const float* dnorm, const float* trailer,
                           const float* xnorm, const float* xdnorm, float c_eps,
                           const float* cand_vals, const long long* cand_idx, int K2, int k, long long clamp_feature,
                           float clamp_value, float* out_vals, long long* out_idx, int* status, int* flag_rows,
                           float* dense_scratch, const float* ext_lower, const __half* w_lo, long long ld_w,
                           int max_ctas, int value_mode, const float* ext_upper, const float* feat_thr,
                           float* out_member, cudaStream_t stream) {
  // feature-sharded calls evaluate only a handful of candidates per token: smaller blocks, more tokens in flight
  const int threads = ext_lower != nullptr ? g_refine_threads_sharded : RF_THREADS;
  // max_ctas > 0: persistent grid (tokens walked with stride gridDim.x), sized by the caller so that a fixed number of
  // CTAs per SM rides beside a resident GEMM grid; 0: one CTA per token
  const unsigned grid = (unsigned)((max_ctas > 0 && max_ctas < T) ? max_ctas : T);
  bool use_lo = false;
  if constexpr (!std::is_same<XT, float>::value) {
    use_lo = w_lo != nullptr;
  }
  if constexpr (!std::is_same<XT, float>::value) if (use_lo) {
    // residual-plane correction (packed mode 4); fp32 activations are rounded on their way into the tensor cores,
    // which the residual of W cannot correct, so they always take the exact route below
    const int d8 = (int)((d + 7) & ~7ll);
    const size_t smem = (size_t)(d8 + 6 * K2) * sizeof(float);
    SAEB_REQUIRE(smem <= 200 * 1024, "refine: d=%lld too large for the shared-memory row buffer", d);
    SAEB_REQUIRE(ld_w >= d8 && (reinterpret_cast<uintptr_t>(w_lo) & 15) == 0 && ld_w % 8 == 0,
                 "refine: residual plane must be 16-byte aligned with rows padded to a multiple of 8");
    auto kern = refine_lo_kernel<XT>;
    SAEB_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    SAEB_CARVEOUT(kern);
    kern<<<grid, threads, smem, stream>>>(x, ld_x, w_lo, ld_w, d, N, bias, wnorm, dnorm, trailer, xnorm, xdnorm, c_eps,
                                            cand_vals, cand_idx, K2, k, clamp_feature, clamp_value, out_vals, out_idx,
                                            status, flag_rows, ext_lower, T, stats_ptr(), value_mode, ext_upper, feat_thr,
                                            out_member);
  }
  // feature-sharded scan calls (value_mode 2 with both external bounds): warp-per-token kernel without shared
  // memory, and fallback grids without dynamic shared memory -- every launch of the call fits beside a resident GEMM CTA
  const bool scan_warp = !use_lo && g_scan_warp && value_mode == 2 && ext_lower != nullptr && ext_upper != nullptr &&
                         out_member != nullptr && K2 <= 32 * RSW_SLOTS;
  if (scan_warp) {
    const int vec_ok = ((d & 3) == 0) && ((reinterpret_cast<uintptr_t>(W) & 15) == 0) &&
                       ((reinterpret_cast<uintptr_t>(x) & (4 * sizeof(XT) - 1)) == 0) && ((ld_x & 3) == 0);
    const int wpb = RSW_THREADS / 32;
    long long blocks = (T + wpb - 1) / wpb;
    if (max_ctas > 0 && max_ctas < blocks) blocks = max_ctas;
    SAEB_CARVEOUT(refine_scan_warp_kernel<XT>);
    refine_scan_warp_kernel<XT><<<(unsigned)blocks, RSW_THREADS, 0, stream>>>(
        x, ld_x, W, d, N, bias, wnorm, dnorm, trailer, xnorm, xdnorm, c_eps, cand_vals, cand_idx, K2, k, clamp_feature,
        clamp_value, out_vals, out_idx, status, flag_rows, ext_lower, T, stats_ptr(), ext_upper, feat_thr, out_member,
        vec_ok);
  }
  if (!use_lo && !scan_warp) {
    const int d4 = (int)((d + 3) & ~3ll);
    const size_t smem = (size_t)(d4 + 6 * K2) * sizeof(float);
    SAEB_REQUIRE(smem <= 200 * 1024, "refine: d=%lld too large for the shared-memory row buffer", d);
    auto kern = refine_kernel<XT>;
    SAEB_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    SAEB_CARVEOUT(kern);
    kern<<<grid, threads, smem, stream>>>(x, ld_x, W, d, N, bias, wnorm, dnorm, trailer, xnorm, xdnorm, c_eps,
                                            cand_vals, cand_idx, K2, k, clamp_feature, clamp_value, out_vals, out_idx,
                                            status, flag_rows, ext_lower, T, stats_ptr(), value_mode, ext_upper, feat_thr,
                                            out_member);
  }
  SAEB_CHECK_CUDA(cudaGetLastError());
  // exact dense fallback for the (normally zero) flagged rows; the grids exit at once when nothing is flagged
  auto ek = exact_rows_kernel<XT>;
  const int stage_x = scan_warp ? 0 : 1;
  const size_t esmem = stage_x ? (size_t)d * sizeof(float) : 0;
  SAEB_CHECK_CUDA(cudaFuncSetAttribute(ek, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)esmem));
  dim3 eg(scan_warp ? 16 : 128, RF_MAX_FLAG);
  SAEB_CARVEOUT(ek);
  ek<<<eg, 256, esmem, stream>>>(x, ld_x, W, d, N, bias, status, flag_rows, clamp_feature, clamp_value,
                                 dense_scratch, stage_x);
  SAEB_CHECK_CUDA(cudaGetLastError());
  int kp2 = 2;
  while (kp2 < k) kp2 <<= 1;
  // beside a resident GEMM grid only small blocks can be scheduled (registers): the fallback grids, which exit at once
  // when nothing is flagged, must never make the stream wait for a GEMM launch boundary
  const int fb_threads = (max_ctas > 0 || scan_warp) ? 256 : 1024;
  SAEB_CARVEOUT(dense_topk_kernel);
  dense_topk_kernel<<<RF_MAX_FLAG, fb_threads, (size_t)kp2 * sizeof(uint2), stream>>>(
      dense_scratch, N, N, k, status, RF_MAX_FLAG, flag_rows, out_vals, out_idx, out_member);
  SAEB_CHECK_CUDA(cudaGetLastError());
  auto ok = overflow_rows_kernel<XT>;
  const size_t osmem = (size_t)kp2 * sizeof(uint2) + (stage_x ? (size_t)d * sizeof(float) : 0);
  SAEB_CHECK_CUDA(cudaFuncSetAttribute(ok, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)osmem));
  SAEB_CARVEOUT(ok);
  ok<<<RF_MAX_FLAG, fb_threads, osmem, stream>>>(x, ld_x, W, d, N, bias, status, flag_rows, clamp_feature, clamp_value,
                                           dense_scratch, k, out_vals, out_idx, out_member, stage_x);
  SAEB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int refine_launch(const void* x, int x_dtype, long long T, long long ld_x, const float* W, long long d, long long N,
                  const float* bias, const float* wnorm, const float* dnorm, const float* trailer, const float* xnorm,
                  const float* xdnorm, float c_eps,
                  const float* cand_vals, const long long* cand_idx, int K2, int k, long long clamp_feature,
                  float clamp_value, float* out_vals, long long* out_idx, int* status, int* flag_rows,
                  float* dense_scratch, const float* ext_lower, const void* w_lo, long long ld_w, int max_ctas,
                  int value_mode, const float* ext_upper, const float* feat_thr, float* out_member,
                  cudaStream_t stream) {
#define SAEB_RF(XT)                                                                                                  \
  return refine_launch_t<XT>(reinterpret_cast<const XT*>(x), T, ld_x, W, d, N, bias, wnorm, dnorm, trailer, xnorm,   \
                             xdnorm, c_eps,                                                                          \
                             cand_vals, cand_idx, K2, k, clamp_feature, clamp_value, out_vals, out_idx, status,      \
                             flag_rows, dense_scratch, ext_lower, reinterpret_cast<const __half*>(w_lo), ld_w,       \
                             max_ctas, value_mode, ext_upper, feat_thr, out_member, stream)
  if (x_dtype == DT_F32) SAEB_RF(float);
  if (x_dtype == DT_BF16) SAEB_RF(__nv_bfloat16);
  if (x_dtype == DT_F16) SAEB_RF(__half);
#undef SAEB_RF
  set_error("refine: unsupported x dtype %d", x_dtype);
  return -1;
}

int dense_topk_launch(const float* dense, long long T, long long ld, long long N, int k, float* out_vals,
                      long long* out_idx, cudaStream_t stream) {
  SAEB_REQUIRE(T >= 0 && N >= 1 && k >= 1 && k <= N && k <= 4096, "dense_topk: bad arguments");
  if (T == 0) return 0;
  int kp2 = 2;
  while (kp2 < k) kp2 <<= 1;
  dense_topk_kernel<<<(unsigned)T, 1024, (size_t)kp2 * sizeof(uint2), stream>>>(dense, ld, N, k, nullptr, 0, nullptr,
                                                                               out_vals, out_idx, nullptr);
  SAEB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace saeb
