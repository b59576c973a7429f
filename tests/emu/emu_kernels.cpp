// TEST INFRASTRUCTURE ONLY -- C entry points that run the device code of multimodal-sae_b200/csrc/kernels_*.cuh on the
// host through tests/emu/cuda_emu.h.  Built by tests/test_kernel_emu.py with g++ from a copy of the kernel headers in
// which `extern __shared__` / `__shared__` were replaced by `extern` / `static`.
#define SAEB_CPU_EMU 1
#include <cstdarg>

#include "kernels_coo_scan.cuh"
#include "kernels_decode.cuh"
#include "kernels_decode_bwd.cuh"
#include "kernels_exchange.cuh"
#include "kernels_kth.cuh"
#include "kernels_pack.cuh"
#include "kernels_refine.cuh"
#include "kernels_topk_select.cuh"

namespace saeb {
void set_error(const char*, ...) {}
// dynamic shared memory of the emulated kernels (`extern __shared__ T name[]` in the sources)
alignas(16) float rsm[1 << 16];
alignas(16) float bsm[1 << 12];
alignas(16) float gsm[1 << 16];
alignas(16) float xsm[1 << 14];
alignas(16) uint2 esm[1 << 14];
alignas(16) uint2 ssm[1 << 15];
alignas(16) uint2 msm[1 << 16];
alignas(16) uint2 dsm[1 << 12];
alignas(16) uint2 osm[1 << 14];
alignas(16) uint32_t hsm[1 << 16];

// compact_row is a device function of the fused tcgen05 kernel's epilogue: run it from a one-warp wrapper
template <int SLOTS>
void compact_row_wrapper(uint2* buf, int cnt_in, int k, float* thr_out, int* cnt_out) {
  static int hist[256];
  float thr = 0.f;
  int cnt = 0;
  compact_row<SLOTS>(buf, cnt_in, k, threadIdx.x & 31u, hist, thr, cnt);
  if (threadIdx.x == 0) {
    *thr_out = thr;
    *cnt_out = cnt;
  }
}
}  // namespace saeb

using namespace saeb;

extern "C" {

void emu_kth(int vpl, const float* g, int R, long long T, int m, int kth, float* out) {
  const unsigned blocks = (unsigned)((T + 7) / 8);
  emu::launch({blocks}, {256}, [&] {
    if (vpl == 4) kth_gathered_reg_kernel<4>(g, R, T, m, kth, out);
    else if (vpl == 16) kth_gathered_reg_kernel<16>(g, R, T, m, kth, out);
    else if (vpl == 64) kth_gathered_reg_kernel<64>(g, R, T, m, kth, out);
    else kth_gathered_mem_kernel(g, R, T, m, kth, out);
  });
}

void emu_decode_bwd_acts(const float* g, long long ld_g, const long long* idx, long long T, int k, const float* W,
                         long long d, long long N, float* d_vals, int* err_flag) {
  emu::launch({(unsigned)T}, {(unsigned)DBW_THREADS},
              [&] { decode_bwd_acts_kernel(g, ld_g, idx, k, W, d, N, d_vals, err_flag); });
}

void emu_decode_bwd_weight(const float* g, long long ld_g, const long long* idx, const float* vals, long long T, int k,
                           long long d, long long N, float* dW, int* err_flag) {
  emu::launch({(unsigned)T}, {(unsigned)DBW_THREADS},
              [&] { decode_bwd_weight_kernel(g, ld_g, idx, vals, k, d, N, dW, err_flag); });
}

// mode 3 / 4 weight pack exactly as pack_weights_f16_launch + pack_weights_lo_launch enqueue it
void emu_pack(const float* W, const float* b_enc, const float* b_dec, long long N, long long d, long long d_pad,
              void* hi, void* lo, float* bias, float* wnorm, float* dnorm, float* trailer) {
  std::memset(trailer, 0, 16);
  const unsigned blocks = (unsigned)((N + 7) / 8);
  emu::launch({blocks}, {256}, [&] {
    w_stats_kernel(W, b_enc, b_dec, N, d, bias, wnorm, reinterpret_cast<unsigned int*>(trailer + 2),
                   reinterpret_cast<unsigned int*>(trailer + 1));
  });
  emu::launch({blocks}, {256}, [&] {
    pack_w_f16_kernel(W, N, d, d_pad, reinterpret_cast<unsigned int*>(trailer + 2), reinterpret_cast<__half*>(hi),
                      dnorm, trailer);
  });
  if (lo != nullptr)
    emu::launch({4}, {256}, [&] { pack_w_lo_kernel(W, N, d, d_pad, trailer, reinterpret_cast<__half*>(lo)); });
}

// bf16 activations (raw 16-bit words) -> scaled fp16 plane + row scale + norms
void emu_prep_x_bf16(const void* x, long long T, long long d, long long ld_x, long long d_pad, void* out,
                     float* row_scale, float* xnorm, float* xdnorm) {
  emu::launch({(unsigned)((T + 7) / 8)}, {256}, [&] {
    prep_x_f16_kernel<__nv_bfloat16>(reinterpret_cast<const __nv_bfloat16*>(x), T, d, ld_x, d_pad,
                                     reinterpret_cast<__half*>(out), row_scale, xnorm, xdnorm);
  });
}
// the same for fp16 activations (the steering path's hidden stream) and fp32 activations (rounded to 11 bits)
void emu_prep_x_f16(const void* x, long long T, long long d, long long ld_x, long long d_pad, void* out,
                    float* row_scale, float* xnorm, float* xdnorm) {
  emu::launch({(unsigned)((T + 7) / 8)}, {256}, [&] {
    prep_x_f16_kernel<__half>(reinterpret_cast<const __half*>(x), T, d, ld_x, d_pad, reinterpret_cast<__half*>(out),
                              row_scale, xnorm, xdnorm);
  });
}
void emu_prep_x_f32(const float* x, long long T, long long d, long long ld_x, long long d_pad, void* out,
                    float* row_scale, float* xnorm, float* xdnorm) {
  emu::launch({(unsigned)((T + 7) / 8)}, {256}, [&] {
    prep_x_f16_kernel<float>(x, T, d, ld_x, d_pad, reinterpret_cast<__half*>(out), row_scale, xnorm, xdnorm);
  });
}

void emu_candidate_bounds(const float* cand_vals, const long long* cand_idx, long long T, int K2, int k,
                          const float* wnorm, const float* dnorm, const float* xnorm, const float* xdnorm, float c_eps,
                          long long clamp_feature, float* lb_out, float* ub_out) {
  emu::launch({(unsigned)T}, {128}, [&] {
    candidate_bounds_kernel(cand_vals, cand_idx, K2, k, wnorm, dnorm, xnorm, xdnorm, c_eps, clamp_feature, lb_out, ub_out);
  });
}

// fused candidate selection + bound lists of the feature-sharded scan (encode_select_bounds_launch's geometry)
void emu_select_bounds(const void* cand, const int* cand_cnt, int T, int S, int CAP, int K2, int m1, const float* wnorm,
                       const float* dnorm, const float* xnorm, const float* xdnorm, float c_eps,
                       long long clamp_feature, float* out_vals, long long* out_idx, float* exch) {
  const int wpb = SSB_THREADS / 32;
  emu::launch({(unsigned)((T + wpb - 1) / wpb)}, {(unsigned)SSB_THREADS}, [&] {
    scan_select_bounds_kernel(reinterpret_cast<const uint2*>(cand), cand_cnt, T, S, CAP, K2, m1, wnorm, dnorm, xnorm,
                              xdnorm, c_eps, clamp_feature, out_vals, out_idx, exch);
  });
}

// saeb_gathered_bounds for R * m1 <= 256
void emu_gathered_bounds(const float* g, int R, long long T, int m1, int k, float* ext_L, float* ext_U) {
  emu::launch({(unsigned)((T + 3) / 4)}, {128}, [&] { gathered_bounds_kernel<8>(g, R, T, m1, k, ext_L, ext_U); });
}

// the refinement kernel on bf16 activations: lo == nullptr -> exact fp32 re-evaluation (refine_kernel), else the
// residual-plane correction (refine_lo_kernel)
void emu_refine_bf16(const void* x, long long T, long long ld_x, const float* W, long long d, long long N,
                     const float* bias, const float* wnorm, const float* dnorm, const float* trailer,
                     const float* xnorm, const float* xdnorm, float c_eps, const float* cand_vals,
                     const long long* cand_idx, int K2, int k, long long clamp_feature, float clamp_value,
                     float* out_vals, long long* out_idx, int* status, int* flag_rows, const float* ext_lower,
                     const void* lo, long long ld_w, int threads, int value_mode, const float* ext_upper,
                     const float* feat_thr, float* out_member) {
  const __nv_bfloat16* xb = reinterpret_cast<const __nv_bfloat16*>(x);
  if (threads < 0) {
    // warp-per-token kernel of the feature-sharded scan (refine_launch_t's `scan_warp` route); fewer warps than tokens
    emu::launch({(unsigned)((T + 7) / 8)}, {(unsigned)RSW_THREADS}, [&] {
      refine_scan_warp_kernel<__nv_bfloat16>(xb, ld_x, W, d, N, bias, wnorm, dnorm, trailer, xnorm, xdnorm, c_eps,
                                             cand_vals, cand_idx, K2, k, clamp_feature, clamp_value, out_vals, out_idx,
                                             status, flag_rows, ext_lower, T, nullptr, ext_upper, feat_thr, out_member,
                                             (d & 3) == 0);
    });
    return;
  }
  // fewer CTAs than tokens: the persistent token loop of the kernels is what runs beside the GEMM on the GPU
  emu::launch({(unsigned)((T + 1) / 2)}, {(unsigned)threads}, [&] {
    if (lo == nullptr)
      refine_kernel<__nv_bfloat16>(xb, ld_x, W, d, N, bias, wnorm, dnorm, trailer, xnorm, xdnorm, c_eps, cand_vals,
                                   cand_idx, K2, k, clamp_feature, clamp_value, out_vals, out_idx, status, flag_rows,
                                   ext_lower, T, nullptr, value_mode, ext_upper, feat_thr, out_member);
    else
      refine_lo_kernel<__nv_bfloat16>(xb, ld_x, reinterpret_cast<const __half*>(lo), ld_w, d, N, bias, wnorm, dnorm,
                                      trailer, xnorm, xdnorm, c_eps, cand_vals, cand_idx, K2, k, clamp_feature,
                                      clamp_value, out_vals, out_idx, status, flag_rows, ext_lower, T, nullptr, value_mode, ext_upper, feat_thr, out_member);
  });
}

void emu_compact_row(int slots, void* buf, int cnt_in, int k, float* thr_out, int* cnt_out) {
  uint2* b = reinterpret_cast<uint2*>(buf);
  emu::launch({1}, {32}, [&] {
    if (slots == 8) compact_row_wrapper<8>(b, cnt_in, k, thr_out, cnt_out);
    else if (slots == 16) compact_row_wrapper<16>(b, cnt_in, k, thr_out, cnt_out);
    else compact_row_wrapper<32>(b, cnt_in, k, thr_out, cnt_out);
  });
}

// exactly the launch geometry of encode_merge_launch for one chunk
void emu_topk_merge(const void* cand, const int* cand_cnt, int T, int S, int CAP, int k, int N, float* out_vals,
                    long long* out_idx) {
  int kp2 = 2;
  while (kp2 < k) kp2 <<= 1;
  const int max_entries = S * CAP;
  const size_t per_warp = (size_t)(max_entries + kp2) * sizeof(uint2);
  int wpb = (int)((200 * 1024) / per_warp);
  if (wpb > 8) wpb = 8;
  const unsigned blocks = (unsigned)((T + wpb - 1) / wpb);
  emu::launch({blocks}, {(unsigned)wpb * 32}, [&] {
    topk_merge_kernel(reinterpret_cast<const uint2*>(cand), cand_cnt, T, S, CAP, k, kp2, N, max_entries, out_vals,
                      out_idx);
  });
}

// the five launches of coo_extract_launch
void emu_coo_extract(const float* vals, const long long* idx, long long T, int k, float threshold,
                     const uint32_t* filter, long long seq_len, long long row_offset, long long* locations,
                     float* activations, long long* nnz_out) {
  std::vector<int> counts(T);
  std::vector<long long> offsets(T + 1);
  const long long nchunks = (T + SCAN_CHUNK - 1) / SCAN_CHUNK;
  std::vector<long long> sums(nchunks + 1);
  const unsigned g8 = (unsigned)((T + 7) / 8);
  emu::launch({g8}, {256}, [&] { coo_count_kernel(vals, idx, T, k, threshold, filter, counts.data()); });
  emu::launch({(unsigned)nchunks}, {(unsigned)SCAN_CHUNK}, [&] { scan_chunk_sum_kernel(counts.data(), T, sums.data()); });
  emu::launch({1}, {(unsigned)SCAN_CHUNK}, [&] { scan_sums_kernel(sums.data(), nchunks, nnz_out); });
  emu::launch({(unsigned)nchunks}, {(unsigned)SCAN_CHUNK},
              [&] { scan_chunks_kernel(counts.data(), T, sums.data(), offsets.data()); });
  int kp2 = 2;
  while (kp2 < k) kp2 <<= 1;
  emu::launch({g8}, {256}, [&] {
    coo_emit_kernel(vals, idx, T, k, kp2, threshold, filter, offsets.data(), seq_len, row_offset, locations, activations, nullptr, 0, nullptr);
  });
}

void emu_scan_pool(const float* vals, const long long* idx, long long T, int k, int ctx_len, float threshold,
                   long long feat_lo, long long feat_hi, long long window_base, const float* tok_thr,
                   const float* feat_thr, void* bucket, int* bucket_cnt, int bucket_cap, int* overflow,
                   const float* member) {
  const long long n_win = (T + ctx_len - 1) / ctx_len;
  int slots = 64;
  while (slots < 2 * (long long)ctx_len * k) slots <<= 1;
  emu::launch({(unsigned)n_win}, {256}, [&] {
    scan_pool_kernel(vals, idx, T, k, ctx_len, threshold, feat_lo, feat_hi, window_base, tok_thr, member, feat_thr,
                     reinterpret_cast<uint2*>(bucket), bucket_cnt, bucket_cap, slots, overflow);
  });
}

// image scan: `grid` persistent CTAs over the images, scratch laid out by the caller (keys = 0xffffffff, sums = 0)
void emu_image_pool(const float* vals, const long long* idx, long long n_images, long long tokens_per_image, int k,
                    int n_base, float threshold, long long feat_lo, long long feat_hi, long long image_base,
                    const float* tok_thr, const float* feat_thr, uint32_t* hkeys, unsigned long long* hsums, int* hlist,
                    int slots, void* bucket, int* bucket_cnt, int bucket_cap, int* overflow, int grid) {
  emu::launch({(unsigned)grid}, {256}, [&] {
    image_pool_kernel(vals, idx, n_images, tokens_per_image, k, n_base, threshold, feat_lo, feat_hi, image_base,
                      tok_thr, feat_thr, hkeys, hsums, hlist, slots, reinterpret_cast<uint2*>(bucket), bucket_cnt,
                      bucket_cap, overflow);
  });
}

void emu_scan_merge(void* bucket, int* bucket_cnt, int bucket_cap, long long F, int n_top, float base_thr,
                    float* top_vals, long long* top_win, float* feat_thr) {
  int sort_n = 2;
  while (sort_n < n_top + bucket_cap) sort_n <<= 1;
  emu::launch({(unsigned)((F + 7) / 8)}, {256}, [&] {
    scan_merge_kernel(reinterpret_cast<uint2*>(bucket), bucket_cnt, bucket_cap, F, n_top, sort_n, base_thr, top_vals,
                      top_win, feat_thr);
  });
}

// decode: W as fp32 (w16 == 0) or as an fp16 copy; x (bf16, optional) + sq_err accumulate; `scalar` forces the
// one-column-per-thread kernel
void emu_decode(const long long* idx, const float* vals, long long T, int k, const void* W, int w16, long long d,
                long long N, const float* b_dec, float* out, const void* x, double* sq_err, int* err_flag, int scalar) {
  const __nv_bfloat16* xb = reinterpret_cast<const __nv_bfloat16*>(x);
  emu::launch({(unsigned)((T + 2) / 3)}, {(unsigned)DEC_THREADS}, [&] {
    if (scalar)
      decode_scalar_kernel<float, float, __nv_bfloat16>(idx, vals, k, reinterpret_cast<const float*>(W), d, N, b_dec,
                                                         out, d, xb, d, sq_err, err_flag, T);
    else if (w16)
      decode_kernel<__half, float, __nv_bfloat16>(idx, vals, k, reinterpret_cast<const __half*>(W), d, N, b_dec, out, d,
                                                  xb, d, sq_err, err_flag, T);
    else
      decode_kernel<float, float, __nv_bfloat16>(idx, vals, k, reinterpret_cast<const float*>(W), d, N, b_dec, out, d,
                                                 xb, d, sq_err, err_flag, T);
  });
}

void emu_dense_topk(const float* dense, long long T, long long ld, long long N, int k, float* out_vals,
                    long long* out_idx, int threads) {
  emu::launch({(unsigned)T}, {(unsigned)threads},
              [&] { dense_topk_kernel(dense, ld, N, k, nullptr, 0, nullptr, out_vals, out_idx, nullptr); });
}

// the three fallback launches of refine_launch_t for flagged rows (bf16 activations): exact dense rows of the first
// RF_MAX_FLAG flagged tokens + their dense TopK, then the overflow kernel for the rest
void emu_refine_fallback(const void* x, long long ld_x, const float* W, long long d, long long N, const float* bias,
                         const int* status, const int* flag_rows, long long clamp_feature, float clamp_value,
                         float* dense_scratch, int k, float* out_vals, long long* out_idx, int gx, int threads,
                         int stage_x) {
  const __nv_bfloat16* xb = reinterpret_cast<const __nv_bfloat16*>(x);
  emu::launch({(unsigned)gx, (unsigned)RF_MAX_FLAG}, {256}, [&] {
    exact_rows_kernel<__nv_bfloat16>(xb, ld_x, W, d, N, bias, status, flag_rows, clamp_feature, clamp_value,
                                     dense_scratch, stage_x);
  });
  emu::launch({(unsigned)RF_MAX_FLAG}, {(unsigned)threads}, [&] {
    dense_topk_kernel(dense_scratch, N, N, k, status, RF_MAX_FLAG, flag_rows, out_vals, out_idx, nullptr);
  });
  emu::launch({3}, {(unsigned)threads}, [&] {
    overflow_rows_kernel<__nv_bfloat16>(xb, ld_x, W, d, N, bias, status, flag_rows, clamp_feature, clamp_value,
                                        dense_scratch, k, out_vals, out_idx, nullptr, stage_x);
  });
}

// FVU denominator: colstats_kernel + totvar_kernel as total_variance_launch enqueues them (bf16 activations)
void emu_total_variance_bf16(const void* x, long long T, long long d, long long ld_x, double* scratch, double* out) {
  std::memset(scratch, 0, sizeof(double) * 2 * d);
  const int rpb = 64;
  emu::launch({(unsigned)((T + rpb - 1) / rpb)}, {256}, [&] {
    colstats_kernel<__nv_bfloat16>(reinterpret_cast<const __nv_bfloat16*>(x), T, d, ld_x, rpb, scratch, scratch + d);
  });
  emu::launch({1}, {256}, [&] { totvar_kernel(scratch, scratch + d, T, d, out); });
}

// one rank of the peer-memory all-gather, launched like push_gather_launch does (`blocks` CTAs); `peer_bases` holds
// this process' mappings of all R symmetric buffers
void emu_push_gather(const void* src, size_t bytes, void* const* peer_bases, int R, int self, size_t region_offset,
                     size_t flags_offset, int channel, unsigned seq, int* counter, int blocks) {
  const size_t n_vec = bytes / 16;
  const size_t slab_off = region_offset + (size_t)self * bytes;
  emu::launch({(unsigned)blocks}, {(unsigned)PUSH_THREADS}, [&] {
    push_gather_kernel(reinterpret_cast<const uint4*>(src), n_vec, peer_bases, R, self, slab_off, nullptr, flags_offset,
                       channel, seq, counter);
  });
}

// refinement on fp16 activations (exact or residual-plane) and on fp32 activations (exact only: refine_launch_t never
// takes the residual route for them)
void emu_refine_f16(const void* x, long long T, long long ld_x, const float* W, long long d, long long N,
                    const float* bias, const float* wnorm, const float* dnorm, const float* trailer,
                    const float* xnorm, const float* xdnorm, float c_eps, const float* cand_vals,
                    const long long* cand_idx, int K2, int k, float* out_vals, long long* out_idx, int* status,
                    int* flag_rows, const void* lo, long long ld_w) {
  const __half* xh = reinterpret_cast<const __half*>(x);
  emu::launch({(unsigned)T}, {256}, [&] {
    if (lo == nullptr)
      refine_kernel<__half>(xh, ld_x, W, d, N, bias, wnorm, dnorm, trailer, xnorm, xdnorm, c_eps, cand_vals, cand_idx,
                            K2, k, -1, 0.f, out_vals, out_idx, status, flag_rows, nullptr, T, nullptr, 0, nullptr, nullptr, nullptr);
    else
      refine_lo_kernel<__half>(xh, ld_x, reinterpret_cast<const __half*>(lo), ld_w, d, N, bias, wnorm, dnorm, trailer,
                               xnorm, xdnorm, c_eps, cand_vals, cand_idx, K2, k, -1, 0.f, out_vals, out_idx, status,
                               flag_rows, nullptr, T, nullptr, 0, nullptr, nullptr, nullptr);
  });
}
void emu_refine_f32(const float* x, long long T, long long ld_x, const float* W, long long d, long long N,
                    const float* bias, const float* wnorm, const float* dnorm, const float* trailer,
                    const float* xnorm, const float* xdnorm, float c_eps, const float* cand_vals,
                    const long long* cand_idx, int K2, int k, float* out_vals, long long* out_idx, int* status,
                    int* flag_rows) {
  emu::launch({(unsigned)T}, {256}, [&] {
    refine_kernel<float>(x, ld_x, W, d, N, bias, wnorm, dnorm, trailer, xnorm, xdnorm, c_eps, cand_vals, cand_idx, K2, k,
                         -1, 0.f, out_vals, out_idx, status, flag_rows, nullptr, T, nullptr, 0, nullptr, nullptr, nullptr);
  });
}

}  // extern "C"
