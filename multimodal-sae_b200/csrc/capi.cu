// extern "C" surface declared in include/saeb200.h: argument checking, workspace carving, launch accounting.
#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstring>

#include <nvtx3/nvToolsExt.h>   // header-only NVTX v3: ranges cost nothing unless a tool (nsys / ncu --nvtx) is attached

#include "../../include/saeb200.h"
#include "common.cuh"

namespace saeb {

static thread_local char g_err[512] = {0};
static thread_local int g_default_margin = 0;   // extra candidates per row in refine mode; 0: max(48, k/2)
static std::atomic<long long> g_launches{0};

// NVTX range around every compute entry point ("saeb:<phase>"): prep / gemm / bounds / refine / decode / coo / scan /
// exchange show up as named ranges on the calling thread's timeline (tracing)
struct NvtxRange {
  explicit NvtxRange(const char* name) { nvtxRangePushA(name); }
  ~NvtxRange() { nvtxRangePop(); }
};
#define SAEB_NVTX(name) saeb::NvtxRange _saeb_nvtx_range(name)

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

// implemented in the kernel translation units
size_t encode_workspace_bytes(long long T, long long d, long long N, int k);
int encode_topk_launch(const void* x_planes, int ap, long long T, long long ld_x, long long x_plane_stride,
                       const void* w_planes, int bp, long long ld_w, const float* bias, long long d, long long N, int k,
                       long long clamp_feature, float clamp_value, float* out_vals, long long* out_idx,
                       float* dense_out, long long ld_dense, void* workspace, size_t workspace_bytes, int pass_mask,
                       int operand_fmt, const float* row_scale, const float* w_unscale, cudaStream_t stream);
int encode_gemm_launch(const void* x_planes, int ap, long long T, long long ld_x, long long x_plane_stride,
                       const void* w_planes, int bp, long long ld_w, const float* bias, long long d, long long N, int k,
                       long long clamp_feature, float clamp_value, bool do_topk, float* dense_out, long long ld_dense,
                       void* workspace, size_t workspace_bytes, int pass_mask, int operand_fmt, const float* row_scale,
                       const float* w_unscale, cudaStream_t stream);
int encode_merge_launch(long long T, long long N, int k, float* out_vals, long long* out_idx, void* workspace,
                        size_t workspace_bytes, int coresident, cudaStream_t stream);
int set_chunking(int v);
int set_reserve_sms(int v);
int set_gemm_stages(int v);
int set_cluster4(int v);
long long query_max_clusters4();
int pack_weights_f16_launch(const float* W_enc, const float* b_enc, const float* b_dec, long long N, long long d,
                            long long d_pad, void* w_plane, float* bias, float* wnorm, float* dnorm, float* trailer,
                            cudaStream_t stream);
int prep_x_f16_launch(const void* x, int x_dtype, long long T, long long d, long long ld_x, long long d_pad, void* out,
                      float* row_scale, float* xnorm, float* xdnorm, cudaStream_t stream);
size_t refine_fallback_bytes(long long N);
int refine_launch(const void* x, int x_dtype, long long T, long long ld_x, const float* W, long long d, long long N,
                  const float* bias, const float* wnorm, const float* dnorm, const float* trailer, const float* xnorm,
                  const float* xdnorm, float c_eps,
                  const float* cand_vals, const long long* cand_idx, int K2, int k, long long clamp_feature,
                  float clamp_value, float* out_vals, long long* out_idx, int* status, int* flag_rows,
                  float* dense_scratch, const float* ext_lower, const void* w_lo, long long ld_w, int max_ctas,
                  int value_mode, const float* ext_upper, const float* feat_thr, float* out_member,
                  cudaStream_t stream);
int pack_weights_lo_launch(const float* W_enc, long long N, long long d, long long d_pad, const float* trailer,
                           void* lo_plane, cudaStream_t stream);
int candidate_bounds_launch(const float* cand_vals, const long long* cand_idx, long long T, int K2, int k,
                            const float* wnorm, const float* dnorm, const float* xnorm, const float* xdnorm,
                            float c_eps, long long clamp_feature, float* lb_out, float* ub_out, cudaStream_t stream);
int dense_topk_launch(const float* dense, long long T, long long ld, long long N, int k, float* out_vals,
                      long long* out_idx, cudaStream_t stream);
int set_splits(int v);
int set_l2_hints(int v);
int set_dbg(int v);
int set_prefetch_b(int v);
int set_stats(int v);
int read_stats(unsigned long long* out8);
int set_persist_a(int v);
long long persist_bytes();
int pack_weights_launch(const float* W_enc, const float* b_enc, const float* b_dec, long long N, long long d,
                        long long d_pad, int planes, void* w_planes, float* bias, cudaStream_t stream);
int split_x_launch(const void* x, int x_dtype, long long T, long long d, long long ld_x, long long d_pad, int planes,
                   void* out, cudaStream_t stream);
int decode_launch(const long long* idx, const float* vals, long long T, int k, const void* W_dec, int w_dtype,
                  long long d, long long N, const float* b_dec, void* out, int out_dtype, long long ld_out,
                  const void* x, int x_dtype, long long ld_x, double* sq_err, int* err_flag, int max_ctas,
                  cudaStream_t stream);
int total_variance_launch(const void* x, int x_dtype, long long T, long long d, long long ld_x, double* scratch,
                          double* out, cudaStream_t stream);
size_t coo_workspace_bytes(long long T);
int coo_extract_launch(const float* vals, const long long* idx, long long T, int k, float threshold,
                       const uint32_t* filter, long long seq_len, long long row_offset, long long* locations,
                       float* activations, long long* nnz_out, void* ws, size_t ws_bytes, long long* cursor,
                       long long capacity, int* overflow, cudaStream_t stream);
int kth_gathered_launch(const float* gathered, int R, long long T, int m, int kth, float* tok_thr,
                        cudaStream_t stream);
int set_kth_impl(int v);
int gathered_bounds_launch(const float* gathered, int R, long long T, int m1, int k, float* ext_L, float* ext_U,
                           cudaStream_t stream);
size_t scan_pool_workspace_bytes(int k, int ctx_len);
int scan_pool_init(void* ws, size_t ws_bytes, int k, int ctx_len, cudaStream_t stream);
int scan_pool_ws_launch(const float* vals, const long long* idx, long long T, int k, int ctx_len, float threshold,
                        long long feat_lo, long long feat_hi, long long window_base, const float* tok_thr,
                        const float* member, const float* feat_thr, void* bucket, int* bucket_cnt, int bucket_cap,
                        int* overflow, void* ws, size_t ws_bytes, cudaStream_t stream);
size_t image_pool_workspace_bytes(int k, int n_base, long long F);
int image_pool_init(void* ws, size_t ws_bytes, int k, int n_base, long long F, cudaStream_t stream);
int image_pool_launch(const float* vals, const long long* idx, long long n_images, long long tokens_per_image, int k,
                      int n_base, float threshold, long long feat_lo, long long feat_hi, long long image_base,
                      const float* tok_thr, const float* feat_thr, void* bucket, int* bucket_cnt, int bucket_cap,
                      int* overflow, void* ws, size_t ws_bytes, cudaStream_t stream);
int set_refine_threads(int v);
int encode_select_bounds_launch(long long T, long long N, int K2, int m1, const float* wnorm, const float* dnorm,
                                const float* xnorm, const float* xdnorm, float c_eps, long long clamp_feature,
                                float* out_vals, long long* out_idx, float* exch, void* workspace,
                                size_t workspace_bytes, cudaStream_t stream);
int set_scan_warp(int v);
int coload_launch(int mode, int ctas, long long iters, const void* buf, size_t bytes, float* sink, cudaStream_t stream);
int scan_warp_enabled();
int decode_bwd_acts_launch(const float* grad_out, long long ld_g, const long long* idx, long long T, int k,
                           const float* W_dec, long long d, long long N, float* d_vals, int* err_flag,
                           cudaStream_t stream);
int decode_bwd_weight_launch(const float* grad_out, long long ld_g, const long long* idx, const float* vals,
                             long long T, int k, long long d, long long N, float* dW, int* err_flag,
                             cudaStream_t stream);
int push_gather_launch(const void* src, size_t bytes, void* const* peer_bases_dev, int R, int self_rank,
                       size_t region_offset, void* multicast_base, size_t flags_offset, int channel, unsigned int seq,
                       int* counter, cudaStream_t stream);
int scan_pool_launch(const float* vals, const long long* idx, long long T, int k, int ctx_len, float threshold,
                     long long feat_lo, long long feat_hi, long long window_base, const float* tok_thr,
                     const float* member, const float* feat_thr, void* bucket, int* bucket_cnt, int bucket_cap,
                     int* overflow, cudaStream_t stream);
int scan_merge_launch(void* bucket, int* bucket_cnt, int bucket_cap, long long F, int n_top, float base_thr,
                      float* top_vals, long long* top_win, float* feat_thr, cudaStream_t stream);

int column_sums_launch(const float* dense, long long T, long long ld, long long N, double* colsum, cudaStream_t stream);
int feature_maps_launch(const void* x, int x_dtype, long long T, long long ld_x, const float* W, const float* b_enc,
                        const float* b_dec, long long d, long long N, const long long* sel, int n_sel, float* out,
                        int* err_flag, cudaStream_t stream);

int coo_window_scores_launch(const long long* feat, const long long* key, const float* act, long long nnz, int mode,
                             float divisor, float* score, int* head, cudaStream_t stream);
int set_cta_pair(int v);
int set_profile(int v);
float last_encode_ms();

static inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }
static inline long long pad8(long long d) { return (d + 7) / 8 * 8; }   // 16-byte row strides for TMA
// planes: 1 / 2 = bf16 planes; 3 = "fp16 + refine" mode (one scaled fp16 plane + per-feature norms + trailer);
// 4 = mode 3 followed by the fp16 residual plane (every mode-3 offset is unchanged: a mode-4 blob IS a mode-3 blob)
static inline int n_planes(int planes) { return planes >= 3 ? 1 : planes; }
static inline size_t mode3_bytes(long long N, long long d);
static inline size_t planes_bytes(long long N, long long d, int planes) {
  return align_up((size_t)n_planes(planes) * (size_t)N * (size_t)pad8(d) * 2, 256);
}
static inline size_t bias_bytes(long long N) { return align_up((size_t)N * sizeof(float), 256); }
// workspace for bf16 activation planes: none for bf16 input consumed in place (d % 8 == 0), one padded plane for
// bf16 with d % 8 != 0, two planes for fp16 / fp32 input
static inline size_t x_split_bytes(long long T, long long d, int x_dtype) {
  if (x_dtype == DT_BF16 && d % 8 == 0) return 0;
  return align_up((size_t)(x_dtype == DT_BF16 ? 1 : 2) * (size_t)T * (size_t)pad8(d) * 2, 1024);
}

static inline size_t mode3_bytes(long long N, long long d) {
  return align_up(planes_bytes(N, d, 3) + 3 * bias_bytes(N) + 256, 1024);
}

}  // namespace saeb

using namespace saeb;

extern "C" {

int saeb_version(void) { return 100; }
const char* saeb_last_error(void) { return g_err; }
long long saeb_launch_count(void) { return g_launches.load(); }
int saeb_set_option(const char* name, int value) {
  g_err[0] = 0;
  SAEB_REQUIRE(name != nullptr, "set_option: null name");
  if (strcmp(name, "cta_pair") == 0) return set_cta_pair(value);
  if (strcmp(name, "profile") == 0) return set_profile(value);
  if (strcmp(name, "splits") == 0) return set_splits(value);
  if (strcmp(name, "l2_hints") == 0) return set_l2_hints(value);
  if (strcmp(name, "debug_tiles") == 0) return set_dbg(value);
  if (strcmp(name, "prefetch_b") == 0) return set_prefetch_b(value);
  if (strcmp(name, "stats") == 0) return set_stats(value);
  if (strcmp(name, "persist_a") == 0) return set_persist_a(value);
  if (strcmp(name, "chunking") == 0) return set_chunking(value);
  if (strcmp(name, "reserve_sms") == 0) return set_reserve_sms(value);
  if (strcmp(name, "gemm_stages") == 0) return set_gemm_stages(value);
  if (strcmp(name, "cluster4") == 0) return set_cluster4(value);
  if (strcmp(name, "kth_impl") == 0) return set_kth_impl(value);
  if (strcmp(name, "refine_threads") == 0) return set_refine_threads(value);
  if (strcmp(name, "scan_warp") == 0) return set_scan_warp(value);
  if (strcmp(name, "refine_margin") == 0) {
    g_default_margin = value;
    return 0;
  }
  set_error("set_option: unknown option '%s'", name);
  return -1;
}

float saeb_profile_last_encode_ms(void) { return last_encode_ms(); }
int saeb_debug_stats(unsigned long long* out8) { return read_stats(out8); }
long long saeb_query(const char* name) {
  if (name == nullptr) return -1;
  int dev = 0, v = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return -2;
  if (strcmp(name, "persisting_l2_max_bytes") == 0) {
    cudaDeviceGetAttribute(&v, cudaDevAttrMaxPersistingL2CacheSize, dev);
    return v;
  }
  if (strcmp(name, "access_policy_max_window_bytes") == 0) {
    cudaDeviceGetAttribute(&v, cudaDevAttrMaxAccessPolicyWindowSize, dev);
    return v;
  }
  if (strcmp(name, "l2_bytes") == 0) {
    cudaDeviceGetAttribute(&v, cudaDevAttrL2CacheSize, dev);
    return v;
  }
  if (strcmp(name, "num_sms") == 0) {
    cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev);
    return v;
  }
  if (strcmp(name, "persisting_l2_in_use_bytes") == 0) return persist_bytes();
  if (strcmp(name, "max_clusters4") == 0) return query_max_clusters4();
  return -1;
}

size_t saeb_packed_bias_offset(int64_t N, int64_t d, int planes) { return planes_bytes(N, d, planes); }
size_t saeb_packed_weights_bytes(int64_t N, int64_t d, int planes) {
  // bias [N]; mode 3 appends ||w_j|| [N], ||w_j - fp16(w_j)|| [N] and a 256-byte trailer
  // {w_unscale, max ||w_j||, scratch, max rounding-error norm}; mode 4 appends the residual plane after that
  if (planes == 4) return mode3_bytes(N, d) + planes_bytes(N, d, 3);
  return planes_bytes(N, d, planes) + bias_bytes(N) + (planes == 3 ? 2 * bias_bytes(N) + 256 : 0);
}

int saeb_pack_weights(const float* W_enc, const float* b_enc, const float* b_dec, int64_t N, int64_t d, int planes,
                      void* packed, void* stream) {
  g_err[0] = 0;
  SAEB_NVTX("saeb:pack_weights");
  SAEB_REQUIRE(W_enc && b_enc && b_dec && packed, "pack_weights: null pointer");
  SAEB_REQUIRE(planes >= 1 && planes <= 4, "pack_weights: planes must be 1, 2, 3 or 4");
  float* bias = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(packed) + planes_bytes(N, d, planes));
  if (planes >= 3) {
    float* wnorm = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(bias) + bias_bytes(N));
    float* dnorm = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(wnorm) + bias_bytes(N));
    float* trailer = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(dnorm) + bias_bytes(N));
    int rc3 = pack_weights_f16_launch(W_enc, b_enc, b_dec, N, d, pad8(d), packed, bias, wnorm, dnorm, trailer,
                                      (cudaStream_t)stream);
    if (rc3 == 0) g_launches += 2;
    if (rc3 == 0 && planes == 4) {
      rc3 = pack_weights_lo_launch(W_enc, N, d, pad8(d), trailer, reinterpret_cast<uint8_t*>(packed) + mode3_bytes(N, d),
                                   (cudaStream_t)stream);
      if (rc3 == 0) g_launches += 1;
    }
    return rc3;
  }
  int rc = pack_weights_launch(W_enc, b_enc, b_dec, N, d, pad8(d), planes, packed, bias, (cudaStream_t)stream);
  if (rc == 0) g_launches += 2;
  return rc;
}

size_t saeb_encode_topk_workspace_bytes(int64_t T, int64_t d, int64_t N, int k, int x_dtype) {
  return x_split_bytes(T, d, x_dtype) + encode_workspace_bytes(T, d, N, k);
}

int saeb_encode_topk(const void* x, int x_dtype, int64_t T, int64_t ld_x, const void* packed, int planes, int64_t d,
                     int64_t N, int k, int64_t clamp_feature, float clamp_value, float* out_vals, int64_t* out_idx,
                     float* dense_out, int64_t ld_dense, void* workspace, size_t workspace_bytes, void* stream) {
  g_err[0] = 0;
  SAEB_NVTX("saeb:encode_topk");
  SAEB_REQUIRE(x && packed, "encode_topk: null pointer");
  SAEB_REQUIRE((out_vals != nullptr) == (out_idx != nullptr), "encode_topk: out_vals and out_idx go together");
  SAEB_REQUIRE(out_vals != nullptr || dense_out != nullptr, "encode_topk: nothing to compute");
  SAEB_REQUIRE(planes == 1 || planes == 2, "encode_topk: planes must be 1 or 2 (mode 3 goes through saeb_encode_topk_refine)");
  SAEB_REQUIRE(x_dtype == DT_BF16 || x_dtype == DT_F16 || x_dtype == DT_F32, "encode_topk: bad x dtype %d", x_dtype);
  SAEB_REQUIRE(clamp_feature < N, "encode_topk: clamp_feature out of range");
  if (T == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  const float* bias =
      reinterpret_cast<const float*>(reinterpret_cast<const uint8_t*>(packed) + planes_bytes(N, d, planes));
  const void* xp = x;
  int ap = 1;
  long long ldx = ld_x, xps = (long long)T * ld_x;
  uint8_t* ws = reinterpret_cast<uint8_t*>(workspace);
  size_t ws_left = workspace_bytes;
  const bool in_place = x_dtype == DT_BF16 && ld_x % 8 == 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0;
  if (!in_place) {
    SAEB_REQUIRE(!(x_dtype == DT_BF16 && d % 8 == 0),
                 "encode_topk: bf16 activations need a 16-byte aligned base and ld_x %% 8 == 0 (got ld_x=%lld)",
                 (long long)ld_x);
    const size_t need = x_split_bytes(T, d, x_dtype);
    SAEB_REQUIRE(workspace != nullptr && workspace_bytes >= need, "encode_topk: workspace too small for x planes");
    ap = (x_dtype == DT_BF16) ? 1 : 2;
    int rc = split_x_launch(x, x_dtype, T, d, ld_x, pad8(d), ap, ws, st);
    if (rc) return rc;
    g_launches += 1;
    xp = ws;
    ldx = pad8(d);
    xps = (long long)T * pad8(d);
    ws += need;
    ws_left -= need;
  }
  // plane products issued: all (a, b) pairs except lo*lo, which is below the residual of the two-plane split
  int pass_mask = 0;
  for (int a = 0; a < ap; ++a)
    for (int b = 0; b < planes; ++b)
      if (!(a == 1 && b == 1)) pass_mask |= 1 << (a * planes + b);
  int rc = encode_topk_launch(xp, ap, T, ldx, xps, packed, planes, pad8(d), bias, d, N, k, clamp_feature, clamp_value, out_vals,
                              reinterpret_cast<long long*>(out_idx), dense_out, ld_dense, ws, ws_left, pass_mask,
                              /*operand_fmt=bf16*/ 1, nullptr, nullptr, st);
  if (rc == 0) g_launches += out_vals ? 2 : 1;
  return rc;
}

// ---- "fp16 + refine": one tensor-core pass + exact fp32 re-evaluation of the candidates near the k-th value
static inline int refine_k2(int k, int margin) {
  const int auto_m = g_default_margin > 0 ? g_default_margin : (k / 2 > 48 ? k / 2 : 48);
  int m = margin > 0 ? margin : auto_m;   // denser spectra (large k / N) need more room below the k-th value
  int K2 = k + m;
  if (K2 > 512) K2 = 512;
  return K2;
}
// prepared activations of a whole batch: x16 [T][d_pad] fp16 | row_scale [T] | xnorm [T] | xdnorm [T]  (f32)
struct PrepLayout {
  size_t x16, row_scale, xnorm, xdnorm, total;
};
static PrepLayout prep_layout(long long T, long long d) {
  PrepLayout p;
  p.x16 = 0;
  p.row_scale = align_up((size_t)T * pad8(d) * 2, 1024);
  p.xnorm = p.row_scale + align_up((size_t)T * 4, 256);
  p.xdnorm = p.xnorm + align_up((size_t)T * 4, 256);
  p.total = p.xdnorm + align_up((size_t)T * 4, 256);
  return p;
}
// per-call (or per-chunk) scratch of the candidate pipeline
struct RefineWs {
  size_t status, flag_rows, mvals, midx, dense, enc, total;
};
static RefineWs refine_ws(long long T, long long d, long long N, int k, int margin) {
  RefineWs w;
  const int K2 = refine_k2(k, margin);
  size_t off = 0;
  auto take = [&](size_t bytes, size_t al) {
    off = align_up(off, al);
    size_t at = off;
    off += bytes;
    return at;
  };
  w.status = take(256, 256);
  w.flag_rows = take((size_t)(T > 64 ? T : 64) * 4, 256);   // every row of the call can be flagged
  w.mvals = take((size_t)T * K2 * 4, 256);
  w.midx = take((size_t)T * K2 * 8, 256);
  w.dense = take(refine_fallback_bytes(N), 256);
  w.enc = take(encode_workspace_bytes(T, d, N, K2), 1024);
  w.total = align_up(off, 1024);
  return w;
}

size_t saeb_prep_bytes(int64_t T, int64_t d) { return prep_layout(T, d).total; }

int saeb_prep_activations(const void* x, int x_dtype, int64_t T, int64_t ld_x, int64_t d, void* prep, void* stream) {
  g_err[0] = 0;
  SAEB_NVTX("saeb:prep");
  SAEB_REQUIRE(x && prep, "prep_activations: null pointer");
  SAEB_REQUIRE(x_dtype == DT_BF16 || x_dtype == DT_F16 || x_dtype == DT_F32, "prep_activations: bad x dtype %d", x_dtype);
  if (T == 0) return 0;
  const PrepLayout p = prep_layout(T, d);
  uint8_t* b = reinterpret_cast<uint8_t*>(prep);
  int rc = prep_x_f16_launch(x, x_dtype, T, d, ld_x, pad8(d), b + p.x16, reinterpret_cast<float*>(b + p.row_scale),
                             reinterpret_cast<float*>(b + p.xnorm), reinterpret_cast<float*>(b + p.xdnorm),
                             (cudaStream_t)stream);
  if (rc == 0) g_launches += 1;
  return rc;
}

size_t saeb_candidates_workspace_bytes(int64_t T, int64_t d, int64_t N, int k, int margin) {
  return refine_ws(T, d, N, k, margin).total;
}

// phase A: single-pass GEMM with fused candidate selection over rows [t0, t0+Tc) of a prepared batch (tensor bound)
int saeb_encode_candidates(const void* prep, int64_t T_total, int64_t t0, int64_t Tc, const void* packed, int64_t d,
                           int64_t N, int k, int margin, int64_t clamp_feature, float clamp_value, void* workspace,
                           size_t workspace_bytes, void* stream) {
  g_err[0] = 0;
  SAEB_NVTX("saeb:gemm");
  SAEB_REQUIRE(prep && packed && workspace, "encode_candidates: null pointer");
  SAEB_REQUIRE(t0 >= 0 && Tc >= 0 && t0 + Tc <= T_total, "encode_candidates: bad row range");
  SAEB_REQUIRE(clamp_feature < N, "encode_candidates: clamp_feature out of range");
  SAEB_REQUIRE(k >= 1 && k <= N && k <= 448, "encode_candidates: k=%d out of range", k);
  if (Tc == 0) return 0;
  const int K2raw = refine_k2(k, margin);
  const int K2 = K2raw < N ? K2raw : (int)N;
  const RefineWs w = refine_ws(Tc, d, N, k, margin);
  SAEB_REQUIRE(workspace_bytes >= w.total, "encode_candidates: workspace too small: have %zu need %zu",
               workspace_bytes, w.total);
  const PrepLayout p = prep_layout(T_total, d);
  const uint8_t* pb = reinterpret_cast<const uint8_t*>(prep);
  uint8_t* ws = reinterpret_cast<uint8_t*>(workspace);
  const uint8_t* pk = reinterpret_cast<const uint8_t*>(packed);
  const float* bias = reinterpret_cast<const float*>(pk + planes_bytes(N, d, 3));
  const float* dnorm = reinterpret_cast<const float*>(pk + planes_bytes(N, d, 3) + 2 * bias_bytes(N));
  const float* trailer = reinterpret_cast<const float*>(pk + planes_bytes(N, d, 3) + 3 * bias_bytes(N));
  int rc = encode_gemm_launch(pb + p.x16 + (size_t)t0 * pad8(d) * 2, 1, Tc, pad8(d), (long long)Tc * pad8(d), packed, 1,
                              pad8(d), bias, d, N, K2, clamp_feature, clamp_value, true, nullptr, 0, ws + w.enc,
                              w.total - w.enc, 1, /*operand_fmt=fp16*/ 0,
                              reinterpret_cast<const float*>(pb + p.row_scale) + t0, trailer, (cudaStream_t)stream);
  if (rc == 0) g_launches += 1;
  return rc;
}

// phase B: candidate merge + exact fp32 re-evaluation (+ dense fallback) for the same rows (HBM bound).
// x points at row t0 of the ORIGINAL activations.
static inline float refine_c_eps(int /*x_dtype*/) {
  // slack for the fp32 accumulation inside the tensor cores (the operand rounding errors are bounded exactly through
  // the stored error norms): 256 block additions of <= 2^-23 relative error each, doubled
  return ldexpf(1.0f, -14);
}

// feature-sharded scan, step 1: merge this shard's candidates and emit, per token, the k largest lower bounds
// (a_j - eps_j, descending).  Leaves the merged candidates in the workspace for saeb_refine_candidates(...,
// already_merged = 1).
int saeb_candidate_bounds(const void* prep, int64_t T_total, int64_t t0, int64_t Tc, const void* packed, int x_dtype,
                          int64_t d, int64_t N, int k, int margin, int64_t clamp_feature, float* lb_out,
                          float* ub_out, void* workspace, size_t workspace_bytes, int coresident, void* stream) {
  g_err[0] = 0;
  SAEB_NVTX("saeb:merge+bounds");
  SAEB_REQUIRE(prep && packed && lb_out && workspace, "candidate_bounds: null pointer");
  SAEB_REQUIRE(t0 >= 0 && Tc >= 0 && t0 + Tc <= T_total, "candidate_bounds: bad row range");
  if (Tc == 0) return 0;
  const int K2raw = refine_k2(k, margin);
  const int K2 = K2raw < N ? K2raw : (int)N;
  const RefineWs w = refine_ws(Tc, d, N, k, margin);
  SAEB_REQUIRE(workspace_bytes >= w.total, "candidate_bounds: workspace too small");
  cudaStream_t st = (cudaStream_t)stream;
  const PrepLayout p = prep_layout(T_total, d);
  const uint8_t* pb = reinterpret_cast<const uint8_t*>(prep);
  uint8_t* ws = reinterpret_cast<uint8_t*>(workspace);
  const uint8_t* pk = reinterpret_cast<const uint8_t*>(packed);
  const float* wnorm = reinterpret_cast<const float*>(pk + planes_bytes(N, d, 3) + bias_bytes(N));
  const float* dnorm = reinterpret_cast<const float*>(pk + planes_bytes(N, d, 3) + 2 * bias_bytes(N));
  float* mvals = reinterpret_cast<float*>(ws + w.mvals);
  long long* midx = reinterpret_cast<long long*>(ws + w.midx);
  int rc = encode_merge_launch(Tc, N, K2, mvals, midx, ws + w.enc, w.total - w.enc, coresident, st);
  if (rc) return rc;
  rc = candidate_bounds_launch(mvals, midx, Tc, K2, k < K2 ? k : K2, wnorm, dnorm,
                               reinterpret_cast<const float*>(pb + p.xnorm) + t0,
                               reinterpret_cast<const float*>(pb + p.xdnorm) + t0, refine_c_eps(x_dtype), clamp_feature,
                               lb_out, ub_out, st);
  if (rc == 0) g_launches += 2;
  return rc;
}

// feature-sharded scan, step 1 as ONE kernel (see include/saeb200.h)
int saeb_candidate_bounds_packed(const void* prep, int64_t T_total, int64_t t0, int64_t Tc, const void* packed,
                                 int x_dtype, int64_t d, int64_t N, int k, int margin, int64_t clamp_feature, int m1,
                                 float* bounds_out, void* workspace, size_t workspace_bytes, void* stream) {
  g_err[0] = 0;
  SAEB_NVTX("saeb:select+bounds");
  SAEB_REQUIRE(prep && packed && bounds_out && workspace, "candidate_bounds_packed: null pointer");
  SAEB_REQUIRE(t0 >= 0 && Tc >= 0 && t0 + Tc <= T_total, "candidate_bounds_packed: bad row range");
  if (Tc == 0) return 0;
  const int K2raw = refine_k2(k, margin);
  const int K2 = K2raw < N ? K2raw : (int)N;
  SAEB_REQUIRE(m1 >= 1 && m1 <= (k < K2 ? k : K2), "candidate_bounds_packed: need 1 <= m1 <= k");
  const RefineWs w = refine_ws(Tc, d, N, k, margin);
  SAEB_REQUIRE(workspace_bytes >= w.total, "candidate_bounds_packed: workspace too small");
  const PrepLayout p = prep_layout(T_total, d);
  const uint8_t* pb = reinterpret_cast<const uint8_t*>(prep);
  uint8_t* ws = reinterpret_cast<uint8_t*>(workspace);
  const uint8_t* pk = reinterpret_cast<const uint8_t*>(packed);
  const float* wnorm = reinterpret_cast<const float*>(pk + planes_bytes(N, d, 3) + bias_bytes(N));
  const float* dnorm = reinterpret_cast<const float*>(pk + planes_bytes(N, d, 3) + 2 * bias_bytes(N));
  int rc = encode_select_bounds_launch(Tc, N, K2, m1, wnorm, dnorm, reinterpret_cast<const float*>(pb + p.xnorm) + t0,
                                       reinterpret_cast<const float*>(pb + p.xdnorm) + t0, refine_c_eps(x_dtype),
                                       clamp_feature, reinterpret_cast<float*>(ws + w.mvals),
                                       reinterpret_cast<long long*>(ws + w.midx), bounds_out, ws + w.enc,
                                       w.total - w.enc, (cudaStream_t)stream);
  if (rc == 0) g_launches += 1;
  return rc;
}

static int refine_candidates_impl(bool lo, const void* x, int x_dtype, int64_t ld_x, const void* prep, int64_t T_total,
                                  int64_t t0, int64_t Tc, const void* packed, const float* W_enc, int64_t d, int64_t N,
                                  int k, int margin, int64_t clamp_feature, float clamp_value, const float* ext_lower,
                                  const float* ext_upper, const float* feat_thr, int already_merged, float* out_vals,
                                  float* out_member, int64_t* out_idx, int32_t* status_out, void* workspace,
                                  size_t workspace_bytes, int max_ctas, int value_mode, void* stream) {
  g_err[0] = 0;
  SAEB_NVTX("saeb:merge+refine");
  SAEB_REQUIRE(x && prep && packed && W_enc && out_vals && out_idx && workspace, "refine_candidates: null pointer");
  SAEB_REQUIRE(max_ctas >= 0, "refine_candidates: max_ctas must be >= 0");
  SAEB_REQUIRE(value_mode >= 0 && value_mode <= 2, "refine_candidates: value_mode must be 0, 1 or 2");
  SAEB_REQUIRE(ext_upper == nullptr || (value_mode == 2 && ext_lower != nullptr && out_member != nullptr),
               "refine_candidates: ext_upper needs value_mode 2, ext_lower and out_member");
  SAEB_REQUIRE(t0 >= 0 && Tc >= 0 && t0 + Tc <= T_total, "refine_candidates: bad row range");
  SAEB_REQUIRE(already_merged >= 0 && already_merged <= 2, "refine_candidates: already_merged must be 0, 1 or 2");
  if (Tc == 0) return 0;
  const int K2raw = refine_k2(k, margin);
  const int K2 = K2raw < N ? K2raw : (int)N;
  // candidates left UNSORTED by saeb_candidate_bounds_packed: only the warp-per-token scan kernel may read them
  SAEB_REQUIRE(already_merged != 2 || (!lo && value_mode == 2 && ext_lower && ext_upper && out_member && K2 <= 128 &&
                                       scan_warp_enabled()),
               "refine_candidates: already_merged = 2 needs the feature-sharded scan form (value_mode 2, ext_lower, "
               "ext_upper, out_member) with option scan_warp = 1");
  const RefineWs w = refine_ws(Tc, d, N, k, margin);
  SAEB_REQUIRE(workspace_bytes >= w.total, "refine_candidates: workspace too small");
  cudaStream_t st = (cudaStream_t)stream;
  const PrepLayout p = prep_layout(T_total, d);
  const uint8_t* pb = reinterpret_cast<const uint8_t*>(prep);
  uint8_t* ws = reinterpret_cast<uint8_t*>(workspace);
  const uint8_t* pk = reinterpret_cast<const uint8_t*>(packed);
  const float* bias = reinterpret_cast<const float*>(pk + planes_bytes(N, d, 3));
  const float* wnorm = reinterpret_cast<const float*>(pk + planes_bytes(N, d, 3) + bias_bytes(N));
  const float* dnorm = reinterpret_cast<const float*>(pk + planes_bytes(N, d, 3) + 2 * bias_bytes(N));
  const float* trailer = reinterpret_cast<const float*>(pk + planes_bytes(N, d, 3) + 3 * bias_bytes(N));
  int* status = reinterpret_cast<int*>(ws + w.status);
  float* mvals = reinterpret_cast<float*>(ws + w.mvals);
  long long* midx = reinterpret_cast<long long*>(ws + w.midx);
  SAEB_CHECK_CUDA(cudaMemsetAsync(status, 0, 256, st));
  int rc = 0;
  if (!already_merged) {
    rc = encode_merge_launch(Tc, N, K2, mvals, midx, ws + w.enc, w.total - w.enc, max_ctas > 0, st);
    if (rc) return rc;
  }
  rc = refine_launch(x, x_dtype, Tc, ld_x, W_enc, d, N, bias, wnorm, dnorm, trailer,
                     reinterpret_cast<const float*>(pb + p.xnorm) + t0,
                     reinterpret_cast<const float*>(pb + p.xdnorm) + t0, refine_c_eps(x_dtype), mvals, midx, K2,
                     k < K2 ? k : K2, clamp_feature, clamp_value, out_vals, reinterpret_cast<long long*>(out_idx), status,
                     reinterpret_cast<int*>(ws + w.flag_rows), reinterpret_cast<float*>(ws + w.dense), ext_lower,
                     lo ? pk + mode3_bytes(N, d) : nullptr, pad8(d), max_ctas, value_mode, ext_upper, feat_thr,
                     out_member, st);
  if (rc) return rc;
  if (status_out != nullptr)
    SAEB_CHECK_CUDA(cudaMemcpyAsync(status_out, status, sizeof(int), cudaMemcpyDeviceToDevice, st));
  g_launches += 6;
  return 0;
}

int saeb_refine_candidates(const void* x, int x_dtype, int64_t ld_x, const void* prep, int64_t T_total, int64_t t0,
                           int64_t Tc, const void* packed, const float* W_enc, int64_t d, int64_t N, int k, int margin,
                           int64_t clamp_feature, float clamp_value, const float* ext_lower, const float* ext_upper,
                           const float* feat_thr, int already_merged, float* out_vals, float* out_member,
                           int64_t* out_idx, int32_t* status_out, void* workspace, size_t workspace_bytes,
                           int max_ctas, int value_mode, void* stream) {
  return refine_candidates_impl(false, x, x_dtype, ld_x, prep, T_total, t0, Tc, packed, W_enc, d, N, k, margin,
                                clamp_feature, clamp_value, ext_lower, ext_upper, feat_thr, already_merged, out_vals,
                                out_member, out_idx, status_out, workspace, workspace_bytes, max_ctas, value_mode,
                                stream);
}

int saeb_refine_candidates_lo(const void* x, int x_dtype, int64_t ld_x, const void* prep, int64_t T_total, int64_t t0,
                              int64_t Tc, const void* packed4, const float* W_enc, int64_t d, int64_t N, int k,
                              int margin, int64_t clamp_feature, float clamp_value, const float* ext_lower,
                              const float* ext_upper, const float* feat_thr, int already_merged, float* out_vals,
                              float* out_member, int64_t* out_idx, int32_t* status_out, void* workspace,
                              size_t workspace_bytes, int max_ctas, int value_mode, void* stream) {
  return refine_candidates_impl(true, x, x_dtype, ld_x, prep, T_total, t0, Tc, packed4, W_enc, d, N, k, margin,
                                clamp_feature, clamp_value, ext_lower, ext_upper, feat_thr, already_merged, out_vals,
                                out_member, out_idx, status_out, workspace, workspace_bytes, max_ctas, value_mode,
                                stream);
}

size_t saeb_encode_topk_refine_workspace_bytes(int64_t T, int64_t d, int64_t N, int k, int margin) {
  return align_up(prep_layout(T, d).total, 1024) + refine_ws(T, d, N, k, margin).total;
}

int saeb_encode_topk_refine(const void* x, int x_dtype, int64_t T, int64_t ld_x, const void* packed,
                            const float* W_enc, int64_t d, int64_t N, int k, int margin, int64_t clamp_feature,
                            float clamp_value, float* out_vals, int64_t* out_idx, int32_t* status_out, void* workspace,
                            size_t workspace_bytes, int value_mode, void* stream) {
  g_err[0] = 0;
  SAEB_REQUIRE(workspace != nullptr, "encode_topk_refine: null workspace");
  if (T == 0) return 0;
  const size_t prep_bytes = align_up(prep_layout(T, d).total, 1024);
  SAEB_REQUIRE(workspace_bytes >= prep_bytes + refine_ws(T, d, N, k, margin).total,
               "encode_topk_refine: workspace too small");
  uint8_t* ws = reinterpret_cast<uint8_t*>(workspace);
  int rc = saeb_prep_activations(x, x_dtype, T, ld_x, d, ws, stream);
  if (rc) return rc;
  rc = saeb_encode_candidates(ws, T, 0, T, packed, d, N, k, margin, clamp_feature, clamp_value, ws + prep_bytes,
                              workspace_bytes - prep_bytes, stream);
  if (rc) return rc;
  return saeb_refine_candidates(x, x_dtype, ld_x, ws, T, 0, T, packed, W_enc, d, N, k, margin, clamp_feature,
                                clamp_value, nullptr, nullptr, nullptr, 0, out_vals, nullptr, out_idx, status_out,
                                ws + prep_bytes, workspace_bytes - prep_bytes, 0, value_mode, stream);
}

int saeb_dense_topk(const float* dense, int64_t T, int64_t ld, int64_t N, int k, float* out_vals, int64_t* out_idx,
                    void* stream) {
  g_err[0] = 0;
  SAEB_NVTX("saeb:dense_topk");
  SAEB_REQUIRE(dense && out_vals && out_idx, "dense_topk: null pointer");
  int rc = dense_topk_launch(dense, T, ld, N, k, out_vals, reinterpret_cast<long long*>(out_idx), (cudaStream_t)stream);
  if (rc == 0 && T > 0) g_launches += 1;
  return rc;
}

int saeb_decode(const int64_t* idx, const float* vals, int64_t T, int k, const void* W_dec, int w_dtype, int64_t d,
                int64_t N, const float* b_dec, void* out, int out_dtype, int64_t ld_out, const void* x, int x_dtype,
                int64_t ld_x, double* sq_err, int* err_flag, int max_ctas, void* stream) {
  g_err[0] = 0;
  SAEB_NVTX("saeb:decode");
  SAEB_REQUIRE(idx && vals && W_dec && out, "decode: null pointer");
  SAEB_REQUIRE(max_ctas >= 0, "decode: max_ctas must be >= 0");
  int rc = decode_launch(reinterpret_cast<const long long*>(idx), vals, T, k, W_dec, w_dtype, d, N, b_dec, out,
                         out_dtype, ld_out, x, x_dtype, ld_x, sq_err, err_flag, max_ctas, (cudaStream_t)stream);
  if (rc == 0 && T > 0) g_launches += 1;
  return rc;
}

int saeb_decode_backward_acts(const float* grad_out, int64_t ld_g, const int64_t* idx, int64_t T, int k,
                              const float* W_dec, int64_t d, int64_t N, float* d_vals, int* err_flag, void* stream) {
  g_err[0] = 0;
  SAEB_NVTX("saeb:decode_bwd_acts");
  SAEB_REQUIRE(grad_out && idx && W_dec && d_vals, "decode_backward_acts: null pointer");
  int rc = decode_bwd_acts_launch(grad_out, ld_g, reinterpret_cast<const long long*>(idx), T, k, W_dec, d, N, d_vals,
                                  err_flag, (cudaStream_t)stream);
  if (rc == 0 && T > 0) g_launches += 1;
  return rc;
}

int saeb_decode_backward_weight(const float* grad_out, int64_t ld_g, const int64_t* idx, const float* vals, int64_t T,
                                int k, int64_t d, int64_t N, float* dW_dec, int* err_flag, void* stream) {
  g_err[0] = 0;
  SAEB_NVTX("saeb:decode_bwd_weight");
  SAEB_REQUIRE(grad_out && idx && vals && dW_dec, "decode_backward_weight: null pointer");
  int rc = decode_bwd_weight_launch(grad_out, ld_g, reinterpret_cast<const long long*>(idx), vals, T, k, d, N, dW_dec,
                                    err_flag, (cudaStream_t)stream);
  if (rc == 0 && T > 0) g_launches += 1;
  return rc;
}

int saeb_total_variance(const void* x, int x_dtype, int64_t T, int64_t d, int64_t ld_x, double* scratch, double* out,
                        void* stream) {
  g_err[0] = 0;
  SAEB_NVTX("saeb:total_variance");
  SAEB_REQUIRE(x && scratch && out, "total_variance: null pointer");
  int rc = total_variance_launch(x, x_dtype, T, d, ld_x, scratch, out, (cudaStream_t)stream);
  if (rc == 0) g_launches += 2;
  return rc;
}

size_t saeb_coo_workspace_bytes(int64_t T) { return coo_workspace_bytes(T); }

int saeb_coo_extract(const float* vals, const int64_t* idx, int64_t T, int k, float threshold,
                     const uint32_t* filter_bitmap, int64_t seq_len, int64_t row_offset, int64_t* locations,
                     float* activations, int64_t* nnz_out, void* workspace, size_t workspace_bytes, void* stream) {
  g_err[0] = 0;
  SAEB_NVTX("saeb:coo_extract");
  SAEB_REQUIRE(vals && idx && locations && activations && nnz_out && workspace, "coo_extract: null pointer");
  int rc = coo_extract_launch(vals, reinterpret_cast<const long long*>(idx), T, k, threshold, filter_bitmap, seq_len,
                              row_offset, reinterpret_cast<long long*>(locations), activations,
                              reinterpret_cast<long long*>(nnz_out), workspace, workspace_bytes, nullptr, 0, nullptr,
                              (cudaStream_t)stream);
  if (rc == 0) g_launches += 5;
  return rc;
}

int saeb_coo_append(const float* vals, const int64_t* idx, int64_t T, int k, float threshold,
                    const uint32_t* filter_bitmap, int64_t seq_len, int64_t row_offset, int64_t* locations,
                    float* activations, int64_t capacity, int64_t* cursor, int* overflow_flag, void* workspace,
                    size_t workspace_bytes, void* stream) {
  g_err[0] = 0;
  SAEB_NVTX("saeb:coo_append");
  SAEB_REQUIRE(vals && idx && locations && activations && cursor && workspace, "coo_append: null pointer");
  SAEB_REQUIRE(capacity >= 0, "coo_append: negative capacity");
  SAEB_REQUIRE(workspace_bytes >= coo_workspace_bytes(T) + 256, "coo_append: workspace too small");
  // the batch's own count lives in the last 256 bytes of the workspace
  uint8_t* ws = reinterpret_cast<uint8_t*>(workspace);
  long long* batch_nnz = reinterpret_cast<long long*>(ws + ((workspace_bytes - 256) & ~(size_t)255));
  int rc = coo_extract_launch(vals, reinterpret_cast<const long long*>(idx), T, k, threshold, filter_bitmap, seq_len,
                              row_offset, reinterpret_cast<long long*>(locations), activations, batch_nnz, workspace,
                              workspace_bytes - 256, reinterpret_cast<long long*>(cursor), capacity, overflow_flag,
                              (cudaStream_t)stream);
  if (rc == 0) g_launches += 6;
  return rc;
}

int saeb_scan_pool(const float* vals, const int64_t* idx, int64_t T, int k, int ctx_len, float threshold,
                   int64_t feat_lo, int64_t feat_hi, int64_t window_base, const float* tok_thr,
                   const float* member, const float* feat_thr, void* bucket, int* bucket_cnt, int bucket_cap,
                   int* overflow_flag, void* stream) {
  g_err[0] = 0;
  SAEB_NVTX("saeb:scan_pool");
  SAEB_REQUIRE(vals && idx && feat_thr && bucket && bucket_cnt, "scan_pool: null pointer");
  int rc = scan_pool_launch(vals, reinterpret_cast<const long long*>(idx), T, k, ctx_len, threshold, feat_lo, feat_hi,
                            window_base, tok_thr, member, feat_thr, bucket, bucket_cnt, bucket_cap, overflow_flag,
                            (cudaStream_t)stream);
  if (rc == 0) g_launches += 1;
  return rc;
}

int saeb_scan_merge(void* bucket, int* bucket_cnt, int bucket_cap, int64_t F, int n_top, float base_threshold,
                    float* top_vals, int64_t* top_win, float* feat_thr, void* stream) {
  g_err[0] = 0;
  SAEB_NVTX("saeb:scan_merge");
  SAEB_REQUIRE(bucket && bucket_cnt && top_vals && top_win && feat_thr, "scan_merge: null pointer");
  int rc = scan_merge_launch(bucket, bucket_cnt, bucket_cap, F, n_top, base_threshold, top_vals,
                             reinterpret_cast<long long*>(top_win), feat_thr, (cudaStream_t)stream);
  if (rc == 0) g_launches += 1;
  return rc;
}

size_t saeb_image_pool_workspace_bytes(int k, int n_base, int64_t F) { return image_pool_workspace_bytes(k, n_base, F); }

int saeb_image_pool_init(void* workspace, size_t workspace_bytes, int k, int n_base, int64_t F, void* stream) {
  g_err[0] = 0;
  return image_pool_init(workspace, workspace_bytes, k, n_base, F, (cudaStream_t)stream);
}

int saeb_image_pool(const float* vals, const int64_t* idx, int64_t n_images, int64_t tokens_per_image, int k,
                    int n_base, float threshold, int64_t feat_lo, int64_t feat_hi, int64_t image_base,
                    const float* tok_thr, const float* feat_thr, void* bucket, int* bucket_cnt, int bucket_cap,
                    int* overflow_flag, void* workspace, size_t workspace_bytes, void* stream) {
  g_err[0] = 0;
  SAEB_NVTX("saeb:image_pool");
  SAEB_REQUIRE(vals && idx && feat_thr && bucket && bucket_cnt && workspace, "image_pool: null pointer");
  int rc = image_pool_launch(vals, reinterpret_cast<const long long*>(idx), n_images, tokens_per_image, k, n_base,
                             threshold, feat_lo, feat_hi, image_base, tok_thr, feat_thr, bucket, bucket_cnt, bucket_cap,
                             overflow_flag, workspace, workspace_bytes, (cudaStream_t)stream);
  if (rc == 0) g_launches += 1;
  return rc;
}

int saeb_kth_largest_gathered(const float* gathered, int R, int64_t T, int m, int kth, float* tok_thr, void* stream) {
  g_err[0] = 0;
  SAEB_NVTX("saeb:kth");
  SAEB_REQUIRE(gathered && tok_thr, "kth_largest_gathered: null pointer");
  if (T == 0) return 0;
  int rc = kth_gathered_launch(gathered, R, T, m, kth, tok_thr, (cudaStream_t)stream);
  if (rc == 0) g_launches += 1;
  return rc;
}

int saeb_debug_coload(int mode, int ctas, int64_t iters, const void* buf, size_t bytes, float* sink, void* stream) {
  g_err[0] = 0;
  return coload_launch(mode, ctas, iters, buf, bytes, sink, (cudaStream_t)stream);
}

int saeb_gathered_bounds(const float* gathered, int R, int64_t T, int m1, int k, float* ext_lower, float* ext_upper,
                         void* stream) {
  g_err[0] = 0;
  SAEB_NVTX("saeb:gathered_bounds");
  SAEB_REQUIRE(gathered && ext_lower && ext_upper, "gathered_bounds: null pointer");
  if (T == 0) return 0;
  int rc = gathered_bounds_launch(gathered, R, T, m1, k, ext_lower, ext_upper, (cudaStream_t)stream);
  if (rc == 0) g_launches += 1;
  return rc;
}

size_t saeb_scan_pool_workspace_bytes(int k, int ctx_len) { return scan_pool_workspace_bytes(k, ctx_len); }

int saeb_scan_pool_init(void* workspace, size_t workspace_bytes, int k, int ctx_len, void* stream) {
  g_err[0] = 0;
  return scan_pool_init(workspace, workspace_bytes, k, ctx_len, (cudaStream_t)stream);
}

int saeb_scan_pool_ws(const float* vals, const int64_t* idx, int64_t T, int k, int ctx_len, float threshold,
                      int64_t feat_lo, int64_t feat_hi, int64_t window_base, const float* tok_thr, const float* member,
                      const float* feat_thr, void* bucket, int* bucket_cnt, int bucket_cap, int* overflow_flag,
                      void* workspace, size_t workspace_bytes, void* stream) {
  g_err[0] = 0;
  SAEB_NVTX("saeb:scan_pool");
  SAEB_REQUIRE(vals && idx && feat_thr && bucket && bucket_cnt && workspace, "scan_pool: null pointer");
  int rc = scan_pool_ws_launch(vals, reinterpret_cast<const long long*>(idx), T, k, ctx_len, threshold, feat_lo, feat_hi,
                               window_base, tok_thr, member, feat_thr, bucket, bucket_cnt, bucket_cap, overflow_flag,
                               workspace, workspace_bytes, (cudaStream_t)stream);
  if (rc == 0) g_launches += 1;
  return rc;
}

int saeb_coo_window_scores(const int64_t* feature, const int64_t* window_key, const float* activations, int64_t nnz,
                           int mode, float divisor, float* score, int* head, void* stream) {
  g_err[0] = 0;
  SAEB_NVTX("saeb:coo_window_scores");
  SAEB_REQUIRE(feature && window_key && activations && score && head, "coo_window_scores: null pointer");
  int rc = coo_window_scores_launch(reinterpret_cast<const long long*>(feature),
                                    reinterpret_cast<const long long*>(window_key), activations, nnz, mode, divisor, score,
                                    head, (cudaStream_t)stream);
  if (rc == 0 && nnz > 0) g_launches += 1;
  return rc;
}

int saeb_column_sums(const float* dense, int64_t T, int64_t ld, int64_t N, double* colsum, void* stream) {
  g_err[0] = 0;
  SAEB_NVTX("saeb:column_sums");
  SAEB_REQUIRE(dense && colsum, "column_sums: null pointer");
  int rc = column_sums_launch(dense, T, ld, N, colsum, (cudaStream_t)stream);
  if (rc == 0 && T > 0) g_launches += 1;
  return rc;
}

int saeb_feature_maps(const void* x, int x_dtype, int64_t T, int64_t ld_x, const float* W_enc, const float* b_enc,
                      const float* b_dec, int64_t d, int64_t N, const int64_t* features, int n_features, float* out,
                      int* err_flag, void* stream) {
  g_err[0] = 0;
  SAEB_NVTX("saeb:feature_maps");
  SAEB_REQUIRE(x && W_enc && b_enc && b_dec && features && out, "feature_maps: null pointer");
  int rc = feature_maps_launch(x, x_dtype, T, ld_x, W_enc, b_enc, b_dec, d, N,
                               reinterpret_cast<const long long*>(features), n_features, out, err_flag,
                               (cudaStream_t)stream);
  if (rc == 0 && T > 0 && n_features > 0) g_launches += 1;
  return rc;
}

int saeb_kth_of_gathered(const float* gathered, int R, int64_t T, int k, float* tok_thr, void* stream) {
  return saeb_kth_largest_gathered(gathered, R, T, k, k, tok_thr, stream);
}

int saeb_push_gather(const void* src, size_t bytes, void* const* peer_bases_dev, int R, int self_rank,
                     size_t region_offset, void* multicast_base, size_t flags_offset, int channel, uint32_t seq,
                     int* counter, void* stream) {
  g_err[0] = 0;
  SAEB_NVTX("saeb:exchange_push");
  SAEB_REQUIRE(src && peer_bases_dev && counter, "push_gather: null pointer");
  if (bytes == 0) return 0;
  int rc = push_gather_launch(src, bytes, peer_bases_dev, R, self_rank, region_offset, multicast_base, flags_offset,
                              channel, seq, counter, (cudaStream_t)stream);
  if (rc == 0) g_launches += 1;
  return rc;
}

}  // extern "C"
