// extern "C" surface declared in include/saeb200.h: argument checking, workspace carving, launch accounting.
#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstring>

#include "../../include/saeb200.h"
#include "common.cuh"

namespace saeb {

static thread_local char g_err[512] = {0};
static std::atomic<long long> g_launches{0};

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
                       cudaStream_t stream);
int pack_weights_launch(const float* W_enc, const float* b_enc, const float* b_dec, long long N, long long d,
                        long long d_pad, int planes, void* w_planes, float* bias, cudaStream_t stream);
int split_x_launch(const void* x, int x_dtype, long long T, long long d, long long ld_x, long long d_pad, int planes,
                   void* out, cudaStream_t stream);
int decode_launch(const long long* idx, const float* vals, long long T, int k, const void* W_dec, int w_dtype,
                  long long d, long long N, const float* b_dec, void* out, int out_dtype, long long ld_out,
                  const void* x, int x_dtype, long long ld_x, double* sq_err, int* err_flag, cudaStream_t stream);
int total_variance_launch(const void* x, int x_dtype, long long T, long long d, long long ld_x, double* scratch,
                          double* out, cudaStream_t stream);
size_t coo_workspace_bytes(long long T);
int coo_extract_launch(const float* vals, const long long* idx, long long T, int k, float threshold,
                       const uint32_t* filter, long long seq_len, long long row_offset, long long* locations,
                       float* activations, long long* nnz_out, void* ws, size_t ws_bytes, cudaStream_t stream);
int kth_gathered_launch(const float* gathered, int R, long long T, int k, float* tok_thr, cudaStream_t stream);
int scan_pool_launch(const float* vals, const long long* idx, long long T, int k, int ctx_len, float threshold,
                     long long feat_lo, long long feat_hi, long long window_base, const float* tok_thr,
                     const float* feat_thr, void* bucket, int* bucket_cnt, int bucket_cap, int* overflow,
                     cudaStream_t stream);
int scan_merge_launch(void* bucket, int* bucket_cnt, int bucket_cap, long long F, int n_top, float base_thr,
                      float* top_vals, long long* top_win, float* feat_thr, cudaStream_t stream);

int set_cta_pair(int v);
int set_profile(int v);
float last_encode_ms();

static inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }
static inline long long pad8(long long d) { return (d + 7) / 8 * 8; }   // 16-byte row strides for TMA
static inline size_t planes_bytes(long long N, long long d, int planes) {
  return align_up((size_t)planes * (size_t)N * (size_t)pad8(d) * 2, 256);
}
// workspace for bf16 activation planes: none for bf16 input consumed in place (d % 8 == 0), one padded plane for
// bf16 with d % 8 != 0, two planes for fp16 / fp32 input
static inline size_t x_split_bytes(long long T, long long d, int x_dtype) {
  if (x_dtype == DT_BF16 && d % 8 == 0) return 0;
  return align_up((size_t)(x_dtype == DT_BF16 ? 1 : 2) * (size_t)T * (size_t)pad8(d) * 2, 1024);
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
  set_error("set_option: unknown option '%s'", name);
  return -1;
}

float saeb_profile_last_encode_ms(void) { return last_encode_ms(); }

size_t saeb_packed_bias_offset(int64_t N, int64_t d, int planes) { return planes_bytes(N, d, planes); }
size_t saeb_packed_weights_bytes(int64_t N, int64_t d, int planes) {
  return planes_bytes(N, d, planes) + align_up((size_t)N * sizeof(float), 256);
}

int saeb_pack_weights(const float* W_enc, const float* b_enc, const float* b_dec, int64_t N, int64_t d, int planes,
                      void* packed, void* stream) {
  g_err[0] = 0;
  SAEB_REQUIRE(W_enc && b_enc && b_dec && packed, "pack_weights: null pointer");
  float* bias = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(packed) + planes_bytes(N, d, planes));
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
  SAEB_REQUIRE(x && packed, "encode_topk: null pointer");
  SAEB_REQUIRE((out_vals != nullptr) == (out_idx != nullptr), "encode_topk: out_vals and out_idx go together");
  SAEB_REQUIRE(out_vals != nullptr || dense_out != nullptr, "encode_topk: nothing to compute");
  SAEB_REQUIRE(planes == 1 || planes == 2, "encode_topk: planes must be 1 or 2");
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
                              reinterpret_cast<long long*>(out_idx), dense_out, ld_dense, ws, ws_left, pass_mask, st);
  if (rc == 0) g_launches += out_vals ? 2 : 1;
  return rc;
}

int saeb_decode(const int64_t* idx, const float* vals, int64_t T, int k, const void* W_dec, int w_dtype, int64_t d,
                int64_t N, const float* b_dec, void* out, int out_dtype, int64_t ld_out, const void* x, int x_dtype,
                int64_t ld_x, double* sq_err, int* err_flag, void* stream) {
  g_err[0] = 0;
  SAEB_REQUIRE(idx && vals && W_dec && out, "decode: null pointer");
  int rc = decode_launch(reinterpret_cast<const long long*>(idx), vals, T, k, W_dec, w_dtype, d, N, b_dec, out,
                         out_dtype, ld_out, x, x_dtype, ld_x, sq_err, err_flag, (cudaStream_t)stream);
  if (rc == 0 && T > 0) g_launches += 1;
  return rc;
}

int saeb_total_variance(const void* x, int x_dtype, int64_t T, int64_t d, int64_t ld_x, double* scratch, double* out,
                        void* stream) {
  g_err[0] = 0;
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
  SAEB_REQUIRE(vals && idx && locations && activations && nnz_out && workspace, "coo_extract: null pointer");
  int rc = coo_extract_launch(vals, reinterpret_cast<const long long*>(idx), T, k, threshold, filter_bitmap, seq_len,
                              row_offset, reinterpret_cast<long long*>(locations), activations,
                              reinterpret_cast<long long*>(nnz_out), workspace, workspace_bytes, (cudaStream_t)stream);
  if (rc == 0) g_launches += 5;
  return rc;
}

int saeb_scan_pool(const float* vals, const int64_t* idx, int64_t T, int k, int ctx_len, float threshold,
                   int64_t feat_lo, int64_t feat_hi, int64_t window_base, const float* tok_thr,
                   const float* feat_thr, void* bucket, int* bucket_cnt, int bucket_cap, int* overflow_flag,
                   void* stream) {
  g_err[0] = 0;
  SAEB_REQUIRE(vals && idx && feat_thr && bucket && bucket_cnt, "scan_pool: null pointer");
  int rc = scan_pool_launch(vals, reinterpret_cast<const long long*>(idx), T, k, ctx_len, threshold, feat_lo, feat_hi,
                            window_base, tok_thr, feat_thr, bucket, bucket_cnt, bucket_cap, overflow_flag,
                            (cudaStream_t)stream);
  if (rc == 0) g_launches += 1;
  return rc;
}

int saeb_scan_merge(void* bucket, int* bucket_cnt, int bucket_cap, int64_t F, int n_top, float base_threshold,
                    float* top_vals, int64_t* top_win, float* feat_thr, void* stream) {
  g_err[0] = 0;
  SAEB_REQUIRE(bucket && bucket_cnt && top_vals && top_win && feat_thr, "scan_merge: null pointer");
  int rc = scan_merge_launch(bucket, bucket_cnt, bucket_cap, F, n_top, base_threshold, top_vals,
                             reinterpret_cast<long long*>(top_win), feat_thr, (cudaStream_t)stream);
  if (rc == 0) g_launches += 1;
  return rc;
}

int saeb_kth_of_gathered(const float* gathered, int R, int64_t T, int k, float* tok_thr, void* stream) {
  g_err[0] = 0;
  SAEB_REQUIRE(gathered && tok_thr, "kth_of_gathered: null pointer");
  int rc = kth_gathered_launch(gathered, R, T, k, tok_thr, (cudaStream_t)stream);
  if (rc == 0) g_launches += 1;
  return rc;
}

}  // extern "C"
