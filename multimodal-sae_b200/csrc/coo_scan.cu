// Feature-activation cache extraction and per-feature top-activation scan, working directly on TopK output.
//
// coo_*      : replaces `zeros_like + scatter_` (features/cache.py:215-217) + Cache.get_nonzeros
//              (features/cache.py:73-92: nonzero(|x|>1e-5), mask gather, torch.isin filter) -- emits the same
//              (row, pos, feature) int64 triples / fp32 activations in the same row-major order, without the dense
//              [tokens, N] tensors.
// scan_*     : replaces, for every feature at once, TensorBuffer.__getitem__ (features/loader.py:74-90) +
//              pool_max_activation_windows (features/constructors.py:11-85): window max-pool of the TopK-masked
//              activations and the `n_top` best windows per feature.
// kth_*      : per-token global k-th value from all-gathered per-shard top-k values (feature-sharded exactness,
//              SURVEY.md section 8(e)).
#include "kernels_coo_scan.cuh"
#include "kernels_kth.cuh"

namespace saeb {

size_t coo_workspace_bytes(long long T) {
  const long long nchunks = (T + SCAN_CHUNK - 1) / SCAN_CHUNK;
  return (size_t)T * sizeof(int) + 256 + (size_t)(T + 1) * sizeof(long long) + 256 +
         (size_t)(nchunks + 1) * sizeof(long long) + 256;
}

// cursor == nullptr: triples are written from entry 0 and *nnz_out = their number (saeb_coo_extract).
// cursor != nullptr (device, in/out): they are appended behind *cursor in an arena of `capacity` entries and the cursor
// advances -- no host round trip between batches (saeb_coo_append); nnz_out then is scratch for the batch's count.
int coo_extract_launch(const float* vals, const long long* idx, long long T, int k, float threshold,
                       const uint32_t* filter, long long seq_len, long long row_offset, long long* locations,
                       float* activations, long long* nnz_out, void* ws, size_t ws_bytes, long long* cursor,
                       long long capacity, int* overflow, cudaStream_t stream) {
  SAEB_REQUIRE(T > 0 && k >= 1 && k <= 1024 && seq_len > 0, "coo: bad arguments T=%lld k=%d seq_len=%lld", T, k,
               seq_len);
  SAEB_REQUIRE(ws_bytes >= coo_workspace_bytes(T), "coo: workspace too small");
  uint8_t* p = reinterpret_cast<uint8_t*>(ws);
  int* counts = reinterpret_cast<int*>(p);
  p += (((size_t)T * sizeof(int)) + 255) & ~(size_t)255;
  long long* offsets = reinterpret_cast<long long*>(p);
  p += (((size_t)(T + 1) * sizeof(long long)) + 255) & ~(size_t)255;
  long long* sums = reinterpret_cast<long long*>(p);
  const long long nchunks = (T + SCAN_CHUNK - 1) / SCAN_CHUNK;
  const int wpb = 8;
  coo_count_kernel<<<(unsigned)((T + wpb - 1) / wpb), wpb * 32, 0, stream>>>(vals, idx, T, k, threshold, filter, counts);
  SAEB_CHECK_CUDA(cudaGetLastError());
  scan_chunk_sum_kernel<<<(unsigned)nchunks, SCAN_CHUNK, 0, stream>>>(counts, T, sums);
  scan_sums_kernel<<<1, SCAN_CHUNK, 0, stream>>>(sums, nchunks, nnz_out);
  scan_chunks_kernel<<<(unsigned)nchunks, SCAN_CHUNK, 0, stream>>>(counts, T, sums, offsets);
  SAEB_CHECK_CUDA(cudaGetLastError());
  int kp2 = 2;
  while (kp2 < k) kp2 <<= 1;
  coo_emit_kernel<<<(unsigned)((T + wpb - 1) / wpb), wpb * 32, (size_t)wpb * kp2 * sizeof(uint2), stream>>>(
      vals, idx, T, k, kp2, threshold, filter, offsets, seq_len, row_offset, locations, activations, cursor, capacity,
      overflow);
  SAEB_CHECK_CUDA(cudaGetLastError());
  if (cursor != nullptr) {
    coo_advance_kernel<<<1, 32, 0, stream>>>(cursor, nnz_out, capacity);
    SAEB_CHECK_CUDA(cudaGetLastError());
  }
  return 0;
}

// ---------------------------------------------------------------------------------------------
// per-token kth-largest of R gathered per-shard value lists  [R][T][m]  (m values per shard and token)
//
// One warp per token.  The R*m values are loaded ONCE into registers (VPL per lane, coalesced along a shard's row) and
// the kth-largest bit pattern is found by a 31-step bit search over the registers: the scan calls this twice per token
// chunk between two collectives, so it sits on the critical path of the multi-GPU schedule (the first version re-read
// the values from memory in every search step: ~0.8 ms per 37 888-token chunk at R*m = 512).
// ---------------------------------------------------------------------------------------------
static thread_local int g_kth_impl = 1;   // 1: register-resident search (default); 0: memory-resident (first version, diagnostics)
int set_kth_impl(int v) {
  g_kth_impl = v ? 1 : 0;
  return 0;
}

int kth_gathered_launch(const float* gathered, int R, long long T, int m, int kth, float* tok_thr,
                        cudaStream_t stream) {
  SAEB_REQUIRE(R >= 1 && T > 0 && m >= 1, "kth_gathered: bad arguments R=%d T=%lld m=%d", R, T, m);
  SAEB_REQUIRE((long long)R * m < (1ll << 30), "kth_gathered: R*m too large");
  SAEB_REQUIRE(kth >= 1 && kth <= R * m, "kth_gathered: kth=%d must be in 1..R*m=%d", kth, R * m);
  const int wpb = 8;
  const unsigned blocks = (unsigned)((T + wpb - 1) / wpb);
  const int M = R * m;
  if (g_kth_impl == 1 && M <= 128)
    SAEB_CARVEOUT(kth_gathered_reg_kernel<4>), kth_gathered_reg_kernel<4><<<blocks, wpb * 32, 0, stream>>>(gathered, R, T, m, kth, tok_thr);
  else if (g_kth_impl == 1 && M <= 512)
    SAEB_CARVEOUT(kth_gathered_reg_kernel<16>), kth_gathered_reg_kernel<16><<<blocks, wpb * 32, 0, stream>>>(gathered, R, T, m, kth, tok_thr);
  else if (g_kth_impl == 1 && M <= 2048)
    SAEB_CARVEOUT(kth_gathered_reg_kernel<64>), kth_gathered_reg_kernel<64><<<blocks, wpb * 32, 0, stream>>>(gathered, R, T, m, kth, tok_thr);
  else
    SAEB_CARVEOUT(kth_gathered_mem_kernel), kth_gathered_mem_kernel<<<blocks, wpb * 32, 0, stream>>>(gathered, R, T, m, kth, tok_thr);
  SAEB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int gathered_bounds_launch(const float* gathered, int R, long long T, int m1, int k, float* ext_L, float* ext_U,
                           cudaStream_t stream) {
  SAEB_REQUIRE(R >= 1 && T > 0 && m1 >= 1 && k >= 1, "gathered_bounds: bad arguments R=%d T=%lld m1=%d k=%d", R, T, m1, k);
  const int M = R * m1;
  SAEB_REQUIRE(M <= 2048, "gathered_bounds: R*m1=%d too large (max 2048)", M);
  const int wpb = 4;
  const unsigned blocks = (unsigned)((T + wpb - 1) / wpb);
  if (M <= 128)
    SAEB_CARVEOUT(gathered_bounds_kernel<4>), gathered_bounds_kernel<4><<<blocks, wpb * 32, 0, stream>>>(gathered, R, T, m1, k, ext_L, ext_U);
  else if (M <= 256)
    SAEB_CARVEOUT(gathered_bounds_kernel<8>), gathered_bounds_kernel<8><<<blocks, wpb * 32, 0, stream>>>(gathered, R, T, m1, k, ext_L, ext_U);
  else if (M <= 512)
    SAEB_CARVEOUT(gathered_bounds_kernel<16>), gathered_bounds_kernel<16><<<blocks, wpb * 32, 0, stream>>>(gathered, R, T, m1, k, ext_L, ext_U);
  else
    SAEB_CARVEOUT(gathered_bounds_kernel<64>), gathered_bounds_kernel<64><<<blocks, wpb * 32, 0, stream>>>(gathered, R, T, m1, k, ext_L, ext_U);
  SAEB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

// ---------------------------------------------------------------------------------------------
// top-activation scan
// ---------------------------------------------------------------------------------------------
int scan_pool_launch(const float* vals, const long long* idx, long long T, int k, int ctx_len, float threshold,
                     long long feat_lo, long long feat_hi, long long window_base, const float* tok_thr,
                     const float* member, const float* feat_thr, void* bucket, int* bucket_cnt, int bucket_cap,
                     int* overflow, cudaStream_t stream) {
  SAEB_REQUIRE(T > 0 && k >= 1 && ctx_len >= 1, "scan_pool: bad arguments");
  const long long n_win = (T + ctx_len - 1) / ctx_len;
  SAEB_REQUIRE(n_win <= bucket_cap, "scan_pool: %lld windows per call exceed the bucket capacity %d", n_win,
               bucket_cap);
  long long ent = (long long)ctx_len * k;
  int slots = 64;
  while (slots < 2 * ent) slots <<= 1;
  SAEB_REQUIRE(slots <= 16384 * 1, "scan_pool: ctx_len*k=%lld too large for the shared-memory hash (max 8192)", ent);
  const size_t smem = (size_t)slots * 8;
  SAEB_CHECK_CUDA(cudaFuncSetAttribute(scan_pool_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  SAEB_CARVEOUT(scan_pool_kernel);
  scan_pool_kernel<<<(unsigned)n_win, 256, smem, stream>>>(vals, idx, T, k, ctx_len, threshold, feat_lo, feat_hi,
                                                          window_base, tok_thr, member, feat_thr,
                                                          reinterpret_cast<uint2*>(bucket), bucket_cnt, bucket_cap,
                                                          slots, overflow);
  SAEB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

// ---- global-hash form of scan_pool (no shared memory: runs beside a resident GEMM CTA) ----
struct ScanPoolPlan {
  int grid, slots;
  size_t keys_off, vals_off, list_off, total;
};
static ScanPoolPlan scan_pool_plan(int k, int ctx_len) {
  ScanPoolPlan p;
  const long long ent = (long long)ctx_len * k;
  p.slots = 64;
  while (p.slots < 2 * ent) p.slots <<= 1;
  int sms = 0, dev = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  p.grid = 2 * (sms > 0 ? sms : 148);
  const size_t tab = (size_t)p.grid * p.slots * 4;
  p.vals_off = 0;
  p.keys_off = tab;
  p.list_off = 2 * tab;
  p.total = 3 * tab;
  return p;
}
size_t scan_pool_workspace_bytes(int k, int ctx_len) { return scan_pool_plan(k, ctx_len).total; }

int scan_pool_init(void* ws, size_t ws_bytes, int k, int ctx_len, cudaStream_t stream) {
  const ScanPoolPlan p = scan_pool_plan(k, ctx_len);
  SAEB_REQUIRE(ws != nullptr && ws_bytes >= p.total, "scan_pool: workspace too small");
  uint8_t* b = reinterpret_cast<uint8_t*>(ws);
  SAEB_CHECK_CUDA(cudaMemsetAsync(b + p.vals_off, 0, p.keys_off - p.vals_off, stream));
  SAEB_CHECK_CUDA(cudaMemsetAsync(b + p.keys_off, 0xff, p.list_off - p.keys_off, stream));   // HASH_EMPTY
  return 0;
}

int scan_pool_ws_launch(const float* vals, const long long* idx, long long T, int k, int ctx_len, float threshold,
                        long long feat_lo, long long feat_hi, long long window_base, const float* tok_thr,
                        const float* member, const float* feat_thr, void* bucket, int* bucket_cnt, int bucket_cap,
                        int* overflow, void* ws, size_t ws_bytes, cudaStream_t stream) {
  SAEB_REQUIRE(T > 0 && k >= 1 && ctx_len >= 1, "scan_pool: bad arguments");
  const long long n_win = (T + ctx_len - 1) / ctx_len;
  SAEB_REQUIRE(n_win <= bucket_cap, "scan_pool: %lld windows per call exceed the bucket capacity %d", n_win,
               bucket_cap);
  const ScanPoolPlan p = scan_pool_plan(k, ctx_len);
  SAEB_REQUIRE(ws != nullptr && ws_bytes >= p.total, "scan_pool: workspace too small");
  uint8_t* b = reinterpret_cast<uint8_t*>(ws);
  const int grid = n_win < p.grid ? (int)n_win : p.grid;
  SAEB_CARVEOUT(scan_pool_g_kernel);
  scan_pool_g_kernel<<<grid, 128, 0, stream>>>(vals, idx, T, k, ctx_len, threshold, feat_lo, feat_hi, window_base,
                                               tok_thr, member, feat_thr, reinterpret_cast<uint2*>(bucket), bucket_cnt,
                                               bucket_cap, p.slots, reinterpret_cast<uint32_t*>(b + p.keys_off),
                                               reinterpret_cast<uint32_t*>(b + p.vals_off),
                                               reinterpret_cast<int*>(b + p.list_off), n_win, overflow);
  SAEB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int scan_merge_launch(void* bucket, int* bucket_cnt, int bucket_cap, long long F, int n_top, float base_thr,
                      float* top_vals, long long* top_win, float* feat_thr, cudaStream_t stream) {
  SAEB_REQUIRE(F > 0 && n_top >= 1 && n_top <= 1024, "scan_merge: bad arguments");
  int sort_n = 2;
  while (sort_n < n_top + bucket_cap) sort_n <<= 1;
  // 4 warps per block: 16 KB at the default bucket capacity, small enough to be scheduled beside a resident GEMM CTA
  const int wpb = ((size_t)8 * sort_n * sizeof(uint2) > 16 * 1024) ? 4 : 8;
  const size_t smem = (size_t)wpb * sort_n * sizeof(uint2);
  SAEB_REQUIRE(smem <= 200 * 1024, "scan_merge: n_top + bucket_cap too large");
  SAEB_CHECK_CUDA(cudaFuncSetAttribute(scan_merge_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  SAEB_CARVEOUT(scan_merge_kernel);
  scan_merge_kernel<<<(unsigned)((F + wpb - 1) / wpb), wpb * 32, smem, stream>>>(
      reinterpret_cast<uint2*>(bucket), bucket_cnt, bucket_cap, F, n_top, sort_n, base_thr, top_vals, top_win,
      feat_thr);
  SAEB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

// ---------------------------------------------------------------------------------------------
// image scan
// ---------------------------------------------------------------------------------------------
struct ImagePoolPlan {
  int grid, slots;
  size_t keys_off, sums_off, list_off, total;
};
static ImagePoolPlan image_pool_plan(int k, int n_base, long long F) {
  ImagePoolPlan p;
  long long distinct = (long long)n_base * k;   // an image cannot touch more features than it has TopK entries ...
  if (distinct > F) distinct = F;               // ... nor more than the shard owns
  p.slots = 64;
  while (p.slots < 2 * distinct) p.slots <<= 1;
  int sms = 0, dev = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  p.grid = sms > 0 ? sms : 148;
  const size_t n = (size_t)p.grid * p.slots;
  p.sums_off = 0;
  p.keys_off = n * sizeof(unsigned long long);
  p.list_off = p.keys_off + n * sizeof(uint32_t);
  p.total = p.list_off + n * sizeof(int);
  return p;
}
size_t image_pool_workspace_bytes(int k, int n_base, long long F) { return image_pool_plan(k, n_base, F).total; }

int image_pool_init(void* ws, size_t ws_bytes, int k, int n_base, long long F, cudaStream_t stream) {
  const ImagePoolPlan p = image_pool_plan(k, n_base, F);
  SAEB_REQUIRE(ws != nullptr && ws_bytes >= p.total, "image_pool: workspace too small");
  uint8_t* b = reinterpret_cast<uint8_t*>(ws);
  SAEB_CHECK_CUDA(cudaMemsetAsync(b + p.sums_off, 0, p.keys_off - p.sums_off, stream));
  SAEB_CHECK_CUDA(cudaMemsetAsync(b + p.keys_off, 0xff, p.list_off - p.keys_off, stream));   // HASH_EMPTY
  return 0;
}

int image_pool_launch(const float* vals, const long long* idx, long long n_images, long long tokens_per_image, int k,
                      int n_base, float threshold, long long feat_lo, long long feat_hi, long long image_base,
                      const float* tok_thr, const float* feat_thr, void* bucket, int* bucket_cnt, int bucket_cap,
                      int* overflow, void* ws, size_t ws_bytes, cudaStream_t stream) {
  SAEB_REQUIRE(n_images > 0 && k >= 1 && n_base >= 1 && tokens_per_image >= n_base,
               "image_pool: need n_images > 0, k >= 1 and 1 <= n_base <= tokens_per_image");
  SAEB_REQUIRE(n_images <= bucket_cap, "image_pool: %lld images per call exceed the bucket capacity %d", n_images,
               bucket_cap);
  SAEB_REQUIRE(image_base >= 0 && image_base + n_images < (1ll << 32), "image_pool: image ids must fit 32 bits");
  const ImagePoolPlan p = image_pool_plan(k, n_base, feat_hi - feat_lo);
  SAEB_REQUIRE(ws != nullptr && ws_bytes >= p.total, "image_pool: workspace too small");
  uint8_t* b = reinterpret_cast<uint8_t*>(ws);
  const int grid = n_images < p.grid ? (int)n_images : p.grid;
  SAEB_CARVEOUT(image_pool_kernel);
  image_pool_kernel<<<grid, 256, 0, stream>>>(vals, idx, n_images, tokens_per_image, k, n_base, threshold, feat_lo,
                                              feat_hi, image_base, tok_thr, feat_thr,
                                              reinterpret_cast<uint32_t*>(b + p.keys_off),
                                              reinterpret_cast<unsigned long long*>(b + p.sums_off),
                                              reinterpret_cast<int*>(b + p.list_off), p.slots,
                                              reinterpret_cast<uint2*>(bucket), bucket_cnt, bucket_cap, overflow);
  SAEB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int coo_window_scores_launch(const long long* feat, const long long* key, const float* act, long long nnz, int mode,
                             float divisor, float* score, int* head, cudaStream_t stream) {
  SAEB_REQUIRE(nnz >= 0 && (mode == 0 || mode == 1), "coo_window_scores: bad arguments");
  if (nnz == 0) return 0;
  coo_window_scores_kernel<<<(unsigned)((nnz + 255) / 256), 256, 0, stream>>>(feat, key, act, nnz, mode, divisor, score, head);
  SAEB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace saeb
