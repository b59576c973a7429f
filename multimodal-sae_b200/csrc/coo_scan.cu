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
#include "kernels_kth.cuh"

namespace saeb {

// ---------------------------------------------------------------------------------------------
// COO extraction
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ bool coo_keep(float v, long long f, float threshold, const uint32_t* filter) {
  if (!(fabsf(v) > threshold)) return false;
  if (filter != nullptr && !((filter[f >> 5] >> (f & 31)) & 1u)) return false;
  return true;
}

// one warp per token: number of surviving entries
__global__ void coo_count_kernel(const float* __restrict__ vals, const long long* __restrict__ idx, long long T, int k,
                                 float threshold, const uint32_t* __restrict__ filter, int* __restrict__ counts) {
  const long long t = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (t >= T) return;
  int c = 0;
  for (int j = lane; j < k; j += 32) c += coo_keep(vals[t * k + j], idx[t * k + j], threshold, filter) ? 1 : 0;
  c = __reduce_add_sync(0xffffffffu, c);
  if (lane == 0) counts[t] = c;
}

// exclusive scan of int counts -> long long offsets[n+1]; three small kernels (chunk sums, scan of sums, chunk scans)
constexpr int SCAN_CHUNK = 1024;
__global__ void scan_chunk_sum_kernel(const int* __restrict__ in, long long n, long long* __restrict__ sums) {
  __shared__ long long red[32];
  const long long i = (long long)blockIdx.x * SCAN_CHUNK + threadIdx.x;
  long long v = (i < n) ? in[i] : 0;
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  if (threadIdx.x < 32) {
    long long s = red[threadIdx.x];
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (threadIdx.x == 0) sums[blockIdx.x] = s;
  }
}
__global__ void scan_sums_kernel(long long* __restrict__ sums, long long nchunks, long long* __restrict__ total) {
  // single thread block, sequential over chunks of 1024 with a block-wide scan
  __shared__ long long buf[SCAN_CHUNK];
  __shared__ long long carry;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (long long base = 0; base < nchunks; base += SCAN_CHUNK) {
    const long long i = base + threadIdx.x;
    const long long v = (i < nchunks) ? sums[i] : 0;
    buf[threadIdx.x] = v;
    __syncthreads();
    for (int o = 1; o < SCAN_CHUNK; o <<= 1) {
      long long a = (threadIdx.x >= o) ? buf[threadIdx.x - o] : 0;
      __syncthreads();
      buf[threadIdx.x] += a;
      __syncthreads();
    }
    if (i < nchunks) sums[i] = carry + buf[threadIdx.x] - v;   // exclusive
    __syncthreads();
    if (threadIdx.x == 0) carry += buf[SCAN_CHUNK - 1];
    __syncthreads();
  }
  if (threadIdx.x == 0) *total = carry;
}
__global__ void scan_chunks_kernel(const int* __restrict__ in, long long n, const long long* __restrict__ sums,
                                   long long* __restrict__ out) {
  __shared__ long long buf[SCAN_CHUNK];
  const long long i = (long long)blockIdx.x * SCAN_CHUNK + threadIdx.x;
  const long long v = (i < n) ? in[i] : 0;
  buf[threadIdx.x] = v;
  __syncthreads();
  for (int o = 1; o < SCAN_CHUNK; o <<= 1) {
    long long a = (threadIdx.x >= o) ? buf[threadIdx.x - o] : 0;
    __syncthreads();
    buf[threadIdx.x] += a;
    __syncthreads();
  }
  if (i < n) out[i] = sums[blockIdx.x] + buf[threadIdx.x] - v;
  if (i == n - 1) out[n] = sums[blockIdx.x] + buf[threadIdx.x];
}

// one warp per token: filter, sort by feature id ascending (torch.nonzero order), write triples
__global__ void coo_emit_kernel(const float* __restrict__ vals, const long long* __restrict__ idx, long long T, int k,
                                int kp2, float threshold, const uint32_t* __restrict__ filter,
                                const long long* __restrict__ offsets, long long seq_len, long long row_offset,
                                long long* __restrict__ locations, float* __restrict__ activations) {
  extern __shared__ uint2 esm[];
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const long long t = (long long)blockIdx.x * (blockDim.x >> 5) + warp;
  if (t >= T) return;
  uint2* e = esm + (size_t)warp * kp2;
  for (int j = lane; j < kp2; j += 32) {
    uint2 ent = make_uint2(0xffffffffu, 0u);
    if (j < k) {
      const float v = vals[t * k + j];
      const long long f = idx[t * k + j];
      if (coo_keep(v, f, threshold, filter)) ent = make_uint2((uint32_t)f, __float_as_uint(v));
    }
    e[j] = ent;
  }
  __syncwarp();
  for (int size = 2; size <= kp2; size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      for (int p = lane; p < (kp2 >> 1); p += 32) {
        const int lo = ((p / stride) * stride * 2) + (p % stride);
        const int hi = lo + stride;
        const bool asc = ((lo & size) == 0);
        const uint2 a = e[lo], b = e[hi];
        if ((a.x > b.x) == asc) {
          e[lo] = b;
          e[hi] = a;
        }
      }
      __syncwarp();
    }
  }
  const long long off = offsets[t];
  const int cnt = (int)(offsets[t + 1] - off);
  const long long r = row_offset + t / seq_len, pos = t % seq_len;
  for (int j = lane; j < cnt; j += 32) {
    locations[(off + j) * 3 + 0] = r;
    locations[(off + j) * 3 + 1] = pos;
    locations[(off + j) * 3 + 2] = (long long)e[j].x;
    activations[off + j] = __uint_as_float(e[j].y);
  }
}

size_t coo_workspace_bytes(long long T) {
  const long long nchunks = (T + SCAN_CHUNK - 1) / SCAN_CHUNK;
  return (size_t)T * sizeof(int) + 256 + (size_t)(T + 1) * sizeof(long long) + 256 +
         (size_t)(nchunks + 1) * sizeof(long long) + 256;
}

int coo_extract_launch(const float* vals, const long long* idx, long long T, int k, float threshold,
                       const uint32_t* filter, long long seq_len, long long row_offset, long long* locations,
                       float* activations, long long* nnz_out, void* ws, size_t ws_bytes, cudaStream_t stream) {
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
      vals, idx, T, k, kp2, threshold, filter, offsets, seq_len, row_offset, locations, activations);
  SAEB_CHECK_CUDA(cudaGetLastError());
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
static int g_kth_impl = 1;   // 1: register-resident search (default); 0: memory-resident (first version, diagnostics)
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
    kth_gathered_reg_kernel<4><<<blocks, wpb * 32, 0, stream>>>(gathered, R, T, m, kth, tok_thr);
  else if (g_kth_impl == 1 && M <= 512)
    kth_gathered_reg_kernel<16><<<blocks, wpb * 32, 0, stream>>>(gathered, R, T, m, kth, tok_thr);
  else if (g_kth_impl == 1 && M <= 2048)
    kth_gathered_reg_kernel<64><<<blocks, wpb * 32, 0, stream>>>(gathered, R, T, m, kth, tok_thr);
  else
    kth_gathered_mem_kernel<<<blocks, wpb * 32, 0, stream>>>(gathered, R, T, m, kth, tok_thr);
  SAEB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

// ---------------------------------------------------------------------------------------------
// top-activation scan
// ---------------------------------------------------------------------------------------------
constexpr uint32_t HASH_EMPTY = 0xffffffffu;

// one CTA per window of `ctx_len` tokens: max-pool per feature in a shared-memory hash table, then append the
// pooled (value, window) pairs that beat the feature's current n-th best to the feature's bucket.
__global__ void scan_pool_kernel(const float* __restrict__ vals, const long long* __restrict__ idx, long long T, int k,
                                 int ctx_len, float threshold, long long feat_lo, long long feat_hi,
                                 long long window_base, const float* __restrict__ tok_thr,
                                 const float* __restrict__ feat_thr, uint2* __restrict__ bucket,
                                 int* __restrict__ bucket_cnt, int bucket_cap, int slots, int* __restrict__ overflow) {
  extern __shared__ uint32_t hsm[];
  uint32_t* hkey = hsm;
  uint32_t* hval = hsm + slots;
  const uint32_t mask = (uint32_t)slots - 1u;
  for (int i = threadIdx.x; i < slots; i += blockDim.x) {
    hkey[i] = HASH_EMPTY;
    hval[i] = 0u;
  }
  __syncthreads();
  const long long w = blockIdx.x;
  const long long t0 = w * ctx_len;
  const int n_ent = ctx_len * k;
  for (int e = threadIdx.x; e < n_ent; e += blockDim.x) {
    const long long t = t0 + e / k;
    if (t >= T) continue;
    const float v = vals[t * k + (e % k)];
    const long long f = idx[t * k + (e % k)];
    if (!(v > threshold) || f < feat_lo || f >= feat_hi) continue;
    if (tok_thr != nullptr && v < tok_thr[t]) continue;   // not in the token's global top-k
    const uint32_t key = (uint32_t)(f - feat_lo);
    uint32_t s = (key * 2654435761u) & mask;
    while (true) {
      const uint32_t old = atomicCAS(&hkey[s], HASH_EMPTY, key);
      if (old == HASH_EMPTY || old == key) {
        atomicMax(&hval[s], __float_as_uint(v));
        break;
      }
      s = (s + 1) & mask;
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < slots; i += blockDim.x) {
    const uint32_t key = hkey[i];
    if (key == HASH_EMPTY) continue;
    const uint32_t vb = hval[i];
    if (!(__uint_as_float(vb) > feat_thr[key])) continue;
    const int pos = atomicAdd(&bucket_cnt[key], 1);
    if (pos < bucket_cap) bucket[(size_t)key * bucket_cap + pos] = make_uint2(vb, (uint32_t)(window_base + w));
    else if (overflow) atomicExch(overflow, 1);
  }
}

// one warp per feature: merge bucket into the feature's sorted top-n list; order (value desc, window asc)
__global__ void scan_merge_kernel(uint2* __restrict__ bucket, int* __restrict__ bucket_cnt, int bucket_cap, long long F,
                                  int n_top, int sort_n, float base_thr, float* __restrict__ top_vals,
                                  long long* __restrict__ top_win, float* __restrict__ feat_thr) {
  extern __shared__ uint2 ssm[];
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const long long f = (long long)blockIdx.x * (blockDim.x >> 5) + warp;
  if (f >= F) return;
  int cnt = bucket_cnt[f];
  if (cnt == 0) return;
  if (cnt > bucket_cap) cnt = bucket_cap;
  uint2* e = ssm + (size_t)warp * sort_n;
  // sort only as many slots as this feature needs (most features receive a handful of new entries per flush)
  int need = 2;
  while (need < n_top + cnt) need <<= 1;
  if (need < sort_n) sort_n = need;
  for (int i = lane; i < sort_n; i += 32) {
    uint2 ent = make_uint2(0u, 0xffffffffu);
    if (i < n_top) {
      const long long wv = top_win[f * n_top + i];
      if (wv >= 0) ent = make_uint2(__float_as_uint(top_vals[f * n_top + i]), (uint32_t)wv);
    } else if (i - n_top < cnt) {
      ent = bucket[(size_t)f * bucket_cap + (i - n_top)];
    }
    e[i] = ent;
  }
  __syncwarp();
  for (int size = 2; size <= sort_n; size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      for (int p = lane; p < (sort_n >> 1); p += 32) {
        const int lo = ((p / stride) * stride * 2) + (p % stride);
        const int hi = lo + stride;
        const bool desc = ((lo & size) == 0);
        const uint2 a = e[lo], b = e[hi];
        const bool a_first = (a.x > b.x) || (a.x == b.x && a.y < b.y);
        if (a_first != desc) {
          e[lo] = b;
          e[hi] = a;
        }
      }
      __syncwarp();
    }
  }
  for (int i = lane; i < n_top; i += 32) {
    const uint2 ent = e[i];
    const bool ok = ent.x != 0u;
    top_vals[f * n_top + i] = ok ? __uint_as_float(ent.x) : 0.f;
    top_win[f * n_top + i] = ok ? (long long)ent.y : -1ll;
  }
  if (lane == 0) {
    const uint2 last = e[n_top - 1];
    feat_thr[f] = (last.x != 0u) ? __uint_as_float(last.x) : base_thr;
    bucket_cnt[f] = 0;
  }
}

int scan_pool_launch(const float* vals, const long long* idx, long long T, int k, int ctx_len, float threshold,
                     long long feat_lo, long long feat_hi, long long window_base, const float* tok_thr,
                     const float* feat_thr, void* bucket, int* bucket_cnt, int bucket_cap, int* overflow,
                     cudaStream_t stream) {
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
  scan_pool_kernel<<<(unsigned)n_win, 256, smem, stream>>>(vals, idx, T, k, ctx_len, threshold, feat_lo, feat_hi,
                                                          window_base, tok_thr, feat_thr,
                                                          reinterpret_cast<uint2*>(bucket), bucket_cnt, bucket_cap,
                                                          slots, overflow);
  SAEB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int scan_merge_launch(void* bucket, int* bucket_cnt, int bucket_cap, long long F, int n_top, float base_thr,
                      float* top_vals, long long* top_win, float* feat_thr, cudaStream_t stream) {
  SAEB_REQUIRE(F > 0 && n_top >= 1 && n_top <= 1024, "scan_merge: bad arguments");
  int sort_n = 2;
  while (sort_n < n_top + bucket_cap) sort_n <<= 1;
  const int wpb = 8;
  const size_t smem = (size_t)wpb * sort_n * sizeof(uint2);
  SAEB_REQUIRE(smem <= 200 * 1024, "scan_merge: n_top + bucket_cap too large");
  SAEB_CHECK_CUDA(cudaFuncSetAttribute(scan_merge_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  scan_merge_kernel<<<(unsigned)((F + wpb - 1) / wpb), wpb * 32, smem, stream>>>(
      reinterpret_cast<uint2*>(bucket), bucket_cnt, bucket_cap, F, n_top, sort_n, base_thr, top_vals, top_win,
      feat_thr);
  SAEB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace saeb
