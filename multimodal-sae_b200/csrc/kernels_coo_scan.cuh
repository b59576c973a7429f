// Device code only (no launch syntax): included by coo_scan.cu for the GPU build and, with SAEB_CPU_EMU defined, by the CPU
// emulation harness under tests/emu, which runs these kernels thread by thread on the host (tests/test_kernel_emu.py).
#pragma once
#include "common.cuh"

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
                                long long* __restrict__ locations, float* __restrict__ activations,
                                const long long* __restrict__ base, long long capacity, int* __restrict__ overflow) {
  // base (optional, device): entries are appended behind *base in an arena of `capacity` entries (saeb_coo_append);
  // whatever does not fit is dropped and reported through *overflow
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
  const int cnt = (int)(offsets[t + 1] - offsets[t]);
  const long long off = offsets[t] + (base != nullptr ? *base : 0);
  const long long r = row_offset + t / seq_len, pos = t % seq_len;
  for (int j = lane; j < cnt; j += 32) {
    if (base != nullptr && off + j >= capacity) {
      if (overflow != nullptr) atomicExch(overflow, 1);
      continue;
    }
    locations[(off + j) * 3 + 0] = r;
    locations[(off + j) * 3 + 1] = pos;
    locations[(off + j) * 3 + 2] = (long long)e[j].x;
    activations[off + j] = __uint_as_float(e[j].y);
  }
}

// *cursor += *count (one thread): advances the arena cursor of saeb_coo_append after the emit kernel has read it
__global__ void coo_advance_kernel(long long* __restrict__ cursor, const long long* __restrict__ count,
                                   long long capacity) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    const long long c = *cursor + *count;
    *cursor = c < capacity ? c : capacity;
  }
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
                                 const float* __restrict__ member, const float* __restrict__ feat_thr,
                                 uint2* __restrict__ bucket,
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
    // not in the token's global top-k?  `member` (optional, same layout as vals): the value that decides membership
    // when it is not the activation itself (value_mode 2 of the sharded refinement: MEMBER_SURE for certain members)
    if (tok_thr != nullptr && (member != nullptr ? member[t * k + (e % k)] : v) < tok_thr[t]) continue;
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

// Same result as scan_pool_kernel with the hash table in GLOBAL scratch instead of shared memory: persistent CTAs of 128
// threads walk the windows (blockIdx.x, blockIdx.x + gridDim.x, ...), each with a private table that cleans itself
// through the list of slots it occupied (like image_pool_kernel), so the launch needs no shared memory and is
// scheduled beside a resident GEMM CTA -- the pipelined scan runs its list update inside the next chunk's GEMM
// launches.  Entries that cannot beat their feature's current n-th best are dropped BEFORE they reach the table
// (the pooled maximum is appended only if it beats that threshold, so the appended set is the same): late in a
// scan almost nothing is inserted.  Scratch: keys [G][slots] (HASH_EMPTY), vals [G][slots] (0), list [G][slots];
// initialised once, left initialised by every call.
__global__ void __launch_bounds__(128)
scan_pool_g_kernel(const float* __restrict__ vals, const long long* __restrict__ idx, long long T, int k, int ctx_len,
                   float threshold, long long feat_lo, long long feat_hi, long long window_base,
                   const float* __restrict__ tok_thr, const float* __restrict__ member,
                   const float* __restrict__ feat_thr, uint2* __restrict__ bucket, int* __restrict__ bucket_cnt,
                   int bucket_cap, int slots, uint32_t* hkeys, uint32_t* hvals, int* hlist, long long n_win,
                   int* __restrict__ overflow) {
  __shared__ int s_n;
  uint32_t* keys = hkeys + (size_t)blockIdx.x * slots;
  uint32_t* hval = hvals + (size_t)blockIdx.x * slots;
  int* list = hlist + (size_t)blockIdx.x * slots;
  const uint32_t mask = (uint32_t)slots - 1u;
  const int n_ent = ctx_len * k;
  for (long long w = blockIdx.x; w < n_win; w += gridDim.x) {
    if (threadIdx.x == 0) s_n = 0;
    __syncthreads();
    const long long t0 = w * ctx_len;
    for (int e = threadIdx.x; e < n_ent; e += blockDim.x) {
      const long long t = t0 + e / k;
      if (t >= T) continue;
      const float v = vals[t * k + (e % k)];
      const long long f = idx[t * k + (e % k)];
      if (!(v > threshold) || f < feat_lo || f >= feat_hi) continue;
      if (tok_thr != nullptr && (member != nullptr ? member[t * k + (e % k)] : v) < tok_thr[t]) continue;
      const uint32_t key = (uint32_t)(f - feat_lo);
      if (!(v > feat_thr[key])) continue;
      uint32_t s = (key * 2654435761u) & mask;
      while (true) {
        const uint32_t old = atomicCAS(&keys[s], HASH_EMPTY, key);
        if (old == HASH_EMPTY) list[atomicAdd(&s_n, 1)] = (int)s;
        if (old == HASH_EMPTY || old == key) {
          atomicMax(&hval[s], __float_as_uint(v));
          break;
        }
        s = (s + 1) & mask;
      }
    }
    __threadfence();
    __syncthreads();
    const int n = s_n;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
      const int s = list[i];
      // the table was written by atomics (L2): read it around L1
      const uint32_t key = *reinterpret_cast<volatile uint32_t*>(&keys[s]);
      const uint32_t vb = *reinterpret_cast<volatile uint32_t*>(&hval[s]);
      keys[s] = HASH_EMPTY;
      hval[s] = 0u;
      const int pos = atomicAdd(&bucket_cnt[key], 1);
      if (pos < bucket_cap) bucket[(size_t)key * bucket_cap + pos] = make_uint2(vb, (uint32_t)(window_base + w));
      else if (overflow) atomicExch(overflow, 1);
    }
    __threadfence();
    __syncthreads();
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

// ---------------------------------------------------------------------------------------------
// cache reader on the device: pooled score of every (feature, window) of a split file
// ---------------------------------------------------------------------------------------------
// Input: the cached triples of a split file grouped by feature (stable, so a feature's entries keep the file's
// (row, pos) order) with a window key per entry (text constructor: row * n_win + pos / ctx_len, reference
// features/constructors.py:11-47; image constructor: the row, for positions below n_base, :109-114; -1 = entry outside
// every window).  Entries of one (feature, window) are consecutive.  The thread that sits on the FIRST entry of a run
// reduces the run in file order -- max for the text windows (max_pool1d), a sequential fp32 sum divided by `divisor` for
// the image mean (the same order as the reference's CPU index_add_ / avg_pool) -- and flags itself as the run's head.
__global__ void __launch_bounds__(256)
coo_window_scores_kernel(const long long* __restrict__ feat, const long long* __restrict__ key,
                         const float* __restrict__ act, long long nnz, int mode, float divisor,
                         float* __restrict__ score, int* __restrict__ head) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nnz) return;
  const long long f = feat[i], k = key[i];
  const bool is_head = k >= 0 && (i == 0 || feat[i - 1] != f || key[i - 1] != k);
  head[i] = is_head ? 1 : 0;
  if (!is_head) {
    score[i] = 0.f;
    return;
  }
  float acc = act[i];
  for (long long j = i + 1; j < nnz && feat[j] == f && key[j] == k; ++j) acc = (mode == 0) ? fmaxf(acc, act[j]) : acc + act[j];
  score[i] = (mode == 0) ? acc : acc / divisor;
}

// ---------------------------------------------------------------------------------------------
// image scan: per-feature MEAN of the TopK-masked activations over the first `n_base` positions of every image row
// (reference pool_max_activations_windows_image, features/constructors.py:109-114: avg_pool1d over the 576 base image
// tokens), feeding the same per-feature buckets / top lists as the window scan (image id in place of the window id).
//
// One persistent CTA walks images blockIdx.x, blockIdx.x + gridDim.x, ...  Its hash table (feature -> sum) lives in
// global scratch: an image can touch up to n_base * k distinct features, far more than shared memory holds.  Sums are
// kept in 32.32 fixed point, so they do not depend on the order in which the threads add (deterministic lists); the
// slots an image occupied are remembered in a list and restored to "empty" when its scores have been emitted, so the
// scratch only has to be initialised once (keys = HASH_EMPTY, sums = 0).
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
image_pool_kernel(const float* __restrict__ vals, const long long* __restrict__ idx, long long n_images,
                  long long tokens_per_image, int k, int n_base, float threshold, long long feat_lo, long long feat_hi,
                  long long image_base, const float* __restrict__ tok_thr, const float* __restrict__ feat_thr,
                  uint32_t* __restrict__ hkeys, unsigned long long* __restrict__ hsums, int* __restrict__ hlist,
                  int slots, uint2* __restrict__ bucket, int* __restrict__ bucket_cnt, int bucket_cap,
                  int* __restrict__ overflow) {
  __shared__ int s_n;
  uint32_t* keys = hkeys + (size_t)blockIdx.x * slots;
  unsigned long long* sums = hsums + (size_t)blockIdx.x * slots;
  int* list = hlist + (size_t)blockIdx.x * slots;
  const uint32_t mask = (uint32_t)slots - 1u;
  const long long n_ent = (long long)n_base * k;
  for (long long img = blockIdx.x; img < n_images; img += gridDim.x) {
    if (threadIdx.x == 0) s_n = 0;
    __syncthreads();
    for (long long e = threadIdx.x; e < n_ent; e += blockDim.x) {
      const long long t = img * tokens_per_image + e / k;
      const float v = vals[t * k + (e % k)];
      const long long f = idx[t * k + (e % k)];
      if (!(v > threshold) || f < feat_lo || f >= feat_hi) continue;
      if (tok_thr != nullptr && v < tok_thr[t]) continue;   // not in the token's global top-k
      const uint32_t key = (uint32_t)(f - feat_lo);
      const unsigned long long q = (unsigned long long)((double)v * 4294967296.0 + 0.5);
      uint32_t s = (key * 2654435761u) & mask;
      while (true) {
        const uint32_t old = atomicCAS(&keys[s], HASH_EMPTY, key);
        if (old == HASH_EMPTY) list[atomicAdd(&s_n, 1)] = (int)s;
        if (old == HASH_EMPTY || old == key) {
          atomicAdd(&sums[s], q);
          break;
        }
        s = (s + 1) & mask;
      }
    }
    __syncthreads();
    const int n = s_n;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
      const int s = list[i];
      const uint32_t key = keys[s];
      const unsigned long long sum = sums[s];
      keys[s] = HASH_EMPTY;
      sums[s] = 0ull;
      const float score = (float)((double)sum * (1.0 / 4294967296.0) / (double)n_base);
      if (!(score > feat_thr[key])) continue;
      const int pos = atomicAdd(&bucket_cnt[key], 1);
      if (pos < bucket_cap) bucket[(size_t)key * bucket_cap + pos] = make_uint2(__float_as_uint(score), (uint32_t)(image_base + img));
      else if (overflow) atomicExch(overflow, 1);
    }
    __syncthreads();
  }
}

}  // namespace saeb
