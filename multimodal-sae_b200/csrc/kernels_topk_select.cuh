// Device code only (no launch syntax): included by encode_topk.cu for the GPU build and, with SAEB_CPU_EMU defined, by the CPU
// emulation harness under tests/emu, which runs these kernels thread by thread on the host (tests/test_kernel_emu.py).
#pragma once
#include "common.cuh"

namespace saeb {

// ---------------------------------------------------------------------------------------------
// warp-cooperative compaction of one row's candidate list
//
// Any threshold that keeps at least k entries is valid (the merge kernel does the exact selection), so the common path
// takes the threshold from a 256-bin histogram of the value bits over the list's [min, max] range: one shared-memory
// atomic per entry + an 8-bins-per-lane suffix scan instead of a 31-step bit search.  If that would keep too many
// entries (heavy ties), the exact path selects the k-th value and keeps exactly k entries, ties by arrival order
// (= ascending column, since appends and compactions preserve order).
// ---------------------------------------------------------------------------------------------
template <int SLOTS>
__device__ __forceinline__ void compact_row(uint2* buf, int cnt_in, int k, uint32_t lane, int* hist, float& thr_out,
                                            int& cnt_out) {
  constexpr int CAP = 32 * SLOTS;
  const uint32_t full = 0xffffffffu;
  const uint32_t lt_mask = (1u << lane) - 1u;
  uint32_t key[SLOTS], col[SLOTS];
  uint32_t kmax = 0, kmin = 0xffffffffu;
#pragma unroll
  for (int s = 0; s < SLOTS; ++s) {
    const int i = s * 32 + lane;
    if (i < cnt_in) {
      const uint2 e = buf[i];
      key[s] = e.x;
      col[s] = e.y;
      kmax = max(kmax, e.x);
      kmin = min(kmin, e.x);
    } else {
      key[s] = 0;
      col[s] = 0;
    }
  }
  kmax = __reduce_max_sync(full, kmax);
  kmin = __reduce_min_sync(full, kmin);
  // values are strictly positive floats, so their bit patterns order like unsigned integers
  const uint32_t range = kmax - kmin;
  const int shift = (range >> 8) ? (32 - __clz(range) - 8) : 0;
#pragma unroll
  for (int b = 0; b < 8; ++b) hist[lane * 8 + b] = 0;
  __syncwarp();
#pragma unroll
  for (int s = 0; s < SLOTS; ++s)
    if (key[s] != 0) atomicAdd(&hist[(key[s] - kmin) >> shift], 1);
  __syncwarp();
  int c8 = 0;
#pragma unroll
  for (int b = 0; b < 8; ++b) c8 += hist[lane * 8 + b];
  int suffix = c8;   // inclusive suffix sum over lanes >= this one
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int up = __shfl_down_sync(full, suffix, o);
    if (lane + o < 32) suffix += up;
  }
  const uint32_t ok = __ballot_sync(full, suffix >= k);
  const int lstar = 31 - __clz(ok);   // highest lane whose suffix still reaches k (lane 0 always does)
  int acc = __shfl_sync(full, suffix, lstar) - __shfl_sync(full, c8, lstar);
  int bin = lstar * 8;
#pragma unroll
  for (int b = 7; b >= 0; --b) {
    const int h = hist[lstar * 8 + b];
    if (acc < k) {
      acc += h;
      bin = lstar * 8 + b;
    }
  }
  uint32_t thr_key = kmin + ((uint32_t)bin << shift);
  int need_eq = CAP;   // ties at thr_key that may be kept (all of them on the histogram path)
  if (acc > CAP - 64) {
    // exact path: k-th largest key by bit search, keep exactly k
    uint32_t prefix = 0;
    for (int bit = 30; bit >= 0; --bit) {
      const uint32_t trial = prefix | (1u << bit);
      int c = 0;
#pragma unroll
      for (int s = 0; s < SLOTS; ++s) c += (key[s] >= trial) ? 1 : 0;
      c = __reduce_add_sync(full, c);
      if (c >= k) prefix = trial;
    }
    int c_gt = 0;
#pragma unroll
    for (int s = 0; s < SLOTS; ++s) c_gt += (key[s] > prefix) ? 1 : 0;
    c_gt = __reduce_add_sync(full, c_gt);
    thr_key = prefix;
    need_eq = k - c_gt;
  }
  int base = 0, ties = 0;
#pragma unroll
  for (int s = 0; s < SLOTS; ++s) {
    const bool tie = key[s] == thr_key;
    const uint32_t mt = __ballot_sync(full, tie);
    const bool keep = key[s] > thr_key || (tie && ties + __popc(mt & lt_mask) < need_eq);
    const uint32_t m = __ballot_sync(full, keep);
    if (keep) buf[base + __popc(m & lt_mask)] = make_uint2(key[s], col[s]);
    base += __popc(m);
    ties += __popc(mt);
  }
  __syncwarp();
  thr_out = __uint_as_float(thr_key);
  cnt_out = base;
}

// ---------------------------------------------------------------------------------------------
// merge: exact top-k per row over the S candidate lists; canonical order (value desc, index asc)
// ---------------------------------------------------------------------------------------------
__global__ void topk_merge_kernel(const uint2* __restrict__ cand, const int* __restrict__ cand_cnt, int T, int S,
                                  int CAP, int k, int kp2, int N, int max_entries, float* __restrict__ out_vals,
                                  long long* __restrict__ out_idx) {
  extern __shared__ uint2 msm[];
  const int warps_per_block = blockDim.x >> 5;
  const int warp = threadIdx.x >> 5;
  const uint32_t lane = threadIdx.x & 31;
  const uint32_t full = 0xffffffffu;
  const uint32_t lt_mask = (1u << lane) - 1u;
  uint2* ent = msm + (size_t)warp * (max_entries + kp2);
  uint2* sel = ent + max_entries;
  const int row = blockIdx.x * warps_per_block + warp;
  if (row >= T) return;

  // 1. stage valid candidates
  int M = 0;
  for (int s = 0; s < S; ++s) {
    const int c = cand_cnt[(size_t)row * S + s];
    const uint2* src = cand + ((size_t)row * S + s) * CAP;
    for (int i = lane; i < c; i += 32) ent[M + i] = src[i];
    M += c;
  }
  __syncwarp();

  int nsel = 0;
  if (M <= k) {
    for (int i = lane; i < M; i += 32) sel[i] = ent[i];
    nsel = M;
  } else {
    // 2. k-th largest value
    uint32_t prefix = 0;
    for (int bit = 30; bit >= 0; --bit) {
      const uint32_t trial = prefix | (1u << bit);
      int c = 0;
      for (int i = lane; i < M; i += 32) c += (ent[i].x >= trial) ? 1 : 0;
      c = __reduce_add_sync(full, c);
      if (c >= k) prefix = trial;
    }
    int c_gt = 0, c_eq = 0;
    for (int i = lane; i < M; i += 32) {
      c_gt += (ent[i].x > prefix) ? 1 : 0;
      c_eq += (ent[i].x == prefix) ? 1 : 0;
    }
    c_gt = __reduce_add_sync(full, c_gt);
    c_eq = __reduce_add_sync(full, c_eq);
    const int need_eq = k - c_gt;   // >= 1
    // 3. among ties at the k-th value keep the `need_eq` smallest indices
    uint32_t idx_cut = 0xffffffffu;
    if (c_eq > need_eq) {
      uint32_t p2 = 0;   // largest t such that #(idx < t) < need_eq  ->  cut = the need_eq-th smallest index
      for (int bit = 31; bit >= 0; --bit) {
        const uint32_t trial = p2 | (1u << bit);
        int c = 0;
        for (int i = lane; i < M; i += 32) c += (ent[i].x == prefix && ent[i].y < trial) ? 1 : 0;
        c = __reduce_add_sync(full, c);
        if (c < need_eq) p2 = trial;
      }
      idx_cut = p2;
    }
    // 4. compact the selection
    for (int i0 = 0; i0 < M; i0 += 32) {
      const int i = i0 + lane;
      bool keep = false;
      uint2 e = make_uint2(0, 0);
      if (i < M) {
        e = ent[i];
        keep = e.x > prefix || (e.x == prefix && e.y <= idx_cut);
      }
      const uint32_t m = __ballot_sync(full, keep);
      if (keep) sel[nsel + __popc(m & lt_mask)] = e;
      nsel += __popc(m);
    }
  }
  for (int i = nsel + lane; i < kp2; i += 32) sel[i] = make_uint2(0u, 0xffffffffu);   // sorts last
  __syncwarp();

  // 5. bitonic sort, descending by (value bits, then ascending index)
  for (int size = 2; size <= kp2; size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      for (int t = lane; t < (kp2 >> 1); t += 32) {
        const int lo = ((t / stride) * stride * 2) + (t % stride);
        const int hi = lo + stride;
        const bool desc = ((lo & size) == 0);
        const uint2 a = sel[lo], b = sel[hi];
        // "a before b" in canonical order
        const bool a_first = (a.x > b.x) || (a.x == b.x && a.y < b.y);
        if (a_first != desc) {
          sel[lo] = b;
          sel[hi] = a;
        }
      }
      __syncwarp();
    }
  }

  // 6. rows with fewer than k positive pre-activations: pad with zeros on distinct unused indices
  //    (torch.topk would return arbitrary zero-valued entries there; they are dropped by the cache's >1e-5 test
  //    and contribute nothing to the decode).
  if (nsel < k) {
    int filled = nsel;
    for (int base = 0; base < N && filled < k; base += 32) {
      const uint32_t j = base + lane;
      bool free_idx = j < (uint32_t)N;
      for (int i = 0; i < nsel && free_idx; ++i) free_idx = sel[i].y != j;
      const uint32_t m = __ballot_sync(full, free_idx);
      const int pos = filled + __popc(m & lt_mask);
      if (free_idx && pos < k) sel[pos] = make_uint2(0u, j);
      filled += __popc(m);
      __syncwarp();
    }
  }
  __syncwarp();
  for (int i = lane; i < k; i += 32) {
    out_vals[(size_t)row * k + i] = __uint_as_float(sel[i].x);
    out_idx[(size_t)row * k + i] = (long long)sel[i].y;
  }
}

// ---------------------------------------------------------------------------------------------
// Feature-sharded scan, step 1 in ONE kernel: select the row's K2 best candidates out of its S lists and emit the
// m1 largest lower bounds a_j - eps_j and the m1 largest upper bounds a_j + eps_j (exchange 1 of saeb200.dist).
//
// One warp per row, everything in registers (S * CAP <= 32 * SSB_VPL entries, 16 per lane) except a 1 KB staging
// buffer per warp: no passes over shared memory (the tensor cores of the GEMM CTA next door read their operands from
// it at full rate), 128-thread CTAs that run beside a resident GEMM CTA.  Replaces topk_merge_kernel +
// candidate_bounds_kernel (+ a concatenation) in the pipelined scan.  Nothing here is sorted: the selected candidates
// are written in list order (the warp-per-token refinement does not need an order: it takes the smallest kept value
// as a_last), and the consumers of the bound lists (k-th largest over all shards; per shard the SMALLEST bound sent)
// are order-independent too.  Selection rule = topk_merge_kernel's: K2 largest values, ties at the K2-th value by
// smallest column.
// ---------------------------------------------------------------------------------------------
constexpr int SSB_VPL = 16;
constexpr int SSB_THREADS = 128;

// the `want` largest of the warp's 4-per-lane non-negative values -> dst[0..want) (any order, zero padded)
__device__ __forceinline__ void ssb_top_values(const float (&v)[4], int want, float* __restrict__ dst, int lane) {
  const uint32_t full = 0xffffffffu, lt_mask = (1u << lane) - 1u;
  uint32_t key[4];
  int npos = 0;
#pragma unroll
  for (int s = 0; s < 4; ++s) {
    key[s] = v[s] > 0.f ? __float_as_uint(v[s]) : 0u;
    npos += __popc(__ballot_sync(full, key[s] != 0u));
  }
  uint32_t prefix = 0;
  int need_eq = 0x7fffffff;
  if (npos > want) {
    for (int bit = 30; bit >= 0; --bit) {
      const uint32_t trial = prefix | (1u << bit);
      int c = 0;
#pragma unroll
      for (int s = 0; s < 4; ++s) c += (key[s] >= trial) ? 1 : 0;
      c = __reduce_add_sync(full, c);
      if (c >= want) prefix = trial;
    }
    int c_gt = 0;
#pragma unroll
    for (int s = 0; s < 4; ++s) c_gt += (key[s] > prefix) ? 1 : 0;
    c_gt = __reduce_add_sync(full, c_gt);
    need_eq = want - c_gt;
  } else {
    prefix = 1u;   // every positive value is kept (as "greater or equal")
  }
  int n = 0, n_eq = 0;
#pragma unroll
  for (int s = 0; s < 4; ++s) {
    const bool gt = (npos > want) ? key[s] > prefix : key[s] >= prefix;
    const bool eq = (npos > want) && key[s] == prefix;
    const uint32_t m_eq = __ballot_sync(full, eq);
    const bool keep = gt || (eq && n_eq + __popc(m_eq & lt_mask) < need_eq);
    n_eq += __popc(m_eq);
    const uint32_t m = __ballot_sync(full, keep);
    const int pos = n + __popc(m & lt_mask);
    if (keep && pos < want) dst[pos] = __uint_as_float(key[s]);
    n += __popc(m);
  }
  for (int j = (n < want ? n : want) + lane; j < want; j += 32) dst[j] = 0.f;
}

__global__ void __launch_bounds__(SSB_THREADS)
scan_select_bounds_kernel(const uint2* __restrict__ cand, const int* __restrict__ cand_cnt, int T, int S, int CAP,
                          int K2, int m1, const float* __restrict__ wnorm, const float* __restrict__ dnorm,
                          const float* __restrict__ xnorm, const float* __restrict__ xdnorm, float c_eps,
                          long long clamp_feature, float* __restrict__ out_vals, long long* __restrict__ out_idx,
                          float* __restrict__ exch) {
  __shared__ uint2 stage_all[SSB_THREADS / 32][128];
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t full = 0xffffffffu, lt_mask = (1u << lane) - 1u;
  const int row = blockIdx.x * (blockDim.x >> 5) + warp;
  if (row >= T) return;   // per warp: the whole warp leaves together
  uint2* stage = stage_all[warp];

  // 1. the row's candidates, list after list, 16 per lane (slot q = v * 32 + lane holds position q)
  uint32_t key[SSB_VPL], col[SSB_VPL];
  int M = 0;
#pragma unroll
  for (int v = 0; v < SSB_VPL; ++v) {
    key[v] = 0u;
    col[v] = 0xffffffffu;
  }
  for (int s = 0; s < S; ++s) {
    const int c = cand_cnt[(size_t)row * S + s];
    const uint2* src = cand + ((size_t)row * S + s) * CAP;
#pragma unroll
    for (int v = 0; v < SSB_VPL; ++v) {
      const int i = v * 32 + lane - M;   // position inside list s
      if (i >= 0 && i < c) {
        const uint2 e = src[i];
        key[v] = e.x;
        col[v] = e.y;
      }
    }
    M += c;
  }
  // 2. K2-th largest value (ties: smallest columns), exactly like topk_merge_kernel
  uint32_t prefix = 0, idx_cut = 0xffffffffu;
  if (M > K2) {
    for (int bit = 30; bit >= 0; --bit) {
      const uint32_t trial = prefix | (1u << bit);
      int c = 0;
#pragma unroll
      for (int v = 0; v < SSB_VPL; ++v) c += (key[v] >= trial) ? 1 : 0;
      c = __reduce_add_sync(full, c);
      if (c >= K2) prefix = trial;
    }
    int c_gt = 0, c_eq = 0;
#pragma unroll
    for (int v = 0; v < SSB_VPL; ++v) {
      c_gt += (key[v] > prefix) ? 1 : 0;
      c_eq += (key[v] == prefix) ? 1 : 0;
    }
    const int both = __reduce_add_sync(full, c_gt | (c_eq << 16));
    c_gt = both & 0xffff;
    c_eq = both >> 16;
    const int need_eq = K2 - c_gt;
    if (c_eq > need_eq) {
      uint32_t p2 = 0;   // largest t with #(col < t among the ties) < need_eq  ->  the need_eq-th smallest column
      for (int bit = 31; bit >= 0; --bit) {
        const uint32_t trial = p2 | (1u << bit);
        int c = 0;
#pragma unroll
        for (int v = 0; v < SSB_VPL; ++v) c += (key[v] == prefix && col[v] < trial) ? 1 : 0;
        c = __reduce_add_sync(full, c);
        if (c < need_eq) p2 = trial;
      }
      idx_cut = p2;
    }
  } else {
    prefix = 1u;   // keep every candidate (value bits >= 1: positive)
  }
  // 3. compact the selection (list order) through the staging buffer
  int nsel = 0;
#pragma unroll
  for (int v = 0; v < SSB_VPL; ++v) {
    const bool keep = (M > K2) ? (key[v] > prefix || (key[v] == prefix && col[v] <= idx_cut)) : key[v] >= prefix;
    const uint32_t m = __ballot_sync(full, keep);
    const int pos = nsel + __popc(m & lt_mask);
    if (keep && pos < 128) stage[pos] = make_uint2(key[v], col[v]);
    nsel += __popc(m);
  }
  __syncwarp();
  // 4. four candidates per lane: merged list out, bounds
  const float xn = xnorm[row], xdn = xdnorm[row];
  float lb[4], ub[4];
#pragma unroll
  for (int s = 0; s < 4; ++s) {
    const int j = lane + 32 * s;
    float av = 0.f;
    long long fj = 0;
    if (j < nsel) {
      const uint2 e = stage[j];
      av = __uint_as_float(e.x);
      fj = (long long)e.y;
    }
    if (j < K2) {
      out_vals[(size_t)row * K2 + j] = av;
      out_idx[(size_t)row * K2 + j] = fj;
    }
    const float wn = wnorm[fj];
    const float eps = (fj == clamp_feature) ? 0.f : 1.001f * (xn * dnorm[fj] + xdn * wn) + c_eps * xn * wn;
    lb[s] = (av > 0.f) ? fmaxf(av - eps, 0.f) : 0.f;
    ub[s] = (av > 0.f) ? av + eps : 0.f;
  }
  // 5. the m1 largest lower bounds | the m1 largest upper bounds of the row
  float* dst = exch + (size_t)row * (2 * m1);
  ssb_top_values(lb, m1, dst, lane);
  ssb_top_values(ub, m1, dst + m1, lane);
}

}  // namespace saeb
