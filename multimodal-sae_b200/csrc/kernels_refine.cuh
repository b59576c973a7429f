// Device code only (no launch syntax): included by refine.cu for the GPU build and, with SAEB_CPU_EMU defined, by the CPU
// emulation harness under tests/emu, which runs these kernels thread by thread on the host (tests/test_kernel_emu.py).
#pragma once
#include <type_traits>

#include "common.cuh"

namespace saeb {

constexpr int RF_THREADS = 256;
constexpr int RF_MAX_FLAG = 64;   // flagged rows handled by the wide (all-SM) fallback; further rows take the
                                  // one-block-per-row overflow path, so any number of flagged rows stays exact

// LO = false: candidates are re-evaluated exactly against the fp32 row W[f] (the parity default).
// LO = true ("fp16 hi + lo", packed mode 4): the approximate value a_j = x . W_hi[f] + bias already comes out of the
// tensor cores; only the missing part x . W_lo[f] is added, W_lo = fp16 plane of the residual (W_hi + W_lo = W to
// 2^-22) -- half the gather bytes.  Exact for bf16 / fp16 activations (they reach the tensor cores unrounded); the
// corrected values carry the tensor cores' fp32 accumulation noise (~1e-6 relative, like any fp32 GEMM) instead of
// the 3e-7 of the exact route.
template <typename XT, bool LO>
__device__ __forceinline__ void
refine_body(const XT* __restrict__ x, long long ld_x, const float* __restrict__ W, long long d, long long N,
            const float* __restrict__ bias, const float* __restrict__ wnorm, const float* __restrict__ dnorm,
            const float* __restrict__ trailer, const float* __restrict__ xnorm, const float* __restrict__ xdnorm,
            float c_eps, const float* __restrict__ cand_vals,
            const long long* __restrict__ cand_idx, int K2, int k, long long clamp_feature, float clamp_value,
            float* __restrict__ out_vals, long long* __restrict__ out_idx, int* __restrict__ status,
            int* __restrict__ flag_rows, const float* __restrict__ ext_lower, const __half* __restrict__ Wlo,
            long long ld_w, long long T, unsigned long long* __restrict__ stats, int value_mode,
            const float* __restrict__ ext_upper, const float* __restrict__ feat_thr, float* __restrict__ out_member) {
  extern __shared__ float rsm[];
  float* xs = rsm;                                   // [d4] activations of this row as fp32
  const int d4 = LO ? (int)((d + 7) & ~7ll) : (int)((d + 3) & ~3ll);   // LO: padded like the packed weight rows
  float* a = xs + d4;                                // [K2] approximate values
  float* lb = a + K2;                                // [K2]
  float* ub = lb + K2;                               // [K2]
  float* ex = ub + K2;                               // [K2] exact values (or -1)
  int* f = reinterpret_cast<int*>(ex + K2);          // [K2] feature ids
  int* st = f + K2;                                  // [K2] boundary-only mode: 1 = certainly in the TopK, 2 = evaluate
  __shared__ float s_L, s_U;
  __shared__ int s_cnt[4];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nthr = blockDim.x;
  // one token per CTA when gridDim.x >= T; a smaller (persistent) grid walks the tokens with stride gridDim.x so that
  // a fixed number of CTAs per SM can ride beside the persistent GEMM grid
  for (long long t = blockIdx.x; t < T; t += gridDim.x) {
  for (int i = tid; i < d4; i += nthr) xs[i] = (i < d) ? (float)x[t * ld_x + i] : 0.f;
  const float xn = xnorm[t], xdn = xdnorm[t];
  const float wmax = trailer[1], dmax = trailer[3];
  for (int j = tid; j < K2; j += nthr) {
    const float av = cand_vals[t * K2 + j];
    const int fj = (int)cand_idx[t * K2 + j];
    const bool valid = av > 0.f;
    const float wn = wnorm[fj];
    const float eps = (fj == clamp_feature) ? 0.f : 1.001f * (xn * dnorm[fj] + xdn * wn) + c_eps * xn * wn;
    a[j] = av;
    f[j] = fj;
    lb[j] = valid ? av - eps : -INFINITY;
    ub[j] = valid ? av + eps : -INFINITY;
    ex[j] = -1.f;
  }
  if (tid == 0) {
    s_L = 0.f;
    s_cnt[0] = 0;
    s_cnt[1] = 0;
    s_cnt[2] = 0;
    s_cnt[3] = 0;
    s_U = 0.f;
  }
  __syncthreads();
  // L = k-th largest lower bound (0 if fewer than k positive candidates exist)
  int my_valid = 0;
  for (int j = tid; j < K2; j += nthr) {
    if (a[j] > 0.f) {
      ++my_valid;
      int rank = 0;
      const float l = lb[j];
      for (int i = 0; i < K2; ++i) rank += (lb[i] > l || (lb[i] == l && i < j)) ? 1 : 0;
      if (rank == k - 1) s_L = fmaxf(l, 0.f);
    }
  }
  if (my_valid) atomicAdd(&s_cnt[0], my_valid);
  __syncthreads();
  const int nv = s_cnt[0];
  float L = (nv >= k) ? s_L : 0.f;
  // feature-sharded use: a lower bound of the GLOBAL k-th value (k-th largest lower bound over all shards)
  if (ext_lower != nullptr) L = fmaxf(L, ext_lower[t]);
  // exact value of candidate j (all lanes of the calling warp): fp32 dot product with the fp32 W_enc row + bias
  const bool vec = (d & 3) == 0 && ((reinterpret_cast<uintptr_t>(W) & 15) == 0);
  auto evaluate = [&](int j) {
    const int fj = f[j];
    float val;
    if (fj == clamp_feature) {
      val = clamp_value;
    } else if constexpr (LO) {
      // residual correction: 8 halves per 16-byte load; rows are zero padded to d_pad (multiple of 8), xs to d4
      const uint4* w8 = reinterpret_cast<const uint4*>(Wlo + (long long)fj * ld_w);
      const float4* x4 = reinterpret_cast<const float4*>(xs);
      const int n8 = d4 >> 3;
      float acc0 = 0.f, acc1 = 0.f, acc2 = 0.f, acc3 = 0.f;
      auto fma8 = [&](uint4 wv, int c8) {
        const float2 w0 = __half22float2(*reinterpret_cast<const __half2*>(&wv.x));
        const float2 w1 = __half22float2(*reinterpret_cast<const __half2*>(&wv.y));
        const float2 w2 = __half22float2(*reinterpret_cast<const __half2*>(&wv.z));
        const float2 w3 = __half22float2(*reinterpret_cast<const __half2*>(&wv.w));
        const float4 xa = x4[2 * c8], xb = x4[2 * c8 + 1];
        acc0 = fmaf(w0.x, xa.x, acc0);
        acc1 = fmaf(w0.y, xa.y, acc1);
        acc2 = fmaf(w1.x, xa.z, acc2);
        acc3 = fmaf(w1.y, xa.w, acc3);
        acc0 = fmaf(w2.x, xb.x, acc0);
        acc1 = fmaf(w2.y, xb.y, acc1);
        acc2 = fmaf(w3.x, xb.z, acc2);
        acc3 = fmaf(w3.y, xb.w, acc3);
      };
      int c = lane;
      for (; c + 7 * 32 < n8; c += 8 * 32) {
        uint4 wv[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) wv[u] = ldg_nc_u4(w8 + c + u * 32);
#pragma unroll
        for (int u = 0; u < 8; ++u) fma8(wv[u], c + u * 32);
      }
      for (; c < n8; c += 32) fma8(ldg_nc_u4(w8 + c), c);
      float acc = (acc0 + acc1) + (acc2 + acc3);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
      // a_j = x . W_hi + folded bias (tensor cores); W_lo carries the residual times 2^11 in the hi plane's scale
      val = fmaf(acc, trailer[0] * (1.0f / 2048.0f), a[j]);
    } else {
      const float* wr = W + (long long)fj * d;
      float acc0 = 0.f, acc1 = 0.f, acc2 = 0.f, acc3 = 0.f;
      if (vec) {
        const float4* w4 = reinterpret_cast<const float4*>(wr);
        const float4* x4 = reinterpret_cast<const float4*>(xs);
        const int n4 = (int)(d >> 2);
        int c = lane;
        for (; c + 7 * 32 < n4; c += 8 * 32) {
          float4 wv[8];
#pragma unroll
          for (int u = 0; u < 8; ++u) wv[u] = ldg_nc_f4(w4 + c + u * 32);
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            const float4 xv = x4[c + u * 32];
            acc0 = fmaf(wv[u].x, xv.x, acc0);
            acc1 = fmaf(wv[u].y, xv.y, acc1);
            acc2 = fmaf(wv[u].z, xv.z, acc2);
            acc3 = fmaf(wv[u].w, xv.w, acc3);
          }
        }
        for (; c < n4; c += 32) {
          const float4 wv = ldg_nc_f4(w4 + c);
          const float4 xv = x4[c];
          acc0 = fmaf(wv.x, xv.x, acc0);
          acc1 = fmaf(wv.y, xv.y, acc1);
          acc2 = fmaf(wv.z, xv.z, acc2);
          acc3 = fmaf(wv.w, xv.w, acc3);
        }
      } else {
        for (long long i = lane; i < d; i += 32) acc0 = fmaf(wr[i], xs[i], acc0);
      }
      float acc = (acc0 + acc1) + (acc2 + acc3);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
      val = acc + bias[fj];
    }
    if (lane == 0) {
      ex[j] = (val > 0.f) ? val : -1.f;
      if (stats != nullptr) atomicAdd(&s_cnt[2], 1);   // diagnostics: rows gathered (saeb_debug_stats slot 7)
    }
  };
  const int nwarps = nthr >> 5;
  // value_mode 1 ("boundary only", unsharded calls): exact values are only needed where they decide the index SET.
  //   U = (k+1)-th largest upper bound: at most k candidates lie above it, so a candidate with a_j - eps_j > U is in
  //   the true TopK whatever the exact values are (non-candidates lie below L <= U unless the row is flagged);
  //   candidates with a_j + eps_j < L are out; only the ones in between are gathered and re-evaluated, and the best
  //   (k - #certain) of them complete the set.  Certain members keep the tensor-core value a_j (|a_j - exact| <= eps_j
  //   rigorously; a member whose bound is looser than 2^-7 of its value is re-evaluated as well).
  // value_mode 2 ("scan"): the caller only wants the members of the TopK that can still enter their feature's top-n
  //   list (upper bound >= feat_thr[feature], the feature's current n-th best; lists only ever rise, so a stale
  //   threshold is merely conservative).  Membership is decided as in mode 1; a certain member that cannot enter its
  //   list is not gathered at all (reported with value 0), one that can is re-evaluated EXACTLY, and so is every
  //   boundary candidate: every value that reaches a list is an exact fp32 value.
  //   Feature-sharded form (ext_upper given = an upper bound of the token's GLOBAL (k+1)-th largest upper bound):
  //   the shard cannot finish the membership decision alone, so it also writes `out_member`: MEMBER_SURE for certain
  //   members, the exact value for boundary candidates -- the k-th largest of all shards' member values is then the
  //   exact value of the weakest member among the boundary candidates (or MEMBER_SURE if there is none), and an entry
  //   belongs to the token's TopK iff its member value reaches it.
  constexpr float MEMBER_SURE = 3.0e38f;
  const bool scan_mode = value_mode == 2;
  const bool sharded_scan = scan_mode && ext_upper != nullptr;
  const bool boundary_only = (value_mode == 1 && ext_lower == nullptr) || scan_mode;
  if (boundary_only) {
    if (!sharded_scan) {
      for (int j = tid; j < K2; j += nthr) {
        if (a[j] > 0.f) {
          int rank = 0;
          const float u = ub[j];
          for (int i = 0; i < K2; ++i) rank += (ub[i] > u || (ub[i] == u && i < j)) ? 1 : 0;
          if (rank == k) s_U = u;
        }
      }
      __syncthreads();
    }
    const float U = sharded_scan ? ext_upper[t] : ((nv > k) ? s_U : 0.f);
    int my_in = 0;
    for (int j = tid; j < K2; j += nthr) {
      int c = 0;
      if (a[j] > 0.f) {
        const float l = lb[j], eps = 0.5f * (ub[j] - l);
        if (l > U && l > 0.f && (scan_mode || eps <= a[j] * 0.0078125f)) {
          ++my_in;
          if (!scan_mode) {
            c = 1;             // certain member, keeps the tensor-core value
            ex[j] = a[j];
          } else if (feat_thr == nullptr || ub[j] >= feat_thr[f[j]]) {
            c = 3;             // certain member that can enter its feature's list: exact value wanted
          } else {
            c = 1;             // certain member that cannot: not gathered
            ex[j] = 0.f;
          }
        } else if (ub[j] >= L) {
          c = 2;               // membership undecided: exact value needed
        }
      }
      st[j] = c;
    }
    if (my_in) atomicAdd(&s_cnt[3], my_in);
    bool flagged0 = false;   // thread 0: a row is flagged at most once (flag_rows holds one slot per row of the call)
    if (tid == 0 && nv == K2) {   // list possibly too short?  (same test as below)
      const float a_last = a[K2 - 1];
      if (a_last + 1.001f * (xn * dmax + xdn * wmax) + c_eps * xn * wmax >= L) {
        const int slot = atomicAdd(&status[0], 1);
        flag_rows[slot] = (int)t;
        flagged0 = true;
      }
    }
    __syncthreads();
    for (int j = warp; j < K2; j += nwarps)
      if (st[j] >= 2) evaluate(j);
    __syncthreads();
    if (!sharded_scan) {
      // keep the best (k - #certain) evaluated boundary candidates, drop the rest: afterwards exactly the members of
      // the TopK (that are wanted) are positive in ex[] and the common final pass only has to order them
      const int need = k - s_cnt[3];
      int* drop = reinterpret_cast<int*>(lb);   // the bounds are not needed any more
      for (int j = tid; j < K2; j += nthr) {
        int dr = 0;
        if (st[j] == 2 && ex[j] > 0.f) {
          const float v = ex[j];
          const int fj = f[j];
          int rank = 0;
          for (int i = 0; i < K2; ++i) rank += (st[i] == 2 && (ex[i] > v || (ex[i] == v && f[i] < fj))) ? 1 : 0;
          dr = rank >= need;
        }
        drop[j] = dr;
      }
      __syncthreads();
      for (int j = tid; j < K2; j += nthr)
        if (drop[j]) ex[j] = -1.f;
      __syncthreads();
    } else {
      // sharded: hand every certain member and every positive boundary candidate to the caller, ordered by member value
      float* mem = lb;   // the bounds are not needed any more
      for (int j = tid; j < K2; j += nthr) {
        float m = -1.f;
        if (st[j] == 1 || st[j] == 3) m = MEMBER_SURE;
        else if (st[j] == 2 && ex[j] > 0.f) m = ex[j];
        mem[j] = m;
      }
      __syncthreads();
      int my_out = 0;
      for (int j = tid; j < K2; j += nthr) {
        const float m = mem[j];
        if (m > 0.f) {
          ++my_out;
          const int fj = f[j];
          int rank = 0;
          for (int i = 0; i < K2; ++i) rank += (mem[i] > m || (mem[i] == m && f[i] < fj)) ? 1 : 0;
          if (rank < k) {
            out_vals[t * k + rank] = fmaxf(ex[j], 0.f);
            out_member[t * k + rank] = m;
            out_idx[t * k + rank] = fj;
          }
        }
      }
      if (my_out) atomicAdd(&s_cnt[1], my_out);
      __syncthreads();
      const int nout = s_cnt[1];
      if (nout > k && tid == 0 && !flagged0) {   // more potential members than output slots: exact dense fallback
        const int slot = atomicAdd(&status[0], 1);
        flag_rows[slot] = (int)t;
      }
      for (int j = (nout < k ? nout : k) + tid; j < k; j += nthr) {
        out_vals[t * k + j] = 0.f;
        out_member[t * k + j] = 0.f;
        out_idx[t * k + j] = 0;
      }
      __syncthreads();
      if (stats != nullptr && tid == 0) atomicAdd(stats + 7, (unsigned long long)s_cnt[2]);
      continue;   // next token of the persistent loop
    }
  } else {
  // stage A: the k best candidates by approximate value (the merged list is sorted by a, descending)
  for (int j = warp; j < k; j += nwarps)
    if (ub[j] >= L && a[j] > 0.f) evaluate(j);
  __syncthreads();
  // k exact values are now known: their smallest is a far tighter lower bound of the k-th value than L (which sits a
  // full eps below it), so fewer of the remaining candidates can still reach the TopK
  if (warp == 0) {
    float mn = INFINITY;
    for (int j = lane; j < k; j += 32) mn = fminf(mn, ex[j]);   // -1 = not evaluated or not positive
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, o));
    if (lane == 0) s_L = (mn > L) ? mn : L;
  }
  __syncthreads();
  L = s_L;
  // list possibly too short?  (only when the list is full: otherwise every positive latent is already in it)
  if (tid == 0 && nv == K2) {
    const float a_last = a[K2 - 1];
    if (a_last + 1.001f * (xn * dmax + xdn * wmax) + c_eps * xn * wmax >= L) {
      const int slot = atomicAdd(&status[0], 1);
      flag_rows[slot] = (int)t;   // capacity = number of rows of the call
    }
  }
  // stage B: every remaining candidate that can still be in the TopK
  for (int j = k + warp; j < K2; j += nwarps)
    if (ub[j] >= L && a[j] > 0.f) evaluate(j);
  __syncthreads();
  }
  // final TopK over the exact values: rank by (value desc, feature id asc)
  int my_pos = 0;
  for (int j = tid; j < K2; j += nthr) {
    const float v = ex[j];
    if (v > 0.f) {
      ++my_pos;
      int rank = 0;
      const int fj = f[j];
      for (int i = 0; i < K2; ++i) rank += (ex[i] > v || (ex[i] == v && f[i] < fj)) ? 1 : 0;
      if (rank < k) {
        out_vals[t * k + rank] = v;
        out_idx[t * k + rank] = fj;
      }
    }
  }
  if (my_pos) atomicAdd(&s_cnt[1], my_pos);
  __syncthreads();
  const int npos = s_cnt[1];
  if (npos < k && ext_lower != nullptr) {
    // feature-sharded call: the tail only has to read as "nothing here" (value 0); ids need not be distinct
    for (int j = npos + tid; j < k; j += nthr) {
      out_vals[t * k + j] = 0.f;
      out_idx[t * k + j] = 0;
    }
  } else if (npos < k && warp == 0) {
    // fewer than k positive latents: pad with zeros on the smallest unused feature ids (as topk_merge_kernel does)
    int filled = npos;
    const uint32_t lt_mask = (1u << lane) - 1u;
    for (long long base = 0; base < N && filled < k; base += 32) {
      const long long jf = base + lane;
      bool free_idx = jf < N;
      for (int i = 0; i < K2 && free_idx; ++i) free_idx = !(ex[i] > 0.f && f[i] == (int)jf);
      const uint32_t m = __ballot_sync(0xffffffffu, free_idx);
      const int pos = filled + __popc(m & lt_mask);
      if (free_idx && pos < k) {
        out_vals[t * k + pos] = 0.f;
        out_idx[t * k + pos] = jf;
      }
      filled += __popc(m);
    }
  }
  __syncthreads();   // the shared row / candidate buffers are reused by the next token
  if (stats != nullptr && tid == 0) atomicAdd(stats + 7, (unsigned long long)s_cnt[2]);
  }
}

template <typename XT>
__global__ void __launch_bounds__(RF_THREADS)
refine_kernel(const XT* __restrict__ x, long long ld_x, const float* __restrict__ W, long long d, long long N,
              const float* __restrict__ bias, const float* __restrict__ wnorm, const float* __restrict__ dnorm,
              const float* __restrict__ trailer, const float* __restrict__ xnorm, const float* __restrict__ xdnorm,
              float c_eps, const float* __restrict__ cand_vals,
              const long long* __restrict__ cand_idx, int K2, int k, long long clamp_feature, float clamp_value,
              float* __restrict__ out_vals, long long* __restrict__ out_idx, int* __restrict__ status,
              int* __restrict__ flag_rows, const float* __restrict__ ext_lower, long long T,
              unsigned long long* __restrict__ stats, int value_mode, const float* __restrict__ ext_upper,
              const float* __restrict__ feat_thr, float* __restrict__ out_member) {
  refine_body<XT, false>(x, ld_x, W, d, N, bias, wnorm, dnorm, trailer, xnorm, xdnorm, c_eps, cand_vals, cand_idx, K2, k,
                         clamp_feature, clamp_value, out_vals, out_idx, status, flag_rows, ext_lower, nullptr, 0, T,
                         stats, value_mode, ext_upper, feat_thr, out_member);
}

template <typename XT>
__global__ void __launch_bounds__(RF_THREADS)
refine_lo_kernel(const XT* __restrict__ x, long long ld_x, const __half* __restrict__ Wlo, long long ld_w, long long d,
                 long long N, const float* __restrict__ bias, const float* __restrict__ wnorm,
                 const float* __restrict__ dnorm, const float* __restrict__ trailer, const float* __restrict__ xnorm,
                 const float* __restrict__ xdnorm, float c_eps, const float* __restrict__ cand_vals,
                 const long long* __restrict__ cand_idx, int K2, int k, long long clamp_feature, float clamp_value,
                 float* __restrict__ out_vals, long long* __restrict__ out_idx, int* __restrict__ status,
                 int* __restrict__ flag_rows, const float* __restrict__ ext_lower, long long T,
                 unsigned long long* __restrict__ stats, int value_mode, const float* __restrict__ ext_upper,
                 const float* __restrict__ feat_thr, float* __restrict__ out_member) {
  refine_body<XT, true>(x, ld_x, nullptr, d, N, bias, wnorm, dnorm, trailer, xnorm, xdnorm, c_eps, cand_vals, cand_idx,
                        K2, k, clamp_feature, clamp_value, out_vals, out_idx, status, flag_rows, ext_lower, Wlo, ld_w, T,
                        stats, value_mode, ext_upper, feat_thr, out_member);
}

// ---------------------------------------------------------------------------------------------
// Warp-per-token form of the refinement for the FEATURE-SHARDED SCAN (value_mode 2 with ext_lower / ext_upper).
//
// A shard of an R-way sharded scan evaluates only a handful of candidates per token exactly (the few members of the
// token's global TopK it holds that can still enter their feature's list, plus the rare boundary candidate), so a
// 256-thread CTA per token with the activation row staged in shared memory spends its time on set-up and barriers.
// Here one WARP owns a token: the K2 <= 128 candidates live in registers (4 per lane), bounds / classification use
// warp collectives only, the activation row is read straight from global memory (L2) during the dot product, and the
// kernel needs no shared memory at all -- 128-thread CTAs that are scheduled BESIDE a resident GEMM CTA, which is
// what the pipelined scan needs (the per-chunk chain runs inside the next chunk's GEMM launches).
// Same classification rules, same bound formula and the same summation order of the exact dot product as
// refine_body (values are bit-identical); the k output slots of a row are filled in candidate order instead of
// (member value, id) order -- both consumers (the member-value exchange + kth, scan_pool) are order-independent.
// ---------------------------------------------------------------------------------------------
constexpr int RSW_THREADS = 128;
constexpr int RSW_SLOTS = 4;   // candidates per lane

template <typename XT>
__device__ __forceinline__ float4 rsw_load_x4(const XT* __restrict__ xr, int c) {
#if defined(SAEB_CPU_EMU)
  return make_float4((float)xr[4 * c], (float)xr[4 * c + 1], (float)xr[4 * c + 2], (float)xr[4 * c + 3]);
#else
  if constexpr (sizeof(XT) == 4) {
    return __ldg(reinterpret_cast<const float4*>(xr) + c);
  } else {
    const uint2 u = __ldg(reinterpret_cast<const uint2*>(xr) + c);
    if constexpr (std::is_same<XT, __nv_bfloat16>::value) {
      return make_float4(__uint_as_float(u.x << 16), __uint_as_float(u.x & 0xffff0000u), __uint_as_float(u.y << 16),
                         __uint_as_float(u.y & 0xffff0000u));
    } else {
      const float2 lo = __half22float2(*reinterpret_cast<const __half2*>(&u.x));
      const float2 hi = __half22float2(*reinterpret_cast<const __half2*>(&u.y));
      return make_float4(lo.x, lo.y, hi.x, hi.y);
    }
  }
#endif
}

template <typename XT>
__global__ void __launch_bounds__(RSW_THREADS)
refine_scan_warp_kernel(const XT* __restrict__ x, long long ld_x, const float* __restrict__ W, long long d, long long N,
                        const float* __restrict__ bias, const float* __restrict__ wnorm,
                        const float* __restrict__ dnorm, const float* __restrict__ trailer,
                        const float* __restrict__ xnorm, const float* __restrict__ xdnorm, float c_eps,
                        const float* __restrict__ cand_vals, const long long* __restrict__ cand_idx, int K2, int k,
                        long long clamp_feature, float clamp_value, float* __restrict__ out_vals,
                        long long* __restrict__ out_idx, int* __restrict__ status, int* __restrict__ flag_rows,
                        const float* __restrict__ ext_lower, long long T, unsigned long long* __restrict__ stats,
                        const float* __restrict__ ext_upper, const float* __restrict__ feat_thr,
                        float* __restrict__ out_member, int vec_ok) {
  constexpr float MEMBER_SURE = 3.0e38f;
  const uint32_t full = 0xffffffffu;
  const int lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
  const uint32_t lt_mask = (1u << lane) - 1u;
  const float wmax = trailer[1], dmax = trailer[3];
  const bool vec = vec_ok != 0;
  for (long long t = (long long)blockIdx.x * wpb + (threadIdx.x >> 5); t < T; t += (long long)gridDim.x * wpb) {
    const float xn = xnorm[t], xdn = xdnorm[t];
    const XT* xr = x + t * ld_x;
    float a[RSW_SLOTS], lb[RSW_SLOTS], ub[RSW_SLOTS], ex[RSW_SLOTS];
    int f[RSW_SLOTS], st[RSW_SLOTS];
    int nv = 0;
#pragma unroll
    for (int s = 0; s < RSW_SLOTS; ++s) {
      const int j = lane + 32 * s;
      float av = 0.f;
      int fj = 0;
      if (j < K2) {
        av = cand_vals[t * K2 + j];
        fj = (int)cand_idx[t * K2 + j];
      }
      const bool valid = av > 0.f;
      const float wn = wnorm[fj];
      const float eps = (fj == clamp_feature) ? 0.f : 1.001f * (xn * dnorm[fj] + xdn * wn) + c_eps * xn * wn;
      a[s] = av;
      f[s] = fj;
      lb[s] = valid ? av - eps : -INFINITY;
      ub[s] = valid ? av + eps : -INFINITY;
      ex[s] = -1.f;
      nv += __popc(__ballot_sync(full, valid));
    }
    // L = max(k-th largest local lower bound (0 if fewer than k positive candidates), external lower bound).  The
    // external bound (k-th largest lower bound over ALL shards) is the tight one; the local search (31 dependent
    // warp-wide steps) only runs where it gave nothing.  A smaller L is still a valid lower bound: at worst a few more
    // candidates are re-evaluated exactly.
    float L = 0.f;
    if (nv >= k && !(ext_lower[t] > 0.f)) {
      uint32_t key[RSW_SLOTS];
#pragma unroll
      for (int s = 0; s < RSW_SLOTS; ++s) key[s] = lb[s] > 0.f ? __float_as_uint(lb[s]) : 0u;
      uint32_t prefix = 0;
      for (int bit = 30; bit >= 0; --bit) {
        const uint32_t trial = prefix | (1u << bit);
        int c = 0;
#pragma unroll
        for (int s = 0; s < RSW_SLOTS; ++s) c += (key[s] >= trial) ? 1 : 0;
        c = __reduce_add_sync(full, c);
        if (c >= k) prefix = trial;
      }
      L = __uint_as_float(prefix);
    }
    L = fmaxf(L, ext_lower[t]);
    const float U = ext_upper[t];
    // list possibly too short?  (same test as refine_body; a_last = the smallest kept approximation -- the last entry
    // of a sorted list, but the lists of saeb_candidate_bounds_packed are not sorted)
    float a_last = INFINITY;
#pragma unroll
    for (int s = 0; s < RSW_SLOTS; ++s)
      if (lane + 32 * s < K2) a_last = fminf(a_last, a[s]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) a_last = fminf(a_last, __shfl_xor_sync(full, a_last, o));
    bool flagged = false;   // lane 0: a row is flagged at most once (flag_rows holds one slot per row of the call)
    if (lane == 0 && nv == K2 && a_last + 1.001f * (xn * dmax + xdn * wmax) + c_eps * xn * wmax >= L) {
      const int slot = atomicAdd(&status[0], 1);
      flag_rows[slot] = (int)t;
      flagged = true;
    }
    // classification: 1 = certain member that cannot enter its feature's list (not gathered), 3 = certain member that
    // can (exact value wanted), 2 = membership undecided (exact value needed), 0 = out
#pragma unroll
    for (int s = 0; s < RSW_SLOTS; ++s) {
      int c = 0;
      if (a[s] > 0.f) {
        const float l = lb[s];
        if (l > U && l > 0.f) {
          c = (feat_thr == nullptr || ub[s] >= feat_thr[f[s]]) ? 3 : 1;
          if (c == 1) ex[s] = 0.f;
        } else if (ub[s] >= L) {
          c = 2;
        }
      }
      st[s] = c;
    }
    // exact re-evaluation, one candidate at a time by the whole warp (same summation order as refine_body::evaluate)
    int n_gathered = 0;
#pragma unroll
    for (int s = 0; s < RSW_SLOTS; ++s) {
      uint32_t m = __ballot_sync(full, st[s] >= 2);
      while (m) {
        const int src = __ffs((int)m) - 1;
        m &= m - 1;
        const int fj = __shfl_sync(full, f[s], src);
        float val;
        if (fj == clamp_feature) {
          val = clamp_value;
        } else {
          const float* wr = W + (long long)fj * d;
          float acc0 = 0.f, acc1 = 0.f, acc2 = 0.f, acc3 = 0.f;
          if (vec) {
            const float4* w4 = reinterpret_cast<const float4*>(wr);
            const int n4 = (int)(d >> 2);
            int c = lane;
            for (; c + 7 * 32 < n4; c += 8 * 32) {
              float4 wv[8];
#pragma unroll
              for (int u = 0; u < 8; ++u) wv[u] = ldg_nc_f4(w4 + c + u * 32);
#pragma unroll
              for (int u = 0; u < 8; ++u) {
                const float4 xv = rsw_load_x4(xr, c + u * 32);
                acc0 = fmaf(wv[u].x, xv.x, acc0);
                acc1 = fmaf(wv[u].y, xv.y, acc1);
                acc2 = fmaf(wv[u].z, xv.z, acc2);
                acc3 = fmaf(wv[u].w, xv.w, acc3);
              }
            }
            for (; c < n4; c += 32) {
              const float4 wv = ldg_nc_f4(w4 + c);
              const float4 xv = rsw_load_x4(xr, c);
              acc0 = fmaf(wv.x, xv.x, acc0);
              acc1 = fmaf(wv.y, xv.y, acc1);
              acc2 = fmaf(wv.z, xv.z, acc2);
              acc3 = fmaf(wv.w, xv.w, acc3);
            }
          } else {
            for (long long i = lane; i < d; i += 32) acc0 = fmaf(wr[i], (float)xr[i], acc0);
          }
          float acc = (acc0 + acc1) + (acc2 + acc3);
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(full, acc, o);
          val = acc + bias[fj];
        }
        if (lane == src) ex[s] = (val > 0.f) ? val : -1.f;
        ++n_gathered;
      }
    }
    // hand every certain member and every positive boundary candidate to the caller (member value: MEMBER_SURE /
    // the exact value); rows with more potential members than output slots take the exact dense fallback
    int nout = 0;
#pragma unroll
    for (int s = 0; s < RSW_SLOTS; ++s) {
      float m = -1.f;
      if (st[s] == 1 || st[s] == 3) m = MEMBER_SURE;
      else if (st[s] == 2 && ex[s] > 0.f) m = ex[s];
      const bool o = m > 0.f;
      const uint32_t mask = __ballot_sync(full, o);
      const int pos = nout + __popc(mask & lt_mask);
      if (o && pos < k) {
        out_vals[t * k + pos] = fmaxf(ex[s], 0.f);
        out_member[t * k + pos] = m;
        out_idx[t * k + pos] = f[s];
      }
      nout += __popc(mask);
    }
    if (nout > k && lane == 0 && !flagged) {
      const int slot = atomicAdd(&status[0], 1);
      flag_rows[slot] = (int)t;
    }
    for (int j = (nout < k ? nout : k) + lane; j < k; j += 32) {
      out_vals[t * k + j] = 0.f;
      out_member[t * k + j] = 0.f;
      out_idx[t * k + j] = 0;
    }
    if (stats != nullptr && lane == 0 && n_gathered) atomicAdd(stats + 7, (unsigned long long)n_gathered);
  }
}

// ---------------------------------------------------------------------------------------------
// per-row lower bounds of the k best candidates (feature-sharded scan): lb_out[t][0..k) = the k largest values of
// a_j - eps_j, descending, floored at 0.  All-gathered across shards, their k-th largest is a lower bound of the
// token's global k-th activation.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
candidate_bounds_kernel(const float* __restrict__ cand_vals, const long long* __restrict__ cand_idx, int K2, int k,
                        const float* __restrict__ wnorm, const float* __restrict__ dnorm,
                        const float* __restrict__ xnorm, const float* __restrict__ xdnorm, float c_eps,
                        long long clamp_feature, float* __restrict__ lb_out, float* __restrict__ ub_out) {
  // ub_out (optional): the k largest UPPER bounds a_j + eps_j as well (descending): all-gathered, their (k+1)-th
  // largest bounds the token's global (k+1)-th value from above (value_mode 2 of the refinement)
  extern __shared__ float bsm[];   // [K2] lower bounds | [K2] upper bounds
  const long long t = blockIdx.x;
  const float xn = xnorm[t], xdn = xdnorm[t];
  for (int j = threadIdx.x; j < K2; j += blockDim.x) {
    const float av = cand_vals[t * K2 + j];
    const long long fj = cand_idx[t * K2 + j];
    const float wn = wnorm[fj];
    const float eps = (fj == clamp_feature) ? 0.f : 1.001f * (xn * dnorm[fj] + xdn * wn) + c_eps * xn * wn;
    bsm[j] = (av > 0.f) ? fmaxf(av - eps, 0.f) : 0.f;
    bsm[K2 + j] = (av > 0.f) ? av + eps : 0.f;
  }
  for (int j = threadIdx.x; j < k; j += blockDim.x) {
    lb_out[t * k + j] = 0.f;
    if (ub_out != nullptr) ub_out[t * k + j] = 0.f;
  }
  __syncthreads();
  for (int j = threadIdx.x; j < K2; j += blockDim.x) {
    const float l = bsm[j];
    int rank = 0;
    for (int i = 0; i < K2; ++i) rank += (bsm[i] > l || (bsm[i] == l && i < j)) ? 1 : 0;
    if (rank < k) lb_out[t * k + rank] = l;
    if (ub_out != nullptr) {
      const float u = bsm[K2 + j];
      rank = 0;
      for (int i = 0; i < K2; ++i) rank += (bsm[K2 + i] > u || (bsm[K2 + i] == u && i < j)) ? 1 : 0;
      if (rank < k) ub_out[t * k + rank] = u;
    }
  }
}

// ---------------------------------------------------------------------------------------------
// exact dense fallback for flagged rows: dense[slot][n] = relu(x_row . W[n] + bias[n]) in fp32, then a dense TopK
// ---------------------------------------------------------------------------------------------
template <typename XT>
__global__ void __launch_bounds__(256)
exact_rows_kernel(const XT* __restrict__ x, long long ld_x, const float* __restrict__ W, long long d, long long N,
                  const float* __restrict__ bias, const int* __restrict__ status, const int* __restrict__ flag_rows,
                  long long clamp_feature, float clamp_value, float* __restrict__ dense, int stage_x) {
  // stage_x = 0: the activation row is read from global memory instead of a shared-memory copy, so that the (normally
  // empty) grid needs no dynamic shared memory and is scheduled beside a resident GEMM CTA (same values either way)
  extern __shared__ float xsm[];
  const int slot = blockIdx.y;
  const int nflag = min(status[0], RF_MAX_FLAG);
  if (slot >= nflag) return;
  const long long t = flag_rows[slot];
  if (stage_x) {
    for (long long i = threadIdx.x; i < d; i += blockDim.x) xsm[i] = (float)x[t * ld_x + i];
    __syncthreads();
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  const long long per_block = (N + gridDim.x - 1) / gridDim.x;
  const long long n0 = (long long)blockIdx.x * per_block;
  const long long n1 = (n0 + per_block < N) ? n0 + per_block : N;
  for (long long n = n0 + warp; n < n1; n += nw) {
    const float* wr = W + n * d;
    float acc = 0.f;
    if (stage_x) {
      for (long long i = lane; i < d; i += 32) acc = fmaf(wr[i], xsm[i], acc);
    } else {
      for (long long i = lane; i < d; i += 32) acc = fmaf(wr[i], (float)x[t * ld_x + i], acc);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) {
      float v = acc + bias[n];
      if (n == clamp_feature) v = clamp_value;
      dense[(long long)slot * N + n] = fmaxf(v, 0.f);
    }
  }
}

// TopK of one dense non-negative row (value desc, index asc) by a whole block; `dsm` = kp2 uint2 of shared memory.
__device__ __forceinline__ void dense_topk_block(const float* row, long long N, int k, uint2* dsm, long long orow,
                                                 float* __restrict__ out_vals, long long* __restrict__ out_idx,
                                                 float* __restrict__ out_member = nullptr) {
  __shared__ int s_red[32];
  __shared__ int s_count;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nw = blockDim.x >> 5;
  auto block_sum = [&](int v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    __syncthreads();
    if (lane == 0) s_red[warp] = v;
    __syncthreads();
    int s = 0;
    for (int w = 0; w < nw; ++w) s += s_red[w];
    return s;
  };
  auto keyof = [&](long long i) { return __float_as_uint(fmaxf(row[i], 0.f)); };
  int kp2 = 2;
  while (kp2 < k) kp2 <<= 1;
  // k-th largest key
  uint32_t prefix = 0;
  for (int bit = 30; bit >= 0; --bit) {
    const uint32_t trial = prefix | (1u << bit);
    int c = 0;
    for (long long i = tid; i < N; i += blockDim.x) c += (keyof(i) >= trial) ? 1 : 0;
    if (block_sum(c) >= k) prefix = trial;
  }
  int c_gt = 0, c_eq = 0;
  for (long long i = tid; i < N; i += blockDim.x) {
    const uint32_t key = keyof(i);
    c_gt += key > prefix;
    c_eq += key == prefix;
  }
  c_gt = block_sum(c_gt);
  c_eq = block_sum(c_eq);
  const int need_eq = k - c_gt;
  uint32_t idx_cut = 0xffffffffu;
  if (c_eq > need_eq) {
    uint32_t p2 = 0;
    for (int bit = 31; bit >= 0; --bit) {
      const uint32_t trial = p2 | (1u << bit);
      int c = 0;
      for (long long i = tid; i < N; i += blockDim.x) c += (keyof(i) == prefix && (uint32_t)i < trial) ? 1 : 0;
      if (block_sum(c) < need_eq) p2 = trial;
    }
    idx_cut = p2;
  }
  __syncthreads();
  if (tid == 0) s_count = 0;
  for (int i = tid; i < kp2; i += blockDim.x) dsm[i] = make_uint2(0u, 0xffffffffu);
  __syncthreads();
  for (long long i = tid; i < N; i += blockDim.x) {
    const uint32_t key = keyof(i);
    if (key > prefix || (key == prefix && (uint32_t)i <= idx_cut)) {
      const int p = atomicAdd(&s_count, 1);
      if (p < kp2) dsm[p] = make_uint2(key, (uint32_t)i);
    }
  }
  __syncthreads();
  for (int size = 2; size <= kp2; size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      for (int p = tid; p < (kp2 >> 1); p += blockDim.x) {
        const int lo = ((p / stride) * stride * 2) + (p % stride);
        const int hi = lo + stride;
        const bool desc = ((lo & size) == 0);
        const uint2 a = dsm[lo], b = dsm[hi];
        const bool a_first = (a.x > b.x) || (a.x == b.x && a.y < b.y);
        if (a_first != desc) {
          dsm[lo] = b;
          dsm[hi] = a;
        }
      }
      __syncthreads();
    }
  }
  for (int i = tid; i < k; i += blockDim.x) {
    out_vals[orow * k + i] = __uint_as_float(dsm[i].x);
    out_idx[orow * k + i] = (long long)dsm[i].y;
    if (out_member != nullptr) out_member[orow * k + i] = __uint_as_float(dsm[i].x);   // exact values decide membership
  }
  __syncthreads();
}

// TopK of dense non-negative rows.  `row_map` (optional) redirects output rows; rows beyond `*n_rows_dev` (optional
// device count) exit.  One block per row.  Also serves Sae.select_topk on dense tensors.
__global__ void __launch_bounds__(1024)
dense_topk_kernel(const float* __restrict__ dense, long long ld, long long N, int k, const int* __restrict__ n_rows_dev,
                  int max_rows, const int* __restrict__ row_map, float* __restrict__ out_vals,
                  long long* __restrict__ out_idx, float* __restrict__ out_member) {
  extern __shared__ uint2 dsm[];   // [kp2] selected (value bits, index)
  const int slot = blockIdx.x;
  if (n_rows_dev != nullptr && slot >= min(*n_rows_dev, max_rows)) return;
  const long long orow = row_map ? row_map[slot] : slot;
  dense_topk_block(dense + (long long)slot * ld, N, k, dsm, orow, out_vals, out_idx, out_member);
}

// Flagged rows beyond the first RF_MAX_FLAG (degenerate inputs: massive ties, k close to N): block b walks the slots
// RF_MAX_FLAG + b, RF_MAX_FLAG + b + gridDim.x, ...; per slot it writes the exact dense row into its private scratch
// row and takes the TopK from it.  Exits at once in the normal case.
template <typename XT>
__global__ void __launch_bounds__(1024)
overflow_rows_kernel(const XT* __restrict__ x, long long ld_x, const float* __restrict__ W, long long d, long long N,
                     const float* __restrict__ bias, const int* __restrict__ status, const int* __restrict__ flag_rows,
                     long long clamp_feature, float clamp_value, float* dense, int k, float* __restrict__ out_vals,
                     long long* __restrict__ out_idx, float* __restrict__ out_member, int stage_x) {
  extern __shared__ uint2 osm[];   // [kp2] uint2 | [d] float (stage_x != 0 only, see exact_rows_kernel)
  const int nflag = status[0];
  if (nflag <= RF_MAX_FLAG) return;
  int kp2 = 2;
  while (kp2 < k) kp2 <<= 1;
  float* xs = reinterpret_cast<float*>(osm + kp2);
  float* my = dense + (long long)blockIdx.x * N;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  for (int slot = RF_MAX_FLAG + blockIdx.x; slot < nflag; slot += gridDim.x) {
    const long long t = flag_rows[slot];
    if (stage_x) {
      for (long long i = threadIdx.x; i < d; i += blockDim.x) xs[i] = (float)x[t * ld_x + i];
    }
    __syncthreads();
    for (long long n = warp; n < N; n += nw) {
      const float* wr = W + n * d;
      float acc = 0.f;
      if (stage_x) {
        for (long long i = lane; i < d; i += 32) acc = fmaf(wr[i], xs[i], acc);
      } else {
        for (long long i = lane; i < d; i += 32) acc = fmaf(wr[i], (float)x[t * ld_x + i], acc);
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
      if (lane == 0) {
        float v = acc + bias[n];
        if (n == clamp_feature) v = clamp_value;
        my[n] = fmaxf(v, 0.f);
      }
    }
    __syncthreads();
    dense_topk_block(my, N, k, osm, t, out_vals, out_idx, out_member);
  }
}

}  // namespace saeb
