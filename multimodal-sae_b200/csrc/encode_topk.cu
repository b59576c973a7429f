// Fused SAE encoder:  pre = relu(x @ W_enc^T + bias_folded)  ->  per-row TopK, dense latents never written.
//
// Replaces the reference chain  Sae.pre_acts (sae/sae.py:172-177, nn.Linear -> fp32 GEMM, dense [T,N] output)
// + Sae.select_topk / torch.topk (sae/sae.py:179-181, features/cache.py:211-213).
//
// Structure (one persistent CTA, or CTA pair with cta_group::2, per SM):
//   warp 0     : TMA producer   (cp.async.bulk.tensor, 128B-swizzled K-major tiles of x and W_enc planes)
//   warp 1     : MMA issuer     (tcgen05.mma kind::f16, fp32 accumulator 128 x 256 per CTA in TMEM, double buffered)
//   warp 2     : TMEM allocator
//   warps 4..7 : epilogue       (tcgen05.ld -> + folded bias -> running per-row threshold filter -> candidate list;
//                                warp-cooperative radix-select compaction when a list fills up)
// A work unit is (m_tile, n_split): a tile of 128*PAIR token rows swept over a contiguous range of 256-wide feature
// tiles.  Each (row, split) produces a candidate list that is a superset of the split's top-k; `topk_merge_kernel`
// selects the exact top-k per row with the canonical order (value desc, index asc).
//
// Precision: W_enc is stored as BP bf16 planes (hi, lo) whose sum reproduces fp32 W_enc to ~2^-17; x as AP planes
// (1 for bf16 activations, which are exact; 2 for fp16/fp32 activations).  All plane products accumulate into the
// same fp32 TMEM accumulator.  The decoder bias is folded: W(x - b_dec) + b_enc = Wx + (b_enc - W b_dec).
#include <cuda.h>
#include <cstdlib>

#include "kernels_topk_select.cuh"

namespace saeb {

constexpr int BM = 128;   // token rows per CTA
constexpr int BN = 256;   // feature columns per MMA tile (UMMA N)
constexpr int BK = 64;    // K elements per pipeline stage (128 bytes of bf16 = one swizzle row)
constexpr int UMMA_K = 16;
constexpr int NUM_THREADS = 256;
constexpr int SMEM_BUDGET = 232448;   // 227 KB
constexpr int SMEM_MISC = 12288;      // barriers + tmem ptr + bias staging + compaction histograms

template <int AP, int BP, int PAIR>
struct EncCfg {
  static constexpr int B_ROWS = BN / PAIR;
  static constexpr int A_PLANE = BM * BK * 2;
  static constexpr int B_PLANE = B_ROWS * BK * 2;
  static constexpr int STAGE = AP * A_PLANE + BP * B_PLANE;
  static constexpr int STAGES_RAW = (SMEM_BUDGET - SMEM_MISC - 1024) / STAGE;
  static constexpr int STAGES = STAGES_RAW > 8 ? 8 : STAGES_RAW;
  static constexpr int SMEM = STAGES * STAGE + SMEM_MISC + 1024;
  static_assert(STAGES >= 2, "need at least two pipeline stages");
};

struct EncodeArgs {
  int T, d, N, k;
  int num_m_tiles;    // ceil(T / (BM*PAIR))
  int num_n_tiles;    // ceil(N / BN)
  int S;              // feature-range splits per m tile
  int num_k_blocks;   // ceil(d / BK)
  int pass_mask;      // bit (a*BP+b) set -> issue MMA for (A plane a, B plane b)
  int clamp_col;      // steering: column forced to clamp_val before TopK (-1 = none)
  float clamp_val;
  unsigned long long* stats;   // optional diagnostics: cycle counters summed over CTAs (see saeb_debug_stats)
  int stages;       // depth of the TMA -> MMA shared-memory ring actually used (<= EncCfg::STAGES)
  int prefetch_b;   // weight tiles are pulled into L2 this many feature tiles ahead (0 = off), one pair per tile
  int dbg;   // diagnostics only (wrong results): bit0 = every cluster loads token tile 0, bit1 = every step loads feature tile 0
  unsigned long long hint_a, hint_b;   // L2 eviction policies of the activation / weight TMA loads
  unsigned int idesc;       // tcgen05 instruction descriptor (operand formats are a run-time choice: bf16 or fp16)
  const float* row_scale;   // optional [T]: per-row power-of-two factor undoing the activation pre-scale
  const float* w_unscale;   // optional device scalar undoing the weight pre-scale
  const float* bias;  // folded bias [N]
  uint2* cand;        // [T][S][CAP] (value bits, column)
  int* cand_cnt;      // [T][S]
  float* dense_out;   // optional dense relu(pre) [T][ld_dense]
  long long ld_dense;
};

// ---------------------------------------------------------------------------------------------
// the fused kernel
// ---------------------------------------------------------------------------------------------
// CL = CTA pairs per cluster.  CL == 2 (clusters of four CTAs, only with PAIR == 2 and S == 2): the two pairs of a
// cluster work on the SAME token tile, pair p on feature-range split p, so they need the same activation tile in every
// pipeline step.  Each of the four CTAs fetches one quarter of it (64 rows) and TMA-multicasts it to its counterpart in
// the other pair: the activation tile is read from L2 once per cluster instead of once per pair (-25 % L2 -> SM
// traffic; the kernel runs at the L2 slices' throughput limit, not at the tensor pipe's).  A stage of the ring is
// written into both pairs, so it is recycled only when BOTH pairs' MMAs have consumed it (empty barriers count CL
// commits, each multicast to all four CTAs); both pairs walk the same number of feature tiles.
template <int AP, int BP, int PAIR, int SLOTS, int CL = 1>
__global__ void __launch_bounds__(NUM_THREADS, 1)
encode_topk_kernel(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_b,
                   const EncodeArgs args) {
  static_assert(CL == 1 || (CL == 2 && PAIR == 2), "clusters of two pairs need the CTA-pair MMA");
  using Cfg = EncCfg<AP, BP, PAIR>;
  const int STAGES = args.stages;   // run-time ring depth: a shallower ring leaves shared memory for co-resident gather CTAs
  constexpr int CAP = 32 * SLOTS;
  constexpr uint32_t TMEM_COLS = 512;

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* misc = smem + STAGES * Cfg::STAGE;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(misc);   // [STAGES]
  uint64_t* empty_bar = full_bar + STAGES;                  // [STAGES]
  uint64_t* tfull_bar = empty_bar + STAGES;                 // [2]
  uint64_t* tempty_bar = tfull_bar + 2;                     // [2]
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tempty_bar + 2);
  float* bias_s = reinterpret_cast<float*>(misc + 1024);    // [4 warps][BN]
  int* hist_s = reinterpret_cast<int*>(misc + 1024 + 4 * BN * 4);   // [4 warps][256]

  const uint32_t warp = threadIdx.x >> 5;
  const uint32_t lane = lane_id();
  const uint32_t crank = (PAIR == 2) ? cluster_ctarank() : 0;   // rank in the cluster: 0 .. PAIR * CL - 1
  const uint32_t cta_rank = crank & (PAIR - 1);                   // position in the CTA pair
  const uint32_t pair_id = (CL == 2) ? (crank >> 1) : 0;          // which pair of the cluster
  const uint32_t leader_rank = pair_id * 2;                       // cluster rank of this pair's leader CTA
  const bool leader = cta_rank == 0;
  const uint16_t pair_mask = (uint16_t)(3u << (2 * pair_id));     // the two CTAs of this pair
  const uint16_t all_mask = (CL == 2) ? (uint16_t)0xF : (uint16_t)3;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_a);
    tma_prefetch_desc(&tm_b);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < STAGES; ++i) {
      mbar_init(&full_bar[i], PAIR);
      mbar_init(&empty_bar[i], CL);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull_bar[i], 1);
      mbar_init(&tempty_bar[i], PAIR * 4);
    }
    fence_mbar_init();
  }
  if constexpr (PAIR == 2) cluster_sync_all();
  if (warp == 2) tmem_alloc<PAIR>(tmem_ptr, TMEM_COLS);
  tc_fence_before();
  if constexpr (PAIR == 2) cluster_sync_all(); else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  const bool timed = args.stats != nullptr;   // cycle accounting only when diagnostics are on
  const int cluster_id = blockIdx.x / (PAIR * CL);
  const int num_clusters = gridDim.x / (PAIR * CL);
  // CL == 1: a unit is (token tile, split), one per pair; CL == 2: a unit is a token tile, pair p takes split p (S == 2)
  const int num_units = (CL == 2) ? args.num_m_tiles : args.num_m_tiles * args.S;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      long long w_prod = 0;
      const long long t_begin = timed ? clock64() : 0;
      for (int u = cluster_id; u < num_units; u += num_clusters) {
        const int split = (CL == 2) ? (int)pair_id : u % args.S, m_tile = (CL == 2) ? u : u / args.S;
        const int nt0 = (int)((long long)split * args.num_n_tiles / args.S);
        const int nt1 = (int)((long long)(split + 1) * args.num_n_tiles / args.S);
        const int m0 = (((args.dbg & 1) ? 0 : m_tile) * PAIR + (int)cta_rank) * BM;
        for (int nt = nt0; nt < nt1; ++nt) {
          const int n0 = ((args.dbg & 2) ? 0 : nt) * BN + (int)cta_rank * Cfg::B_ROWS;
          if (args.prefetch_b > 0) {
            // every pair of a split sweeps the same weight tiles; exactly one of them (round robin over the token
            // tiles of this launch) asks L2 for a tile a few steps before the whole group needs it, so that the
            // group's demand loads hit in L2 instead of stalling on HBM together
            const int ntp = nt + args.prefetch_b;
            if (ntp < nt1 && (ntp % args.num_m_tiles) == m_tile) {
              const int np0 = ntp * BN + (int)cta_rank * Cfg::B_ROWS;
              for (int kb = 0; kb < args.num_k_blocks; ++kb)
#pragma unroll
                for (int b = 0; b < BP; ++b) tma_prefetch_3d(&tm_b, kb * BK, np0, b);
            }
          }
          for (int kb = 0; kb < args.num_k_blocks; ++kb) {
            const long long tw0 = timed ? clock64() : 0;
            mbar_wait(&empty_bar[stage], phase ^ 1);
            if (timed) w_prod += clock64() - tw0;
            uint8_t* sa = smem + stage * Cfg::STAGE;
            uint8_t* sb = sa + AP * Cfg::A_PLANE;
            if constexpr (PAIR == 1) {
              mbar_arrive_expect_tx(&full_bar[stage], Cfg::STAGE);
#pragma unroll
              for (int a = 0; a < AP; ++a)
                tma_load_3d(sa + a * Cfg::A_PLANE, &tm_a, &full_bar[stage], kb * BK, m0, a, args.hint_a);
#pragma unroll
              for (int b = 0; b < BP; ++b)
                tma_load_3d(sb + b * Cfg::B_PLANE, &tm_b, &full_bar[stage], kb * BK, n0, b, args.hint_b);
            } else {
              if (leader) mbar_arrive_expect_tx(&full_bar[stage], Cfg::STAGE * 2);
              else mbar_arrive_cluster(&full_bar[stage], leader_rank);
              if constexpr (CL == 2) {
                // my quarter of the activation tile (64 rows), delivered to me and to my counterpart in the other pair
                const uint16_t mask_a = (uint16_t)((1u << cta_rank) | (1u << (2 + cta_rank)));
#pragma unroll
                for (int a = 0; a < AP; ++a)
                  tma_load_3d_pair_mc(sa + a * Cfg::A_PLANE + pair_id * (Cfg::A_PLANE / 2), &tm_a, &full_bar[stage],
                                      kb * BK, m0 + (int)pair_id * (BM / 2), a, mask_a, args.hint_a);
              } else {
#pragma unroll
                for (int a = 0; a < AP; ++a)
                  tma_load_3d_pair(sa + a * Cfg::A_PLANE, &tm_a, &full_bar[stage], kb * BK, m0, a, args.hint_a);
              }
#pragma unroll
              for (int b = 0; b < BP; ++b)
                tma_load_3d_pair(sb + b * Cfg::B_PLANE, &tm_b, &full_bar[stage], kb * BK, n0, b, args.hint_b);
            }
            if (++stage == STAGES) { stage = 0; phase ^= 1; }
          }
        }
      }
      if (args.stats != nullptr && leader) {
        atomicAdd(args.stats + 0, (unsigned long long)w_prod);
        atomicAdd(args.stats + 5, (unsigned long long)(clock64() - t_begin));
        atomicAdd(args.stats + 6, 1ull);
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (leader CTA only) =====================
    if (leader && lane == 0) {
      const uint32_t idesc = args.idesc;
      int stage = 0;
      uint32_t phase = 0;
      uint32_t tile_iter = 0;
      long long w_tempty = 0, w_full = 0;
      for (int u = cluster_id; u < num_units; u += num_clusters) {
        const int split = (CL == 2) ? (int)pair_id : u % args.S;
        const int nt0 = (int)((long long)split * args.num_n_tiles / args.S);
        const int nt1 = (int)((long long)(split + 1) * args.num_n_tiles / args.S);
        for (int nt = nt0; nt < nt1; ++nt, ++tile_iter) {
          const uint32_t acc_stage = tile_iter & 1, acc_phase = (tile_iter >> 1) & 1;
          const long long tw0 = timed ? clock64() : 0;
          mbar_wait(&tempty_bar[acc_stage], acc_phase ^ 1);
          if (timed) w_tempty += clock64() - tw0;
          tc_fence_after();
          const uint32_t tmem_d = tmem_base + acc_stage * BN;
          for (int kb = 0; kb < args.num_k_blocks; ++kb) {
            const long long tw1 = timed ? clock64() : 0;
            mbar_wait(&full_bar[stage], phase);
            if (timed) w_full += clock64() - tw1;
            tc_fence_after();
            const uint32_t sa = smem_u32(smem + stage * Cfg::STAGE);
            const uint32_t sb = sa + AP * Cfg::A_PLANE;
            uint32_t first = (kb == 0) ? 1u : 0u;
#pragma unroll
            for (int a = 0; a < AP; ++a) {
#pragma unroll
              for (int b = 0; b < BP; ++b) {
                if (!((args.pass_mask >> (a * BP + b)) & 1)) continue;
                const uint64_t da = make_smem_desc_sw128(sa + a * Cfg::A_PLANE);
                const uint64_t db = make_smem_desc_sw128(sb + b * Cfg::B_PLANE);
#pragma unroll
                for (int ks = 0; ks < BK / UMMA_K; ++ks) {
                  // advance 32 bytes (16 bf16) along K inside the 128B swizzle row: +2 in 16-byte units
                  umma_f16<PAIR>(tmem_d, da + 2 * ks, db + 2 * ks, idesc, first ? 0u : 1u);
                  first = 0;
                }
              }
            }
            umma_commit<PAIR>(&empty_bar[stage], all_mask);   // frees the stage in every CTA that writes into it
            if (kb == args.num_k_blocks - 1) umma_commit<PAIR>(&tfull_bar[acc_stage], pair_mask);
            if (++stage == STAGES) { stage = 0; phase ^= 1; }
          }
        }
      }
      if (args.stats != nullptr) {
        atomicAdd(args.stats + 1, (unsigned long long)w_tempty);
        atomicAdd(args.stats + 2, (unsigned long long)w_full);
      }
    }
  } else if (warp >= 4) {
    // ===================== epilogue: bias + ReLU + running TopK filter =====================
    const uint32_t q = warp & 3;                 // TMEM lane quarter owned by this warp
    float* bias_w = bias_s + q * BN;
    const uint32_t full = 0xffffffffu;
    uint32_t tile_iter = 0;
    long long w_tfull = 0, w_compact = 0;
    for (int u = cluster_id; u < num_units; u += num_clusters) {
      const int split = (CL == 2) ? (int)pair_id : u % args.S, m_tile = (CL == 2) ? u : u / args.S;
      const int nt0 = (int)((long long)split * args.num_n_tiles / args.S);
      const int nt1 = (int)((long long)(split + 1) * args.num_n_tiles / args.S);
      const int row = (m_tile * PAIR + (int)cta_rank) * BM + (int)(q * 32 + lane);
      const bool valid = row < args.T;
      const bool do_topk = args.cand != nullptr;
      uint2* cand = do_topk && valid ? args.cand + ((size_t)row * args.S + split) * CAP : nullptr;
      float thr = (valid && do_topk) ? 0.0f : __int_as_float(0x7f800000);   // +inf: never append
      int cnt = 0;
      float sc = 1.0f;   // acc * sc + bias; exact powers of two when the operands were pre-scaled
      if (args.row_scale != nullptr && valid) sc = __ldg(args.row_scale + row) * __ldg(args.w_unscale);
      for (int nt = nt0; nt < nt1; ++nt, ++tile_iter) {
        const uint32_t acc_stage = tile_iter & 1, acc_phase = (tile_iter >> 1) & 1;
        const int n_base = nt * BN;
        // stage this tile's folded bias (columns >= N get -inf so they can never be selected)
        __syncwarp();
#pragma unroll
        for (int i = 0; i < BN / 32; ++i) {
          int c = n_base + i * 32 + lane;
          bias_w[i * 32 + lane] = (c < args.N) ? __ldg(args.bias + c) : __int_as_float(0xff800000);
        }
        __syncwarp();
        const long long twe = timed ? clock64() : 0;
        mbar_wait(&tfull_bar[acc_stage], acc_phase);
        if (timed) w_tfull += clock64() - twe;
        tc_fence_after();
        const uint32_t taddr = tmem_base + ((q * 32u) << 16) + acc_stage * BN;
#pragma unroll 1
        for (int c = 0; c < BN / 32; ++c) {
          uint32_t r[32];
          tmem_ld_32x32(taddr + c * 32, r);
          tmem_ld_wait();
          if (c == BN / 32 - 1) {
            // accumulator fully read: hand the TMEM stage back to the MMA issuer (leader CTA's barrier)
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
              if constexpr (PAIR == 1) mbar_arrive(&tempty_bar[acc_stage]);
              else mbar_arrive_cluster(&tempty_bar[acc_stage], leader_rank);
            }
          }
          const int col0 = n_base + c * 32;
          const int cj = args.clamp_col - col0;   // in [0,32) iff the clamped column is in this chunk
          float v[32];
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            float4 b4 = *reinterpret_cast<const float4*>(bias_w + c * 32 + j);
            v[j + 0] = fmaf(__uint_as_float(r[j + 0]), sc, b4.x);
            v[j + 1] = fmaf(__uint_as_float(r[j + 1]), sc, b4.y);
            v[j + 2] = fmaf(__uint_as_float(r[j + 2]), sc, b4.z);
            v[j + 3] = fmaf(__uint_as_float(r[j + 3]), sc, b4.w);
          }
          if (cj >= 0 && cj < 32) {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (j == cj) v[j] = args.clamp_val;
          }
          if (args.dense_out != nullptr && valid) {
            float* o = args.dense_out + (size_t)row * args.ld_dense + col0;
            if (col0 + 32 <= args.N) {
#pragma unroll
              for (int j = 0; j < 32; j += 4)
                *reinterpret_cast<float4*>(o + j) =
                    make_float4(fmaxf(v[j], 0.f), fmaxf(v[j + 1], 0.f), fmaxf(v[j + 2], 0.f), fmaxf(v[j + 3], 0.f));
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j)
                if (col0 + j < args.N) o[j] = fmaxf(v[j], 0.f);
            }
          }
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            if (v[j] > thr) {
              cand[cnt] = make_uint2(__float_as_uint(v[j]), (uint32_t)(col0 + j));
              ++cnt;
            }
          }
          // compaction when a list could overflow during the next chunk
          uint32_t need = __ballot_sync(full, cnt > CAP - 32);
          const bool had = timed && need != 0;
          const long long twc = had ? clock64() : 0;
          while (need) {
            const int src = __ffs(need) - 1;
            need &= need - 1;
            const int src_cnt = __shfl_sync(full, cnt, src);
            const unsigned long long p = __shfl_sync(full, (unsigned long long)(uintptr_t)cand, src);
            float nthr;
            int ncnt;
            __syncwarp();
            compact_row<SLOTS>(reinterpret_cast<uint2*>((uintptr_t)p), src_cnt, args.k, lane, hist_s + q * 256, nthr,
                               ncnt);
            if ((int)lane == src) {
              thr = nthr;
              cnt = ncnt;
            }
          }
          if (had) w_compact += clock64() - twc;
        }
      }
      if (do_topk && valid) args.cand_cnt[(size_t)row * args.S + split] = cnt;
    }
    if (args.stats != nullptr && leader && warp == 4 && lane == 0) {
      atomicAdd(args.stats + 3, (unsigned long long)w_tfull);
      atomicAdd(args.stats + 4, (unsigned long long)w_compact);
    }
  }

  // ===================== teardown =====================
  tc_fence_before();
  if constexpr (PAIR == 2) cluster_sync_all(); else __syncthreads();
  if (warp == 2) tmem_dealloc<PAIR>(tmem_base, TMEM_COLS);
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode_fn() {
  static PFN_encodeTiled fn = nullptr;
  if (fn) return fn;
  void* p = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
      q != cudaDriverEntryPointSuccess)
    return nullptr;
  fn = reinterpret_cast<PFN_encodeTiled>(p);
  return fn;
}

// 3-D map over [planes][rows][cols] of 16-bit elements, box = (64 cols, box_rows rows, 1 plane), 128B swizzle.
static int make_map(CUtensorMap* map, const void* base, long long cols, long long rows, long long planes,
                    long long row_stride_bytes, long long plane_stride_bytes, int box_rows) {
  PFN_encodeTiled fn = get_encode_fn();
  SAEB_REQUIRE(fn != nullptr, "cuTensorMapEncodeTiled entry point not available");
  cuuint64_t gdim[3] = {(cuuint64_t)cols, (cuuint64_t)rows, (cuuint64_t)planes};
  cuuint64_t gstr[2] = {(cuuint64_t)row_stride_bytes, (cuuint64_t)plane_stride_bytes};
  cuuint32_t box[3] = {(cuuint32_t)BK, (cuuint32_t)box_rows, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), gdim, gstr, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  SAEB_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
  return 0;
}

static int g_num_sms = 0;
static int num_sms() {
  if (g_num_sms == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev);
  }
  return g_num_sms;
}

int cap_for_k(int k) { return k <= 128 ? 256 : (k <= 256 ? 512 : 1024); }

static thread_local int g_l2_hints = 1;   // 1: activation tiles evict_last (they are re-read once per feature tile), weights normal
int set_l2_hints(int v) {
  g_l2_hints = v;
  return 0;
}
static thread_local int g_persist_a = 1;   // 1: pin the activation planes in the persisting part of L2 for the duration of the launch
static size_t g_persist_bytes = 0;
int set_persist_a(int v) {
  g_persist_a = v;
  if (v && g_persist_bytes == 0) {
    int dev = 0, max_persist = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&max_persist, cudaDevAttrMaxPersistingL2CacheSize, dev);
    if (max_persist > 0 && cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, (size_t)max_persist) == cudaSuccess)
      g_persist_bytes = (size_t)max_persist;
  }
  return 0;
}
long long persist_bytes() { return (long long)g_persist_bytes; }
static unsigned long long* g_stats = nullptr;   // device counters [8] when option "stats" is on
int set_stats(int v) {
  if (v && g_stats == nullptr) {
    if (cudaMalloc(&g_stats, 64) != cudaSuccess) return -2;
  }
  if (g_stats) cudaMemset(g_stats, 0, 64);
  if (!v && g_stats) {
    cudaFree(g_stats);
    g_stats = nullptr;
  }
  return 0;
}
unsigned long long* stats_ptr() { return g_stats; }
int read_stats(unsigned long long* out8) {
  if (!g_stats) return -1;
  return cudaMemcpy(out8, g_stats, 64, cudaMemcpyDeviceToHost) == cudaSuccess ? 0 : -2;
}
static thread_local int g_prefetch_b = 0;   // measured: no gain (7.07-7.12 ms per wave without, 7.03-7.19 with), kept as an option
int set_prefetch_b(int v) {
  g_prefetch_b = v < 0 ? 0 : v;
  return 0;
}
static thread_local int g_dbg = 0;
int set_dbg(int v) {
  g_dbg = v;
  return 0;
}
static thread_local int g_splits = 0;     // 0 = automatic
int set_splits(int v) {
  if (v < 0 || v > 64) {
    set_error("splits must be in 0..64 (0 = automatic)");
    return -1;
  }
  g_splits = v;
  return 0;
}

static int merge_max_splits(int cap, int k) {
  int kp2 = 2;
  while (kp2 < k) kp2 <<= 1;
  const int s = (200 * 1024 - kp2 * 8) / (cap * 8);   // the merge kernel stages a row's candidates in shared memory
  return s < 1 ? 1 : s;
}

// Feature-range splits for a launch over `num_m_tiles` token tiles: (token tiles x splits) should fill the machine in
// whole waves.  Every split restarts its running threshold from zero (a compaction-heavy first few tiles), so the
// smallest S with the best wave efficiency wins.
static int choose_splits(int num_m_tiles, int num_n_tiles, int num_clusters, int cap, int k) {
  int s_cap = merge_max_splits(cap, k);
  if (s_cap > num_n_tiles) s_cap = num_n_tiles;
  if (s_cap < 1) s_cap = 1;
  int S;
  if (g_splits > 0) {
    S = g_splits;
  } else if (num_m_tiles > num_clusters) {
    S = 2;
  } else {
    int s0 = num_clusters / num_m_tiles;
    if (s0 < 1) s0 = 1;
    S = s0;
    double best = -1.0;
    for (int c = s0; c <= s0 + 2; ++c) {
      const long long units = (long long)num_m_tiles * c;
      const long long waves = (units + num_clusters - 1) / num_clusters;
      const double eff = (double)units / (double)(waves * num_clusters);
      if (eff > best + 1e-9) {
        best = eff;
        S = c;
      }
    }
  }
  if (S > s_cap) S = s_cap;
  if (S < 1) S = 1;
  return S;
}

// A call is executed as a sequence of launches over row chunks.  One launch covers at most (clusters / 2) token tiles
// = one wave of (tile, split) units at S = 2: measured with ncu, a single-wave launch keeps its 37 live activation
// tiles (74 MB) L2-resident (3.6 GB of DRAM reads per 9472 tokens with the persisting window, 13 GB without), while
// a multi-wave launch lets waves drift apart and re-fetches the activations from HBM (233 GB per 65536 tokens),
// which costs ~20 % of the SM clock under the power cap.
struct ChunkPlan {
  long long t0, rows;
  int num_m_tiles, S, grid;
  size_t cnt_off, cand_off;
};
struct EncodePlan {
  int pair, cap, num_n_tiles, n_chunks;
  static constexpr int MAX_CHUNKS = 4096;   // 38.8 M tokens per call at one 9472-token wave per chunk
  ChunkPlan chunks[MAX_CHUNKS];
  size_t total_bytes;
};
static thread_local int g_chunking = 1;   // 0: one launch for the whole call
// SMs left free by the persistent GEMM grid.  The feature-sharded scan runs its collectives, refinement and list
// update on a second stream while the next chunk's GEMM is in flight; a grid that owns every SM would make those
// kernels (NCCL's in particular: too many registers to co-reside) wait for a launch boundary and then displace GEMM
// CTAs, which doubles that launch.
static thread_local int g_reserve_sms = 0;
int set_reserve_sms(int v) {
  g_reserve_sms = v < 0 ? 0 : v;
  return 0;
}
int set_chunking(int v) {
  g_chunking = v;
  return 0;
}

static bool make_plan(EncodePlan& p, long long T, long long N, int k, int pair) {
  p.pair = pair;
  p.cap = cap_for_k(k);
  p.num_n_tiles = (int)((N + BN - 1) / BN);
  int sms = num_sms() > 0 ? num_sms() : 148;
  if (g_reserve_sms > 0 && sms - g_reserve_sms >= 2 * pair) sms -= g_reserve_sms;
  const int clusters = sms / pair;
  const long long tile_rows = (long long)BM * pair;
  long long rows_full = (long long)(clusters / 2 > 0 ? clusters / 2 : 1) * tile_rows;
  if (!g_chunking || g_splits > 0) rows_full = T;
  p.n_chunks = 0;
  size_t off = 0;
  for (long long t0 = 0; t0 < T; t0 += rows_full) {
    if (p.n_chunks >= EncodePlan::MAX_CHUNKS) return false;
    ChunkPlan& c = p.chunks[p.n_chunks++];
    c.t0 = t0;
    c.rows = (T - t0 < rows_full) ? T - t0 : rows_full;
    c.num_m_tiles = (int)((c.rows + tile_rows - 1) / tile_rows);
    c.S = choose_splits(c.num_m_tiles, p.num_n_tiles, clusters, p.cap, k);
    const int units = c.num_m_tiles * c.S;
    c.grid = (units < clusters ? units : clusters) * pair;
    c.cnt_off = off;
    off += (((size_t)c.rows * c.S * sizeof(int)) + 255) & ~(size_t)255;
    c.cand_off = off;
    off += (size_t)c.rows * c.S * p.cap * sizeof(uint2);
  }
  p.total_bytes = off;
  return true;
}

static thread_local int g_profile = 0;
static thread_local cudaEvent_t g_ev0 = nullptr, g_ev1 = nullptr;   // per calling thread, recorded on the call's stream
static thread_local bool g_ev_valid = false;
int set_profile(int v) {
  g_profile = v ? 1 : 0;
  if (g_profile && g_ev0 == nullptr) {
    if (cudaEventCreate(&g_ev0) != cudaSuccess || cudaEventCreate(&g_ev1) != cudaSuccess) {
      set_error("profile: cudaEventCreate failed");
      return -2;
    }
  }
  return 0;
}
// elapsed milliseconds of the most recent fused encode kernel (main kernel only, merge excluded); < 0 if none
float last_encode_ms() {
  if (!g_ev_valid) return -1.f;
  float ms = -1.f;
  if (cudaEventSynchronize(g_ev1) != cudaSuccess) return -1.f;
  if (cudaEventElapsedTime(&ms, g_ev0, g_ev1) != cudaSuccess) return -1.f;
  return ms;
}
static thread_local int g_cta_pair = 0;   // 0 = not initialised (env SAEB_CTA_PAIR or default 2)
static int default_pair() {
  if (g_cta_pair == 0) {
    const char* e = getenv("SAEB_CTA_PAIR");
    g_cta_pair = (e && e[0] == '1') ? 1 : 2;
  }
  return g_cta_pair;
}
int set_cta_pair(int v) {
  if (v != 1 && v != 2) {
    set_error("cta_pair must be 1 or 2");
    return -1;
  }
  g_cta_pair = v;
  return 0;
}

size_t encode_workspace_bytes(long long T, long long d, long long N, int k) {
  // upper bound over both tile modes (the plan depends on the option values at call time)
  (void)d;
  size_t best = 0;
  for (int pair = 1; pair <= 2; ++pair) {
    static thread_local EncodePlan plan;
    if (!make_plan(plan, T, N, k, pair)) continue;
    if (plan.total_bytes > best) best = plan.total_bytes;
  }
  // headroom for option changes between the query and the call (splits override, chunking off)
  const size_t cap = cap_for_k(k);
  const size_t s8 = (size_t)T * 8 * cap * sizeof(uint2) + (size_t)T * 8 * sizeof(int) + 4096;
  if (T > 4096 && s8 > best && merge_max_splits((int)cap, k) >= 8) best = s8;
  return best + 1024;
}

// depth of the smem ring (0 = as deep as fits: 6 stages of 32 KB in the single-pass pair mode).  5 leaves ~54 KB of
// shared memory per SM to the gather CTAs that run beside the GEMM (saeb200.overlap).
static thread_local int g_gemm_stages = 0;
int set_gemm_stages(int v) {
  if (v != 0 && (v < 2 || v > 8)) {
    set_error("gemm_stages must be 0 (automatic) or 2..8");
    return -1;
  }
  g_gemm_stages = v;
  return 0;
}

struct PersistWindow {
  void* base = nullptr;
  size_t bytes = 0;
};
static thread_local PersistWindow g_window;   // set by the launching thread just before each launch

template <int AP, int BP, int PAIR, int SLOTS, int CL = 1>
static int launch_cfg(const CUtensorMap& ta, const CUtensorMap& tb, const EncodeArgs& args, int grid,
                      cudaStream_t stream, int* max_clusters_out = nullptr) {
  using Cfg = EncCfg<AP, BP, PAIR>;
  auto kern = encode_topk_kernel<AP, BP, PAIR, SLOTS, CL>;
  SAEB_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM));
  SAEB_CARVEOUT(kern);
  EncodeArgs largs = args;
  largs.stages = (g_gemm_stages >= 2 && g_gemm_stages < Cfg::STAGES) ? g_gemm_stages : Cfg::STAGES;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(NUM_THREADS);
  cfg.dynamicSmemBytes = largs.stages * Cfg::STAGE + SMEM_MISC + 1024;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = PAIR * CL;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  if (max_clusters_out != nullptr) {   // occupancy query only: how many clusters of this shape can be resident at once
    int n = 0;
    cfg.gridDim = dim3(PAIR * CL * 1024);
    if (cudaOccupancyMaxActiveClusters(&n, kern, &cfg) != cudaSuccess) {
      cudaGetLastError();
      n = 0;
    }
    *max_clusters_out = n;
    return 0;
  }
  if (g_persist_a && g_persist_bytes > 0 && g_window.bytes > 0) {
    // the activation tiles are re-read once per feature tile: keep them in the persisting set-aside of L2
    attr[1].id = cudaLaunchAttributeAccessPolicyWindow;
    attr[1].val.accessPolicyWindow.base_ptr = g_window.base;
    attr[1].val.accessPolicyWindow.num_bytes = g_window.bytes;
    double ratio = (double)g_persist_bytes / (double)g_window.bytes;
    // mode 1: persisting fraction limited to the set-aside size, the rest streams; 2: everything persisting, the
    // hardware arbitrates; 3: limited fraction, the rest normal
    attr[1].val.accessPolicyWindow.hitRatio = (ratio > 1.0 || g_persist_a == 2) ? 1.0f : (float)ratio;
    attr[1].val.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
    attr[1].val.accessPolicyWindow.missProp = g_persist_a == 1 ? cudaAccessPropertyStreaming : cudaAccessPropertyNormal;
    cfg.numAttrs = 2;
  }
  SAEB_CHECK_CUDA(cudaLaunchKernelEx(&cfg, kern, ta, tb, largs));
  return 0;
}

template <int AP, int BP, int PAIR>
static int launch_slots(const CUtensorMap& ta, const CUtensorMap& tb, const EncodeArgs& args, int grid, int cap,
                        cudaStream_t stream) {
  switch (cap) {
    case 256: return launch_cfg<AP, BP, PAIR, 8>(ta, tb, args, grid, stream);
    case 512: return launch_cfg<AP, BP, PAIR, 16>(ta, tb, args, grid, stream);
    case 1024: return launch_cfg<AP, BP, PAIR, 32>(ta, tb, args, grid, stream);
  }
  set_error("unsupported candidate capacity %d", cap);
  return -1;
}

template <int PAIR>
static int launch_planes(int ap, int bp, const CUtensorMap& ta, const CUtensorMap& tb, const EncodeArgs& args,
                         int grid, int cap, cudaStream_t stream) {
  if (ap == 1 && bp == 1) return launch_slots<1, 1, PAIR>(ta, tb, args, grid, cap, stream);
  if (ap == 1 && bp == 2) return launch_slots<1, 2, PAIR>(ta, tb, args, grid, cap, stream);
  if (ap == 2 && bp == 1) return launch_slots<2, 1, PAIR>(ta, tb, args, grid, cap, stream);
  if (ap == 2 && bp == 2) return launch_slots<2, 2, PAIR>(ta, tb, args, grid, cap, stream);
  set_error("unsupported plane counts AP=%d BP=%d", ap, bp);
  return -1;
}

// clusters of two CTA pairs (single-pass modes only: one activation plane, one weight plane)
// Measured on B200 (round 2, profiles/r02_cluster4.md): bit-identical results, but the device offers this kernel only
// 33 resident clusters of four CTAs (132 of 148 SMs; GPC geometry) and the per-round time drops by < 2 % -- the L2
// slices already merge the simultaneous unicast reads of up to ~4 CTAs, so multicast at cluster size 4 removes no L2
// traffic (B300_MICROARCH: "at csz <= 4, MC ~ UC").  Kept as an option, off by default.
static thread_local int g_cluster4 = 0;          // 0: never, 1: when every token tile of the launch gets a resident cluster, 2: always
static int g_max_clusters4 = -1;    // resident clusters of 4 CTAs the device offers this kernel (queried once)
int set_cluster4(int v) {
  if (v < 0 || v > 2) {
    set_error("cluster4 must be 0, 1 or 2");
    return -1;
  }
  g_cluster4 = v;
  return 0;
}
static int max_clusters4() {
  if (g_max_clusters4 < 0) {
    CUtensorMap dummy_a = {}, dummy_b = {};
    EncodeArgs dummy = {};
    int n = 0;
    launch_cfg<1, 1, 2, 8, 2>(dummy_a, dummy_b, dummy, 4, nullptr, &n);
    g_max_clusters4 = n;
  }
  return g_max_clusters4;
}
long long query_max_clusters4() { return max_clusters4(); }
static int launch_cluster4(const CUtensorMap& ta, const CUtensorMap& tb, const EncodeArgs& args, int grid, int cap,
                           cudaStream_t stream) {
  switch (cap) {
    case 256: return launch_cfg<1, 1, 2, 8, 2>(ta, tb, args, grid, stream);
    case 512: return launch_cfg<1, 1, 2, 16, 2>(ta, tb, args, grid, stream);
  }
  set_error("unsupported candidate capacity %d for the 4-CTA cluster kernel", cap);
  return -1;
}

// x_planes: [ap][T][ld_x] bf16 (ap==1: the caller's bf16 activations in place)
// w_planes: [bp][N][d] bf16; bias: folded bias [N]
// Phase 1: the fused GEMM launches (one per row chunk) -> candidate lists in the workspace (and/or dense output)
int encode_gemm_launch(const void* x_planes, int ap, long long T, long long ld_x, long long x_plane_stride,
                       const void* w_planes, int bp, long long ld_w, const float* bias, long long d, long long N, int k,
                       long long clamp_feature, float clamp_value, bool do_topk, float* dense_out, long long ld_dense,
                       void* workspace, size_t workspace_bytes, int pass_mask, int operand_fmt, const float* row_scale,
                       const float* w_unscale, cudaStream_t stream) {
  SAEB_REQUIRE(T > 0 && d > 0 && N > 0, "empty problem T=%lld d=%lld N=%lld", T, d, N);
  SAEB_REQUIRE(ld_w % 8 == 0 && ld_x % 8 == 0, "ld_w and ld_x must be multiples of 8 (16-byte TMA strides)");
  SAEB_REQUIRE(k >= 1 && k <= 512 && k <= N, "k=%d out of range (1..min(512,N))", k);
  SAEB_REQUIRE(T < (1ll << 31) && N < (1ll << 31), "T/N too large");
  SAEB_REQUIRE((reinterpret_cast<uintptr_t>(x_planes) & 15) == 0 && (reinterpret_cast<uintptr_t>(w_planes) & 15) == 0,
               "x / packed weights must be 16-byte aligned");
  const int pair = default_pair();
  static thread_local EncodePlan plan;
  SAEB_REQUIRE(make_plan(plan, T, N, k, pair), "too many row chunks");
  if (do_topk)
    SAEB_REQUIRE(workspace != nullptr && workspace_bytes >= plan.total_bytes,
                 "workspace too small: have %zu need %zu", workspace_bytes, plan.total_bytes);
  if (g_persist_a && g_persist_bytes == 0) set_persist_a(1);

  CUtensorMap tb;
  int rc = make_map(&tb, w_planes, d, N, bp, ld_w * 2, N * ld_w * 2, BN / pair);
  if (rc) return rc;
  if (g_profile) cudaEventRecord(g_ev0, stream);
  for (int ci = 0; ci < plan.n_chunks; ++ci) {
    const ChunkPlan& c = plan.chunks[ci];
    const uint8_t* xa = reinterpret_cast<const uint8_t*>(x_planes) + (size_t)c.t0 * ld_x * 2;
    // two CTA pairs per cluster sharing the activation tile: needs S == 2 splits of equal length and (unless forced)
    // a resident cluster for every token tile of the launch
    const bool cl2 = pair == 2 && ap == 1 && bp == 1 && c.S == 2 && plan.num_n_tiles % 2 == 0 && plan.cap <= 512 &&
                     g_cluster4 > 0 && g_reserve_sms == 0 && (g_cluster4 == 2 || max_clusters4() >= c.num_m_tiles);
    CUtensorMap ta;
    rc = make_map(&ta, xa, d, c.rows, ap, ld_x * 2, x_plane_stride * 2, cl2 ? BM / 2 : BM);
    if (rc) return rc;
    EncodeArgs args;
    args.T = (int)c.rows; args.d = (int)d; args.N = (int)N; args.k = k;
    args.num_m_tiles = c.num_m_tiles;
    args.num_n_tiles = plan.num_n_tiles;
    args.S = c.S;
    args.num_k_blocks = (int)((d + BK - 1) / BK);
    args.pass_mask = pass_mask;
    args.clamp_col = (int)clamp_feature;
    args.clamp_val = clamp_value;
    args.dbg = g_dbg;
    args.prefetch_b = g_prefetch_b;
    args.stats = g_stats;
    args.hint_a = (g_l2_hints & 1) ? L2_EVICT_LAST : L2_EVICT_NORMAL;
    args.hint_b = (g_l2_hints & 2) ? L2_EVICT_FIRST : L2_EVICT_NORMAL;
    // 0 = fp16 operands, 1 = bf16.  Mixed formats (A = bf16 rows in place against the fp16 weight plane, which would
    // save the fp16 activation plane) are not an option: tcgen05.mma kind::f16 with a_format != b_format raises
    // "illegal instruction" on sm_100a (tried in round 2, profiles/r02y_pytest_gpu.log).
    args.idesc = make_idesc_f16(BM * pair, BN, operand_fmt, operand_fmt);
    args.row_scale = row_scale ? row_scale + c.t0 : nullptr;
    args.w_unscale = w_unscale;
    args.bias = bias;
    uint8_t* ws = reinterpret_cast<uint8_t*>(workspace);
    args.cand_cnt = do_topk ? reinterpret_cast<int*>(ws + c.cnt_off) : nullptr;
    args.cand = do_topk ? reinterpret_cast<uint2*>(ws + c.cand_off) : nullptr;
    args.dense_out = dense_out ? dense_out + (size_t)c.t0 * ld_dense : nullptr;
    args.ld_dense = ld_dense;
    g_window.base = const_cast<uint8_t*>(xa);
    g_window.bytes = (plan.n_chunks > 1 || c.rows * ld_x * 2 <= (long long)g_persist_bytes + (8 << 20))
                         ? (size_t)c.rows * (size_t)ld_x * 2 : 0;   // plane 0 of this chunk
    if (cl2) {
      int clusters = max_clusters4() > 0 && max_clusters4() < c.num_m_tiles ? max_clusters4() : c.num_m_tiles;
      rc = launch_cluster4(ta, tb, args, clusters * 4, plan.cap, stream);
    } else {
      rc = (pair == 2) ? launch_planes<2>(ap, bp, ta, tb, args, c.grid, plan.cap, stream)
                       : launch_planes<1>(ap, bp, ta, tb, args, c.grid, plan.cap, stream);
    }
    if (rc) return rc;
  }
  if (g_profile) {
    cudaEventRecord(g_ev1, stream);
    g_ev_valid = true;
  }
  return 0;
}

// Phase 2: exact top-k per row over the candidate lists of phase 1
// coresident != 0: blocks small enough (<= 20 KB of shared memory, 4 warps) to be scheduled beside a resident GEMM CTA
int encode_merge_launch(long long T, long long N, int k, float* out_vals, long long* out_idx, void* workspace,
                        size_t workspace_bytes, int coresident, cudaStream_t stream) {
  const int pair = default_pair();
  static thread_local EncodePlan plan;
  SAEB_REQUIRE(make_plan(plan, T, N, k, pair), "too many row chunks");
  SAEB_REQUIRE(workspace != nullptr && workspace_bytes >= plan.total_bytes, "merge: workspace too small");
  int kp2 = 2;
  while (kp2 < k) kp2 <<= 1;
  uint8_t* ws = reinterpret_cast<uint8_t*>(workspace);
  for (int ci = 0; ci < plan.n_chunks; ++ci) {
    const ChunkPlan& c = plan.chunks[ci];
    // consecutive chunks with the same split count share one launch (their lists are contiguous per chunk only, so
    // each chunk is launched on its own row range)
    const int max_entries = c.S * plan.cap;
    const size_t per_warp = (size_t)(max_entries + kp2) * sizeof(uint2);
    int wpb = (int)((200 * 1024) / per_warp);
    if (wpb > 8) wpb = 8;
    if (coresident && (size_t)wpb * per_warp > 20 * 1024) wpb = (int)((20 * 1024) / per_warp) > 0 ? (int)((20 * 1024) / per_warp) : 1;
    SAEB_REQUIRE(wpb >= 1, "merge: candidate set too large for shared memory");
    const size_t smem = per_warp * wpb;
    SAEB_CHECK_CUDA(cudaFuncSetAttribute(topk_merge_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int blocks = (int)((c.rows + wpb - 1) / wpb);
    SAEB_CARVEOUT(topk_merge_kernel);
    topk_merge_kernel<<<blocks, wpb * 32, smem, stream>>>(
        reinterpret_cast<const uint2*>(ws + c.cand_off), reinterpret_cast<const int*>(ws + c.cnt_off), (int)c.rows, c.S,
        plan.cap, k, kp2, (int)N, max_entries, out_vals + (size_t)c.t0 * k, out_idx + (size_t)c.t0 * k);
    SAEB_CHECK_CUDA(cudaGetLastError());
  }
  return 0;
}

// Feature-sharded scan: candidate selection + bound lists in one register-resident kernel (scan_select_bounds_kernel).
// Returns 1 (nothing launched) when the row's lists do not fit its registers (S * CAP > 512 or K2 > 128): the caller
// then takes the two-kernel route (encode_merge_launch + candidate_bounds_launch).
int encode_select_bounds_launch(long long T, long long N, int K2, int m1, const float* wnorm, const float* dnorm,
                                const float* xnorm, const float* xdnorm, float c_eps, long long clamp_feature,
                                float* out_vals, long long* out_idx, float* exch, void* workspace,
                                size_t workspace_bytes, cudaStream_t stream) {
  const int pair = default_pair();
  static thread_local EncodePlan plan;
  SAEB_REQUIRE(make_plan(plan, T, N, K2, pair), "too many row chunks");
  SAEB_REQUIRE(workspace != nullptr && workspace_bytes >= plan.total_bytes, "select_bounds: workspace too small");
  SAEB_REQUIRE(m1 >= 1 && m1 <= K2, "select_bounds: need 1 <= m1 <= K2");
  if (K2 > 128) return 1;
  for (int ci = 0; ci < plan.n_chunks; ++ci)
    if (plan.chunks[ci].S * plan.cap > 32 * SSB_VPL) return 1;
  uint8_t* ws = reinterpret_cast<uint8_t*>(workspace);
  const int wpb = SSB_THREADS / 32;
  for (int ci = 0; ci < plan.n_chunks; ++ci) {
    const ChunkPlan& c = plan.chunks[ci];
    const int blocks = (int)((c.rows + wpb - 1) / wpb);
    SAEB_CARVEOUT(scan_select_bounds_kernel);
    scan_select_bounds_kernel<<<blocks, SSB_THREADS, 0, stream>>>(
        reinterpret_cast<const uint2*>(ws + c.cand_off), reinterpret_cast<const int*>(ws + c.cnt_off), (int)c.rows, c.S,
        plan.cap, K2, m1, wnorm, dnorm, xnorm + c.t0, xdnorm + c.t0, c_eps, clamp_feature,
        out_vals + (size_t)c.t0 * K2, out_idx + (size_t)c.t0 * K2, exch + (size_t)c.t0 * 2 * m1);
    SAEB_CHECK_CUDA(cudaGetLastError());
  }
  return 0;
}

// x_planes: [ap][T][ld_x] 16-bit (ap==1: the caller's bf16 activations in place)
// w_planes: [bp][N][ld_w] 16-bit; bias: folded bias [N]
int encode_topk_launch(const void* x_planes, int ap, long long T, long long ld_x, long long x_plane_stride,
                       const void* w_planes, int bp, long long ld_w, const float* bias, long long d, long long N, int k,
                       long long clamp_feature, float clamp_value, float* out_vals, long long* out_idx,
                       float* dense_out, long long ld_dense, void* workspace, size_t workspace_bytes, int pass_mask,
                       int operand_fmt, const float* row_scale, const float* w_unscale, cudaStream_t stream) {
  const bool do_topk = out_vals != nullptr;
  int rc = encode_gemm_launch(x_planes, ap, T, ld_x, x_plane_stride, w_planes, bp, ld_w, bias, d, N, k, clamp_feature,
                              clamp_value, do_topk, dense_out, ld_dense, workspace, workspace_bytes, pass_mask,
                              operand_fmt, row_scale, w_unscale, stream);
  if (rc) return rc;
  if (do_topk) rc = encode_merge_launch(T, N, k, out_vals, out_idx, workspace, workspace_bytes, 0, stream);
  return rc;
}

}  // namespace saeb
