/* saeb200 -- C ABI of the B200-native SAE encode / TopK / sparse-decode / activation-cache engine.
 *
 * Drop-in boundary for the hot path of EvolvingLMMs-Lab/multimodal-sae.  The reference has no FFI: its seam is the
 * Python function pointer `decoder_impl(top_indices, top_acts, W_dec_T)` (sae_auto_interp/sae/utils.py:107-129) and
 * the methods of `sae_auto_interp.sae.Sae` (sae/sae.py:172-247).  Each entry point below names the reference code it
 * replaces; `INTEGRATION.md` shows the ctypes binding a maintainer adds on the reference side.
 *
 * Conventions
 *   - plain pointers and sizes only; every pointer is a DEVICE pointer unless stated otherwise;
 *   - the caller owns every buffer including workspaces (query the size first);
 *   - functions only enqueue work on `stream` (a cudaStream_t passed as void*); no allocation, no host sync;
 *   - return 0 on success, negative on error (-1 bad argument, -2 CUDA error); `saeb_last_error()` returns a
 *     thread-local message; nothing throws across the boundary;
 *   - dtype codes: SAEB_F32 = 0, SAEB_BF16 = 1, SAEB_F16 = 2;
 *   - TopK indices are int64 like `EncoderOutput.top_indices` (sae/sae.py:17-22).
 * There is no CPU fallback: without an sm_100a device the compute entry points fail with -2.
 */
#ifndef SAEB200_H
#define SAEB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SAEB_F32 0
#define SAEB_BF16 1
#define SAEB_F16 2

/* Library version (major*10000 + minor*100 + patch). */
int saeb_version(void);
/* Thread-local description of the last error returned on this thread ("" if none). */
const char* saeb_last_error(void);
/* Number of kernel launches enqueued by this library in this process (bench.py's `gpu_launches` claim). */
long long saeb_launch_count(void);
/* Tuning knobs.  "cta_pair": 2 (default) = CTA pairs with tcgen05 cta_group::2 (256-row MMA tiles), 1 = single-CTA
 * tiles.  Results are identical; only throughput differs.  "profile": see saeb_profile_last_encode_ms.  "splits": feature-range splits per token tile
 * (0 = automatic).  "chunking": 1 (default) = long calls run as one launch per wave of token tiles (keeps the live
 * activation tiles L2-resident).  "persist_a": 1 (default) = pin each launch's activation rows in the persisting part
 * of L2.  "reserve_sms": SMs the persistent GEMM grid leaves free (default 0) so that kernels of another stream --
 * the collectives of the feature-sharded scan -- have somewhere to run while a GEMM launch is in flight.
 * "refine_threads": threads per refinement CTA in feature-sharded calls (default 128; 256 keeps more loads in flight
 * when the kernel runs beside a GEMM launch, where only one CTA fits per SM).  "kth_impl": see
 * saeb_kth_largest_gathered.  "gemm_stages": depth of the GEMM's shared-memory ring (0 = as deep as fits: 6; 5 leaves
 * ~54 KB per SM to gather CTAs running beside it).  "cluster4": 4-CTA clusters sharing the activation tile by TMA
 * multicast (0 = off, default; 1 = when every token tile gets a resident cluster; 2 = always).  "scan_warp": 1
 * (default) = feature-sharded scan calls of the refinement (value_mode 2 with ext_lower and ext_upper) run as the
 * warp-per-token kernel without shared memory, which is scheduled beside a resident GEMM CTA; 0 = CTA per token.
 * "stats": cycle /
 * gather counters for saeb_debug_stats.  "l2_hints", "debug_tiles", "prefetch_b": diagnostics.
 * Options are PER CALLING THREAD (thread-local state, like the error string): two threads driving two streams or two
 * SAEs never see each other's settings, and the profiling events belong to the thread that set "profile". */
int saeb_set_option(const char* name, int value);
/* With option "profile" = 1 the library brackets the fused encode kernel (main kernel only) with CUDA events on the
 * launching stream; this returns the duration in ms of the most recent one (synchronises on it), < 0 if none. */
float saeb_profile_last_encode_ms(void);
/* Diagnostics: a synthetic load of `ctas` 128-thread CTAs without shared memory (they are scheduled beside a resident
 * GEMM CTA), `iters` iterations per thread: mode 0 = dependent FMA chains (issue slots only), 1 = streaming 16-byte
 * reads of buf (DRAM traffic, L2 pollution), 2 = reads confined to the first 32 MB of buf (L2 bandwidth only).
 * tools/probe_overlap.py uses it to separate what slows a GEMM launch that shares its SMs.  sink: one device float. */
int saeb_debug_coload(int mode, int ctas, int64_t iters, const void* buf, size_t bytes, float* sink, void* stream);
/* Diagnostics (option "stats" = 1 zeroes and enables device cycle counters of the fused encode kernel): host array of 8
 * counters summed over CTA pairs = {producer waiting for a free smem stage, MMA issuer waiting for a free TMEM stage,
 * MMA issuer waiting for TMA data, epilogue warp waiting for an accumulator, epilogue compaction time, kernel time,
 * number of CTA pairs summed, rows gathered by the refinement kernels since the option was set}.  Synchronises. */
int saeb_debug_stats(unsigned long long* out8);
/* Device facts used by the launch heuristics: "num_sms", "l2_bytes", "persisting_l2_max_bytes",
 * "access_policy_max_window_bytes", "persisting_l2_in_use_bytes"; < 0 if unknown. */
long long saeb_query(const char* name);

/* ---- one-time weight repack -------------------------------------------------------------------------------
 * Reference parameters: `encoder.weight [N,d]`, `encoder.bias [N]`, `b_dec [d]` fp32 (sae/sae.py:59-66, loaded by
 * Sae.load_from_disk sae/sae.py:126-148).  Produces `planes` bf16 planes of W_enc (1: bf16(W); 2: hi + lo, the
 * parity-grade mode; 3: the "fp16 + refine" layout used by saeb_encode_topk_refine; 4: layout 3 plus the fp16 residual
 * plane used by saeb_refine_candidates_lo) followed by the folded bias b_enc - W_enc b_dec (sae/sae.py:174-175 rewritten as
 * W x + (b_enc - W b_dec)).  Layout of `packed`: [planes][N][d_pad] bf16 with d_pad = d rounded up to 8 (16-byte
 * rows for TMA, zero padded), then [N] fp32 at saeb_packed_bias_offset().  Any d >= 1 is accepted. */
size_t saeb_packed_weights_bytes(int64_t N, int64_t d, int planes);
size_t saeb_packed_bias_offset(int64_t N, int64_t d, int planes);
int saeb_pack_weights(const float* W_enc, const float* b_enc, const float* b_dec, int64_t N, int64_t d, int planes,
                      void* packed, void* stream);

/* ---- fused encode + TopK ------------------------------------------------------------------------------------
 * Replaces Sae.encode = select_topk(pre_acts(x)) (sae/sae.py:172-185) and the cache path's
 * `pre_acts -> torch.topk` (features/cache.py:210-213, :407-412) without materialising the dense [T,N] latents.
 *   x            [T, ld_x] activations.  bf16 with d % 8 == 0 is consumed in place (needs a 16-byte aligned base and
 *                ld_x % 8 == 0); f16 / f32 are split into two bf16 planes, bf16 with d % 8 != 0 is copied to a padded
 *                plane (both in the workspace)
 *   clamp_feature / clamp_value: steering (features/steering.py:113-114): that latent is forced to clamp_value
 *                before TopK; pass -1 to disable
 *   out_vals     [T,k] f32, out_idx [T,k] i64, each row ordered by (value desc, index asc); rows with fewer than k
 *                positive pre-activations are padded with value 0 on distinct unused indices
 *   dense_out    optional [T, ld_dense] f32: relu(pre-activations) for callers that need Sae.pre_acts
 *                (sae/sae.py:172-177); out_vals/out_idx may be NULL when only dense_out is wanted. */
size_t saeb_encode_topk_workspace_bytes(int64_t T, int64_t d, int64_t N, int k, int x_dtype);
int saeb_encode_topk(const void* x, int x_dtype, int64_t T, int64_t ld_x, const void* packed, int planes, int64_t d,
                     int64_t N, int k, int64_t clamp_feature, float clamp_value, float* out_vals, int64_t* out_idx,
                     float* dense_out, int64_t ld_dense, void* workspace, size_t workspace_bytes, void* stream);

/* ---- fused encode + TopK, single tensor-core pass with exact refinement ("fp16 + refine") -------------------------
 * Same result contract as saeb_encode_topk (parity grade) at half the tensor-core work.  `packed` must have been
 * produced by saeb_pack_weights(..., planes = 3): one fp16 plane of W_enc (power-of-two scaled), folded bias,
 * per-feature norms.  The GEMM ranks by approximate values; every candidate that can still belong to the TopK under the
 * rigorous rounding bound  (2^-11 [+2^-11 for fp32 x] + 2^-12) * ||x||_2 * ||w_j||_2  is re-evaluated exactly in fp32
 * against `W_enc` (the [N,d] fp32 parameter itself), so the returned values are fp32-exact and the index set is the
 * fp32 reference's.  `margin` = extra candidates kept per row (0 = max(48, k/2), or the value of option "refine_margin").  Rows whose candidate list could be too short
 * for the bound (never observed) are recomputed by an exact dense fp32 kernel, up to 64 per call; *status_out (device
 * int, may be NULL) receives the number of such rows -- more than 64 means the call must be repeated with a larger
 * margin. */
size_t saeb_encode_topk_refine_workspace_bytes(int64_t T, int64_t d, int64_t N, int k, int margin);
int saeb_encode_topk_refine(const void* x, int x_dtype, int64_t T, int64_t ld_x, const void* packed,
                            const float* W_enc, int64_t d, int64_t N, int k, int margin, int64_t clamp_feature,
                            float clamp_value, float* out_vals, int64_t* out_idx, int32_t* status_out, void* workspace,
                            size_t workspace_bytes, int value_mode, void* stream);

/* The stages of saeb_encode_topk_refine as separate calls, so that a caller can run the tensor-core-bound GEMM of the
 * next token chunk concurrently (on another stream) with the HBM-bound refinement and decode of the previous one:
 *   saeb_prep_activations   whole batch: activations -> power-of-two scaled fp16 plane + per-row scale + norms
 *                           (`prep`, saeb_prep_bytes(T, d) bytes);
 *   saeb_encode_candidates  rows [t0, t0+Tc): single-pass GEMM with fused candidate selection (tensor bound);
 *   saeb_refine_candidates  same rows: candidate merge + exact fp32 re-evaluation + dense fallback (HBM bound); `x`
 *                           points at row t0 of the original activations; ext_lower = NULL, already_merged = 0 unless
 *                           feature sharded (below).  already_merged: 0 = merge the GEMM's candidate lists first,
 *                           1 = saeb_candidate_bounds already left sorted merged candidates in the workspace,
 *                           2 = saeb_candidate_bounds_packed left UNSORTED ones (feature-sharded scan form only).
 * Both row-range calls share a scratch buffer of saeb_candidates_workspace_bytes(Tc, ...) bytes. */
size_t saeb_prep_bytes(int64_t T, int64_t d);
int saeb_prep_activations(const void* x, int x_dtype, int64_t T, int64_t ld_x, int64_t d, void* prep, void* stream);
size_t saeb_candidates_workspace_bytes(int64_t T, int64_t d, int64_t N, int k, int margin);
int saeb_encode_candidates(const void* prep, int64_t T_total, int64_t t0, int64_t Tc, const void* packed, int64_t d,
                           int64_t N, int k, int margin, int64_t clamp_feature, float clamp_value, void* workspace,
                           size_t workspace_bytes, void* stream);
int saeb_refine_candidates(const void* x, int x_dtype, int64_t ld_x, const void* prep, int64_t T_total, int64_t t0,
                           int64_t Tc, const void* packed, const float* W_enc, int64_t d, int64_t N, int k, int margin,
                           int64_t clamp_feature, float clamp_value, const float* ext_lower, const float* ext_upper,
                           const float* feat_thr, int already_merged, float* out_vals, float* out_member,
                           int64_t* out_idx, int32_t* status_out, void* workspace, size_t workspace_bytes,
                           int max_ctas, int value_mode, void* stream);
/* value_mode (saeb_encode_topk_refine, saeb_refine_candidates[_lo]): which outputs are re-evaluated exactly.
 *   0  every member of the TopK carries its exact fp32 value (reference grade, ~3e-7 relative);
 *   1  "boundary only": the index SET is still decided rigorously -- every candidate whose error interval
 *      [a_j - eps_j, a_j + eps_j] straddles the k-th boundary is re-evaluated exactly -- but members that are in the
 *      TopK whatever their exact value keep the tensor-core value a_j (W_enc rounded to fp16: |a_j - exact| <= eps_j
 *      rigorously, ~5e-5 relative in practice at d = 4096; a member whose bound exceeds 2^-7 of its value is
 *      re-evaluated too).  About 10 instead of 70 gathered rows per token at k = 64, N = 131072.  Ignored (treated as
 *      0) for feature-sharded calls (ext_lower != NULL).
 *   2  "scan": for the top-activation scan.  Membership is decided as in mode 1; a certain member is re-evaluated
 *      EXACTLY only if it can still enter its feature's top-n list (a_j + eps_j >= feat_thr[feature], feat_thr [N]
 *      = the scan's current n-th best per feature, may be NULL = always) and is reported with value 0 otherwise;
 *      every boundary candidate is re-evaluated exactly.  Every value that reaches a list is an exact fp32 value.
 *      Feature-sharded form: ext_lower [Tc] AND ext_upper [Tc] (an upper bound of the token's global (k+1)-th largest
 *      upper bound, from the all-gathered ub lists of saeb_candidate_bounds) are given; the shard then also writes
 *      out_member [Tc,k]: 3e38 for certain members, the exact value for boundary candidates, 0 for padding.  The
 *      k-th largest of all shards' member values is the token's membership threshold (saeb_kth_largest_gathered);
 *      pass it as tok_thr and out_member as `member` to saeb_scan_pool.
 *   ext_upper / feat_thr / out_member are NULL outside mode 2.
 * max_ctas (saeb_refine_candidates[_lo], saeb_decode): 0 = one CTA per token.  > 0 = a persistent grid of at most that
 * many CTAs walks the tokens, and every helper launch of the call uses blocks small enough (<= 256 threads, <= 21 KB
 * of shared memory) to be scheduled on an SM that already hosts a CTA of the fused GEMM: with max_ctas = (1..2) x
 * number of SMs the HBM-bound gathers of chunk c run INSIDE the tensor-core-bound GEMM launches of chunk c+1 instead
 * of queueing behind them (saeb200.overlap.OverlappedForward).  Results do not depend on max_ctas. */
/* "fp16 hi + lo" refinement (packed mode 4 = saeb_pack_weights(..., planes = 4): the mode-3 blob followed by an fp16
 * plane of the residual W - fp16(W), scaled by 2^11; every mode-3 entry point accepts a mode-4 blob unchanged).
 * saeb_refine_candidates_lo is saeb_refine_candidates with the exact re-evaluation replaced by a correction: the
 * approximate value x . W_hi + bias already came out of the tensor cores, only x . W_lo is added -- the gather reads
 * 2*d instead of 4*d bytes per candidate.  Exact in the operands for bf16 / fp16 activations (fp32 activations are
 * rounded on their way into the tensor cores, so they silently take the exact route); the corrected values carry the
 * tensor cores' fp32 accumulation noise (~1e-6 relative, the grade of any fp32 GEMM) instead of 3e-7.  The flagged-row
 * fallback is the same exact dense path (needs W_enc). */
int saeb_refine_candidates_lo(const void* x, int x_dtype, int64_t ld_x, const void* prep, int64_t T_total, int64_t t0,
                              int64_t Tc, const void* packed4, const float* W_enc, int64_t d, int64_t N, int k,
                              int margin, int64_t clamp_feature, float clamp_value, const float* ext_lower,
                              const float* ext_upper, const float* feat_thr, int already_merged, float* out_vals,
                              float* out_member, int64_t* out_idx, int32_t* status_out, void* workspace,
                              size_t workspace_bytes, int max_ctas, int value_mode, void* stream);
/* Feature-sharded use (every GPU holds N/R features, sees all tokens): after saeb_encode_candidates,
 * saeb_candidate_bounds merges this shard's candidates and writes, per token, its k largest LOWER bounds
 * a_j - eps_j (descending, lb_out [Tc,k]).  All-gather them, take the per-token k-th largest (saeb_kth_of_gathered):
 * that is a lower bound of the token's GLOBAL k-th activation; pass it as `ext_lower` (with already_merged = 1) and
 * the shard only re-evaluates candidates that can still be in the global TopK (about k/R + a few per token instead of
 * k + 25).  ub_out (optional, [Tc,k]): the k largest UPPER bounds a_j + eps_j as well, for value_mode 2.
 * coresident != 0: launch shapes that fit beside a resident GEMM CTA (the call then can run on another stream INSIDE
 * the next chunk's GEMM launches instead of between them). */
int saeb_candidate_bounds(const void* prep, int64_t T_total, int64_t t0, int64_t Tc, const void* packed, int x_dtype,
                          int64_t d, int64_t N, int k, int margin, int64_t clamp_feature, float* lb_out, float* ub_out,
                          void* workspace, size_t workspace_bytes, int coresident, void* stream);

/* saeb_candidate_bounds as ONE register-resident kernel for the pipelined feature-sharded scan: selects the row's K2
 * best candidates and writes bounds_out [Tc, 2*m1] = the m1 largest lower bounds | the m1 largest upper bounds of the
 * row (zero padded, NOT sorted -- saeb_gathered_bounds does not need an order), the payload of exchange 1.  The merged
 * candidates are left in the workspace UNSORTED: pass already_merged = 2 to saeb_refine_candidates, which then
 * insists on the feature-sharded scan form (value_mode 2, ext_lower, ext_upper, out_member) whose warp-per-token
 * kernel is order independent.  No passes over shared memory, 128-thread CTAs: runs beside a resident GEMM CTA.
 * Returns 1 (nothing launched, no error) when the shape does not fit the kernel's registers (more than 512 list
 * entries per row, or K2 > 128): use saeb_candidate_bounds then. */
int saeb_candidate_bounds_packed(const void* prep, int64_t T_total, int64_t t0, int64_t Tc, const void* packed,
                                 int x_dtype, int64_t d, int64_t N, int k, int margin, int64_t clamp_feature, int m1,
                                 float* bounds_out, void* workspace, size_t workspace_bytes, void* stream);

/* TopK of dense non-negative rows, (value desc, index asc): Sae.select_topk (sae/sae.py:179-181) for callers that hold
 * a dense [T, ld] latent tensor. */
int saeb_dense_topk(const float* dense, int64_t T, int64_t ld, int64_t N, int k, float* out_vals, int64_t* out_idx,
                    void* stream);

/* ---- sparse decode ------------------------------------------------------------------------------------------
 * Replaces decoder_impl (sae/utils.py:107-129: triton_sparse_dense_matmul sae/kernels.py:178-284, or eager_decode
 * sae/utils.py:108-111) + `+ b_dec` (sae/sae.py:191) + the steering hook's output cast (features/steering.py:116-118).
 *   out[t,:] = sum_j vals[t,j] * W_dec[idx[t,j],:] + b_dec        (fp32 accumulate in j order, zeros skipped)
 *   W_dec [N,d] contiguous (the reference passes the transposed view W_dec.mT; pass the parameter itself),
 *   w_dtype SAEB_F32 (parity grade) or SAEB_BF16; b_dec may be NULL; out_dtype any of the three codes.
 *   If sq_err != NULL, x (same shape as out) is read and sum((out - x)^2) is ADDED to *sq_err (double) -- the
 *   numerator of the FVU (sae/sae.py:201,229).  err_flag (int, may be NULL) is set to 1 if an index is out of
 *   range (tl.device_assert at sae/kernels.py:276).  max_ctas: see saeb_refine_candidates. */
int saeb_decode(const int64_t* idx, const float* vals, int64_t T, int k, const void* W_dec, int w_dtype, int64_t d,
                int64_t N, const float* b_dec, void* out, int out_dtype, int64_t ld_out, const void* x, int x_dtype,
                int64_t ld_x, double* sq_err, int* err_flag, int max_ctas, void* stream);

/* ---- cache reader on the device ----------------------------------------------------------------------------
 * Replaces the per-feature densify + pool of the example constructors (features/constructors.py:11-47 `_to_dense` +
 * max_pool1d for text windows, :104-114 avg_pool1d over the base image tokens) for ALL features of a split file at
 * once.  Input: the file's triples grouped by feature id (stable: file order kept inside a feature) as parallel arrays
 * feature[nnz], window_key[nnz] (text: row * n_win + pos / ctx_len; image: row if pos < n_base; -1 = outside every
 * window), activations[nnz].  For the first entry of every (feature, window) run: head = 1 and score = max of the run
 * (mode 0) or (sequential fp32 sum of the run) / divisor (mode 1, divisor = n_base); other entries: head = 0. */
int saeb_coo_window_scores(const int64_t* feature, const int64_t* window_key, const float* activations, int64_t nnz,
                           int mode, float divisor, float* score, int* head, void* stream);

/* ---- probing queries -----------------------------------------------------------------------------------------
 * Replaces the hook body of the reference's probing tool (tools/probe_activations.py:109-126):
 *   latents = sae.pre_acts(hidden); top = latents.mean(0).topk(k).indices; maps = latents[:, top]
 * saeb_column_sums ADDS the column sums of a dense latent chunk [T, ld] (as written by saeb_encode_topk's dense
 * output) to colsum [N] (double; zero it first) -- chunks of tokens can be streamed through one scratch buffer.
 * saeb_feature_maps evaluates out[j][t] = relu((x_t - b_dec) . W_enc[features[j]] + b_enc[features[j]]) exactly in
 * fp32 for a few selected features (out [n_features, T]); err_flag is set to 1 on an out-of-range feature id. */
int saeb_column_sums(const float* dense, int64_t T, int64_t ld, int64_t N, double* colsum, void* stream);
int saeb_feature_maps(const void* x, int x_dtype, int64_t T, int64_t ld_x, const float* W_enc, const float* b_enc,
                      const float* b_dec, int64_t d, int64_t N, const int64_t* features, int n_features, float* out,
                      int* err_flag, void* stream);

/* ---- backward of the sparse decode -------------------------------------------------------------------------
 * The decoder seam is a torch.autograd.Function in the reference (TritonDecoder, sae/kernels.py:403-429); these are
 * its two backward products, fp32:
 *   saeb_decode_backward_acts    d_vals[t,j] = grad_out[t,:] . W_dec[idx[t,j],:]   (triton_dense_dense_sparseout_matmul,
 *                                sae/kernels.py:287-400; computed for zero activations too)
 *   saeb_decode_backward_weight  dW_dec[n,:] += sum over (t,j) with idx[t,j] == n of vals[t,j] * grad_out[t,:]
 *                                (triton_sparse_transpose_dense_matmul, sae/kernels.py:10-175; zero values skipped).
 *                                ACCUMULATES into dW_dec [N,d] with fp32 atomics (like the reference's tl.atomic_add,
 *                                the summation order is not fixed): zero it before the first call.
 * grad_out [T, ld_g] f32; err_flag as in saeb_decode. */
int saeb_decode_backward_acts(const float* grad_out, int64_t ld_g, const int64_t* idx, int64_t T, int k,
                              const float* W_dec, int64_t d, int64_t N, float* d_vals, int* err_flag, void* stream);
int saeb_decode_backward_weight(const float* grad_out, int64_t ld_g, const int64_t* idx, const float* vals, int64_t T,
                                int k, int64_t d, int64_t N, float* dW_dec, int* err_flag, void* stream);

/* FVU denominator  sum((x - x.mean(0))^2)  (sae/sae.py:204); scratch = 2*d doubles. */
int saeb_total_variance(const void* x, int x_dtype, int64_t T, int64_t d, int64_t ld_x, double* scratch, double* out,
                        void* stream);

/* ---- activation-cache extraction ----------------------------------------------------------------------------
 * Replaces `zeros_like + scatter_` (features/cache.py:215-217) + Cache.get_nonzeros (features/cache.py:73-92).
 * Input is TopK output of T = batch*seq_len tokens; output is the reference's `locations [nnz,3]` int64
 * (row_offset + t / seq_len, t % seq_len, feature) in torch.nonzero order and `activations [nnz]` f32, keeping
 * entries with |v| > threshold (1e-5 in the reference) whose feature bit is set in `filter_bitmap` (N bits packed in
 * uint32 words, NULL = keep all; replaces torch.isin at features/cache.py:89-92).  locations/activations must hold
 * T*k entries; *nnz_out (device int64) receives the count. */
size_t saeb_coo_workspace_bytes(int64_t T);
int saeb_coo_extract(const float* vals, const int64_t* idx, int64_t T, int k, float threshold,
                     const uint32_t* filter_bitmap, int64_t seq_len, int64_t row_offset, int64_t* locations,
                     float* activations, int64_t* nnz_out, void* workspace, size_t workspace_bytes, void* stream);
/* Same extraction, but the triples are APPENDED behind *cursor (device int64, in/out) in a caller-owned arena of
 * `capacity` entries (locations [capacity,3], activations [capacity]) and the cursor advances on the device: the cache
 * of a whole run (reference Cache.add, features/cache.py:42-57, which moves every batch to the host) accumulates in HBM
 * with no host round trip per batch; one device-to-host copy at Cache.save().  Entries that do not fit are dropped and
 * *overflow_flag (device int, may be NULL) is set.  Workspace: saeb_coo_workspace_bytes(T) + 256 bytes. */
int saeb_coo_append(const float* vals, const int64_t* idx, int64_t T, int k, float threshold,
                    const uint32_t* filter_bitmap, int64_t seq_len, int64_t row_offset, int64_t* locations,
                    float* activations, int64_t capacity, int64_t* cursor, int* overflow_flag, void* workspace,
                    size_t workspace_bytes, void* stream);

/* ---- per-feature top-activation scan ------------------------------------------------------------------------
 * Replaces, for all features at once, TensorBuffer.__getitem__ (features/loader.py:74-90) +
 * pool_max_activation_windows (features/constructors.py:11-85): max-pool of the TopK-masked activations over
 * windows of ctx_len tokens and the n_top best windows per feature, ordered (score desc, window asc).
 * saeb_scan_pool processes one chunk of tokens (at most bucket_cap windows): only features in [feat_lo, feat_hi)
 * (this GPU's shard) are kept, local index f - feat_lo.  tok_thr (NULL or [T]) is the per-token global k-th value
 * used under feature sharding: an entry is dropped if its activation -- or, when `member` ([T,k], optional) is given,
 * member[t][j] instead (value_mode 2 of the refinement) -- is below tok_thr[t].  saeb_scan_merge folds the buckets into the sorted per-feature lists
 * top_vals [F,n_top] f32 / top_win [F,n_top] i64 (-1 = empty), refreshes feat_thr and clears the bucket counters.
 * Initialise feat_thr to the activation threshold (1e-5), top_win to -1, bucket_cnt to 0. */
int saeb_scan_pool(const float* vals, const int64_t* idx, int64_t T, int k, int ctx_len, float threshold,
                   int64_t feat_lo, int64_t feat_hi, int64_t window_base, const float* tok_thr,
                   const float* member, const float* feat_thr, void* bucket, int* bucket_cnt, int bucket_cap,
                   int* overflow_flag, void* stream);
int saeb_scan_merge(void* bucket, int* bucket_cnt, int bucket_cap, int64_t F, int n_top, float base_threshold,
                    float* top_vals, int64_t* top_win, float* feat_thr, void* stream);
/* saeb_scan_pool with its hash tables in global scratch instead of shared memory (same appended entries): persistent
 * 128-thread CTAs without shared memory, which are scheduled beside a resident CTA of the fused GEMM -- the form the
 * pipelined scan uses, where the list update of chunk c runs inside the GEMM launches of chunk c+1.  The scratch
 * (saeb_scan_pool_workspace_bytes) is initialised ONCE with saeb_scan_pool_init and left initialised by every call. */
size_t saeb_scan_pool_workspace_bytes(int k, int ctx_len);
int saeb_scan_pool_init(void* workspace, size_t workspace_bytes, int k, int ctx_len, void* stream);
int saeb_scan_pool_ws(const float* vals, const int64_t* idx, int64_t T, int k, int ctx_len, float threshold,
                      int64_t feat_lo, int64_t feat_hi, int64_t window_base, const float* tok_thr, const float* member,
                      const float* feat_thr, void* bucket, int* bucket_cnt, int bucket_cap, int* overflow_flag,
                      void* workspace, size_t workspace_bytes, void* stream);
/* Image form of the scan (replaces, for all features at once, pool_max_activations_windows_image,
 * features/constructors.py:88-148): the score of (feature, image) is the MEAN of the feature's TopK-masked activations
 * over the first n_base positions of the image's token row (avg_pool1d over the 576 base image tokens, :109-114).
 * saeb_image_pool processes the TopK output of n_images rows of tokens_per_image tokens each (at most bucket_cap images
 * per call) and appends (score, image_base + i) to the same per-feature buckets as saeb_scan_pool; saeb_scan_merge
 * then keeps the n_top best images per feature, ordered (score desc, image id asc) -- ask for max_examples + 50 and
 * drop repeated dataset ids on the host like the reference does (:117-135).  Sums are accumulated in 32.32 fixed
 * point (order independent); scores agree with the reference's fp32 mean to ~1e-7 relative.  The scratch
 * (saeb_image_pool_workspace_bytes) must be initialised ONCE with saeb_image_pool_init and is left initialised by
 * every call. */
size_t saeb_image_pool_workspace_bytes(int k, int n_base, int64_t F);
int saeb_image_pool_init(void* workspace, size_t workspace_bytes, int k, int n_base, int64_t F, void* stream);
int saeb_image_pool(const float* vals, const int64_t* idx, int64_t n_images, int64_t tokens_per_image, int k,
                    int n_base, float threshold, int64_t feat_lo, int64_t feat_hi, int64_t image_base,
                    const float* tok_thr, const float* feat_thr, void* bucket, int* bucket_cnt, int bucket_cap,
                    int* overflow_flag, void* workspace, size_t workspace_bytes, void* stream);
/* Per-token kth-largest of R gathered per-shard value lists, gathered [R][T][m] f32 (after the all-gather of each
 * shard's m best values per token): tok_thr[t] = the kth-largest of the R*m values of token t, values
 * <= 0 count as 0; 1 <= kth <= R*m.  The feature-sharded scan uses it twice per token chunk: on the gathered lower
 * bounds (m may be smaller than k: the kth-largest of ANY >= kth lower bounds of distinct latents is a lower bound of
 * the global k-th activation) and on the gathered exact local TopK values (m = k).  saeb_kth_of_gathered is the
 * m = kth = k form.  Option "kth_impl" = 0 selects the first, memory-resident version of the kernel (diagnostics). */
int saeb_kth_largest_gathered(const float* gathered, int R, int64_t T, int m, int kth, float* tok_thr, void* stream);
int saeb_kth_of_gathered(const float* gathered, int R, int64_t T, int k, float* tok_thr, void* stream);
/* Exchange 1 of the feature-sharded scan in one pass.  gathered [R][T][2*m1] f32: per shard and token the m1 largest
 * lower bounds followed by the m1 largest upper bounds of saeb_candidate_bounds (both descending).  Writes
 *   ext_lower[t] = k-th largest of the R*m1 lower bounds (0 if R*m1 < k)             -> `ext_lower` of the refinement
 *   ext_upper[t] = max((k+1)-th largest of the R*m1 upper bounds (0 if R*m1 < k+1),
 *                      max over shards of the SMALLEST upper bound the shard sent)   -> `ext_upper` (value_mode 2)
 * (whatever a shard did not send is no larger than the smallest bound it sent; a shard's lists need not be sorted).
 * R*m1 <= 2048. */
int saeb_gathered_bounds(const float* gathered, int R, int64_t T, int m1, int k, float* ext_lower, float* ext_upper,
                         void* stream);

/* ---- peer-memory all-gather for the feature-sharded scan (optional replacement of the two per-chunk NCCL
 * all-gathers) -------------------------------------------------------------------------------------
 * Every rank holds a SYMMETRIC buffer (same size and offsets on all R ranks of one NVLink/NVSwitch box; mapped into
 * each other's address space, e.g. by torch.distributed._symmetric_memory) laid out by the caller as
 *   [flags: uint32 [channels][R], zero-initialised] ... [gathered region: R slabs of `bytes` each] ...
 * saeb_push_gather copies this rank's slab `src` (device, 16-byte aligned, bytes % 16 == 0) into slab `self_rank` of the
 * region at `region_offset` in EVERY rank's buffer -- 16-byte stores to the mapped peer pointers
 * `peer_bases_dev` (DEVICE array of R base pointers, own buffer included), or one multimem.st per vector through
 * `multicast_base` (the buffer's NVSwitch multicast mapping; NULL = unicast stores) -- then publishes sequence number
 * `seq` for (channel, self_rank) on every rank (system-scope release) and waits until all R ranks' numbers for this
 * channel have reached `seq` in the local flags (acquire).  When the kernel ends the local region [R][bytes] is
 * complete; work enqueued later on `stream` may read it.  All ranks must call with the same channel / seq / bytes /
 * offsets; `seq` must grow by one per call on a channel (start at 1); `counter` is a zero-initialised device int per
 * channel (local memory, restored to 0 by the kernel).  A rank that never arrives makes the kernel trap after 20 s. */
int saeb_push_gather(const void* src, size_t bytes, void* const* peer_bases_dev, int R, int self_rank,
                     size_t region_offset, void* multicast_base, size_t flags_offset, int channel, uint32_t seq,
                     int* counter, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SAEB200_H */
