"""Multi-GPU forms of the path, one process per GPU (torchrun), `torch.distributed` for the plumbing.

* token-parallel forward: tokens are independent, so ranks take disjoint token slices with the full SAE replicated and
  no data-path collective (the reference's own strategy, `dataset.shard`, launch/cache/cache_image.py:60-71);
* feature-sharded top-activation scan (the north-star layout): rank r owns SAE features [r*N/R, (r+1)*N/R); every rank
  sees every token; per-feature top lists are rank-local and disjoint, so the job ends with ONE all-gather of the
  per-shard lists.  Because the cache keeps a latent only if it is in the token's *global* top-k
  (features/cache.py:210-218), exact mode additionally all-gathers each chunk's local top-k VALUES ([Tc, k] fp32 per
  rank, 256 B/token/rank) and derives the per-token global k-th value; `exact=False` skips that exchange and ranks by
  the shard-local top-k instead (not reference semantics, reported separately).

The choreography is written against a small `ops` object so that the same code runs on CUDA (`EngineOps`, NCCL) and in
the CPU `gloo` tests (where `tests/` inject oracle-backed ops).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Iterable, Optional, Tuple

import os

import torch
import torch.distributed as dist

LOOKAHEAD_MIN_WORLD = 1 << 30   # see sharded_scan: the lookahead schedule never beat the two-stream one so far


def scan_schedule(world: int) -> str:
    """schedule sharded_scan uses for CUDA ops (SAEB_SCAN_SCHEDULE = lookahead | streams overrides, diagnostics)"""
    return os.environ.get("SAEB_SCAN_SCHEDULE") or ("lookahead" if world >= LOOKAHEAD_MIN_WORLD else "streams")


def bounds_width(k: int, k_local: int, world: int) -> int:
    """Columns of the per-token lower-bound lists a shard contributes to exchange 1.

    The k-th largest of ANY set of >= k lower bounds of distinct latents is a valid lower bound of the token's global
    k-th activation, so a shard need not send all k of its bounds: with the global TopK spread over `world` shards a
    shard holds about k / world of them, and 2 k / world + 8 columns lose nothing in practice (a shard that does hold
    more only makes the bound a little looser, i.e. a few more exact re-evaluations -- the result stays exact).  The
    width never drops below ceil(k / world), so the union always has k entries.  SAEB_SCAN_BOUNDS_WIDTH overrides."""
    env = os.environ.get("SAEB_SCAN_BOUNDS_WIDTH")
    lo = -(-k // max(world, 1))
    m = int(env) if env else max(16, 2 * lo + 8)
    return max(1, min(k_local, max(m, lo)))


def shard_range(num_latents: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous, balanced feature range of `rank` (sizes differ by at most one)."""
    base, rem = divmod(num_latents, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def token_slice(num_tokens: int, world: int, rank: int) -> Tuple[int, int]:
    base, rem = divmod(num_tokens, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


class EngineOps:
    """CUDA implementation of the per-rank steps of the feature-sharded scan (thin calls into the C ABI).

    Interface used by `sharded_scan` (the CPU gloo tests inject oracle-backed objects with the same methods):
      local_bounds(x, k, slot)        -> (lb, ub), both [Tc, k] descending: lower / upper bounds a_j -/+ eps_j of the
                                         shard's k best candidates of the chunk
      local_topk(ext_L, ext_U, slot)  -> (vals, member, ids), all [Tc, k]: the shard's entries that can be in the token's
                                         GLOBAL TopK given the bounds of the union (ext_L: lower bound of the global k-th
                                         value, ext_U: upper bound of the (k+1)-th); member = 3e38 for certain members,
                                         the exact value for undecided ones; ids as the ops like (scan_update takes them)
      kth_of_gathered, scan_update, scan_finalize
      optional, used when present: local_gemm + local_bounds_finish(slot, coresident, pack_m1) (GEMM and bounds as
      separate steps, bounds as one unsorted payload [Tc, 2*m1]), gathered_bounds (fused bounds of the union),
      push_gather (own transport of the two exchanges), local_prep, begin_pipeline / end_pipeline
    `slot` (0/1) selects one of two private scratch sets, so that the GEMM of chunk c+1 can be in flight on
    `stream_gemm` while chunk c is exchanged / refined / scanned on `stream_aux` (`pipelined = True`).
    """

    pipelined = True

    def __init__(self, W_enc_shard, b_enc_shard, b_dec, feat_lo, feat_hi, n_top, ctx_len, device, planes=3,
                 bucket_cap=256, aux_priority=None, image_n_base=None):
        """`image_n_base` (optional): the IMAGE form of the scan -- `ctx_len` is then the number of tokens per image row,
        a feature's score for an image the mean of its TopK-masked activations over the first `image_n_base` positions
        (engine.TopImageScan; reference pool_max_activations_windows_image, features/constructors.py:88-148).  A mean
        needs every member's exact value, so the refinement runs with every member exact instead of the scan mode."""
        from . import _capi, engine

        self.engine, self._capi = engine, _capi
        self.enc = engine.PackedEncoder.pack(W_enc_shard, b_enc_shard, b_dec, planes)
        self.feat_lo, self.feat_hi = feat_lo, feat_hi
        if image_n_base is None:
            self.scan = engine.TopActivationScan(feat_lo, feat_hi, n_top, ctx_len, device, bucket_cap=bucket_cap)
        else:
            self.scan = engine.TopImageScan(feat_lo, feat_hi, n_top, ctx_len, int(image_n_base), device,
                                            bucket_cap=bucket_cap)
        self._x, self._k, self._prep, self._ws = [None, None], [0, 0], [None, None], [None, None]
        self._lb, self._ub, self._cached = [None, None], [None, None], [None, None]
        self._merged = [1, 1]   # how the slot's candidates were merged (already_merged of saeb_refine_candidates)
        # Candidates kept per token and shard = k + margin (0 = the library default max(48, k/2)).  A shard of an R-way
        # scan (R >= 4) holds ~k/R of a token's TopK, so its list can be far shorter than the unsharded k + 48: k + 8 makes the
        # GEMM's epilogue and the whole chain cheaper (one rank of 8, ms per chunk: 5.67 -> 5.24, GEMM stream alone
        # 4.80 -> 4.52; profiles/r02ah_scan_rank_emul_margin.log).  Exactness does not depend on it: a list that could be
        # too short is flagged and the row recomputed by the exact dense kernels.  Fixed per slot at local_prep time.
        self.margin_sharded = int(os.environ.get("SAEB_SCAN_MARGIN", "8"))
        self.margin = 0
        self._margin = [0, 0]
        self.packed_bounds = os.environ.get("SAEB_SCAN_PACKED_BOUNDS", "1") != "0"
        self.status = torch.zeros(1, dtype=torch.int32, device=device)
        # SAEB_SCAN_AUX_PRIORITY = low (default: GEMM CTAs are placed first at launch boundaries; measured 1 % faster
        # at 8 GPUs, profiles/r02t_scan_sweep_n8_1M.log) | high (exchange / refine / list update get the first pick of
        # whatever SM resources the GEMM grid leaves free)
        aux_priority = aux_priority or os.environ.get("SAEB_SCAN_AUX_PRIORITY", "low")
        self.stream_gemm = torch.cuda.Stream(device, priority=0 if aux_priority == "high" else -1)
        self.stream_aux = torch.cuda.Stream(device, priority=-1 if aux_priority == "high" else 0)
        # > 0: the sharded refinement runs as a bounded persistent grid of that many CTAs and its helper launches use
        # small blocks, so that nothing of the per-chunk chain has to wait for a GEMM launch boundary (saeb200.h,
        # `max_ctas`); 0: one CTA per token
        self.refine_max_ctas = int(os.environ.get("SAEB_SCAN_REFINE_CTAS", "0"))
        # 2 (default): the refinement only gathers latents that can still enter their feature's list ("scan" mode of
        # the refinement, exact values, rigorous membership); 0: every member of every token's TopK is re-evaluated
        self.scan_value_mode = int(os.environ.get("SAEB_SCAN_VALUE_MODE", "2")) if image_n_base is None else 0
        # SMs the GEMM grid leaves idle while a multi-rank scan is pipelined (see begin_pipeline).  Measured at 8 GPUs
        # (1 M tokens): 0 -> 4.85 M tokens/s, 4 (with NCCL_MAX_CTAS=4) -> 4.65-4.76 M: off by default
        self.reserve_sms = 0
        # per-chunk exchanges: "push" (default: saeb_push_gather, this library's own peer-memory all-gather over NVLink
        # / NVSwitch multicast, 31 registers and no shared memory, so it runs beside the GEMM CTAs; verified bit-exact
        # against NCCL on 2 and 8 GPUs, profiles/r02d_push_gather_n2.json, r02e_scan_sweep_n8.log; every rank falls
        # back to NCCL together when symmetric memory cannot be set up) or "nccl" (all_gather_into_tensor, whose
        # kernels wait for a GEMM launch boundary); SAEB_SCAN_EXCHANGE / bench.py --scan-exchange
        self.exchange = os.environ.get("SAEB_SCAN_EXCHANGE", "push")
        self.push_widths = None   # (exchange-1 width, exchange-2 width), set by sharded_scan
        self._push = None
        # Pipelined scans run the per-chunk chain (merge, bounds, exchange, kth, refinement, list update) INSIDE the next
        # chunk's GEMM launches.  That only works if every kernel of the chain fits beside a resident GEMM CTA: the
        # GEMM's shared-memory ring is one stage shorter (5 x 32 KB: ~48 KB per SM left, < 1 % slower GEMM) and the
        # chain uses its small-footprint launch shapes (`coresident`).  A kernel that does not fit waits for the next
        # GEMM launch boundary (~1 ms each).  SAEB_SCAN_GEMM_STAGES=0 / SAEB_SCAN_CORESIDENT=0 restore the round-1 shapes.
        self.gemm_stages = int(os.environ.get("SAEB_SCAN_GEMM_STAGES", "5"))
        # activation prep of chunk c+1 on the small-kernel stream (see _pipelined_loop)
        # (measured: no gain -- the step is energy-bound, moving work between streams does not remove it; opt-in)
        self.prep_ahead = os.environ.get("SAEB_SCAN_PREP_AHEAD", "0") != "0"
        self.coresident = os.environ.get("SAEB_SCAN_CORESIDENT", "1") != "0"

    def push_gather(self, t, group, channel, slot):
        """[T, m] -> [R, T, m] through the peer-memory exchange, or None when it is not selected (caller uses NCCL).
        The symmetric buffer is created at the first call (a collective: every rank gets here with the same chunk)."""
        if self.exchange != "push" or self.push_widths is None:
            return None
        if self._push is None or self._push.max_rows < t.shape[0] or tuple(self._push.widths) != tuple(self.push_widths):
            from .p2p import PushExchange

            # setting up the symmetric buffer can fail on one rank only (no peer access, no symmetric-memory support):
            # every rank tries, then all agree -- either everybody pushes or everybody stays on NCCL
            err = None
            try:
                push = PushExchange(group, t.device, t.shape[0], self.push_widths)
            except Exception as exc:   # noqa: BLE001 - any failure means "fall back", the reason is reported below
                push, err = None, exc
            ok = torch.tensor([0 if push is None else 1], dtype=torch.int32, device=t.device)
            dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=group)
            if int(ok.item()) == 0:
                if dist.get_rank(group) == 0 or err is not None:
                    print(f"saeb200: peer-memory exchange unavailable ({err!r}); using NCCL all-gathers", flush=True)
                self.exchange, self._push = "nccl", None
                return None
            self._push = push
        return self._push.gather(t, channel, slot)

    def begin_pipeline(self, world: int):
        """The persistent GEMM grid normally owns every SM; with `reserve_sms` > 0 it leaves a few free while the scan
        is pipelined so that NCCL's kernels (which cannot co-reside with a GEMM CTA) never have to displace one
        (pair it with NCCL_MAX_CTAS <= reserve_sms in the environment)."""
        if world > 1 and self.reserve_sms > 0:
            self._capi.check(self._capi.lib().saeb_set_option(b"reserve_sms", self.reserve_sms), "set_option")
        rt = os.environ.get("SAEB_REFINE_THREADS")   # tuning knob of the co-resident refinement (see saeb200.h)
        if rt:
            self._capi.check(self._capi.lib().saeb_set_option(b"refine_threads", int(rt)), "set_option")
        if self.coresident:
            self._capi.check(self._capi.lib().saeb_set_option(b"gemm_stages", self.gemm_stages), "set_option")
            self.scan.coresident = True
        self.margin = self.margin_sharded if world >= 4 else 0   # measured at 8 ranks; 2 ranks keep the default lists

    def chunk_tokens(self, world: int, waves: Optional[int] = None) -> int:
        """Tokens per scan chunk = `waves` full single-wave GEMM launches (256-row tiles on half of the CTA pairs the
        grid may use), so that no launch runs partly empty.  Default 4 waves (SAEB_SCAN_WAVES overrides): every chunk
        costs two cross-rank synchronisation points under sharding, larger chunks mean fewer of them."""
        if waves is None:
            waves = int(os.environ.get("SAEB_SCAN_WAVES", "4"))
        sms = int(self._capi.lib().saeb_query(b"num_sms")) or 148
        if world > 1:
            sms -= self.reserve_sms
        return waves * 256 * max(1, (sms // 2) // 2)

    def end_pipeline(self):
        self.margin = 0
        self._capi.check(self._capi.lib().saeb_set_option(b"reserve_sms", 0), "set_option")
        if self.coresident:
            self._capi.check(self._capi.lib().saeb_set_option(b"gemm_stages", 0), "set_option")
            self.scan.coresident = False

    # ---- mode 3: GEMM -> bounds -> (exchange) -> refinement restricted by the global lower bound
    def _scratch(self, store, slot, nbytes, dev):
        """per-slot scratch that is never handed back to the allocator while the scan runs (it is shared between the
        two streams, ordered by events)"""
        if store[slot] is None or store[slot].numel() < nbytes:
            store[slot] = torch.empty(nbytes, dtype=torch.uint8, device=dev)
        return store[slot]

    def local_prep(self, x, k, slot=0):
        """activation prep of a chunk (row scales, fp16 plane, row norms) into the slot's scratch, on the current stream.
        The pipelined scan runs it on the SMALL-kernel stream one chunk ahead, so that the tensor-core stream holds
        nothing but GEMM launches; `local_gemm(..., prepped=True)` then skips it."""
        eng, L = self.engine, self._capi.lib()
        enc = self.enc
        x2 = eng._as_2d(x, enc.d_in)
        self._x[slot], self._k[slot], self._margin[slot] = x2, k, int(self.margin)
        if enc.planes < 3:
            return
        T = x2.shape[0]
        dev = x2.device
        ldx = x2.stride(0) if T > 1 else enc.d_in
        with torch.cuda.device(dev):
            prep = self._scratch(self._prep, slot, L.saeb_prep_bytes(T, enc.d_in), dev)
            self._capi.check(L.saeb_prep_activations(x2.data_ptr(), eng._code(x2), T, ldx, enc.d_in, prep.data_ptr(),
                                                     torch.cuda.current_stream().cuda_stream), "saeb_prep_activations")

    def local_gemm(self, x, k, slot=0, prepped=False):
        """tensor-core half of a chunk: activation prep (unless `local_prep` already did it) + the fused GEMM launches
        with candidate selection (enqueued on the current stream).  `local_bounds_finish(slot)` completes it;
        `local_bounds` = both."""
        eng, L = self.engine, self._capi.lib()
        enc = self.enc
        if not prepped:
            self.local_prep(x, k, slot)
        x2 = self._x[slot]
        T = x2.shape[0]
        if enc.planes < 3:
            vals, idx, _ = eng.encode_topk(x2, enc, k)
            self._cached[slot] = (vals, idx)
            return
        dev = x2.device
        with torch.cuda.device(dev):
            prep = self._prep[slot]
            ws = self._scratch(self._ws, slot, L.saeb_candidates_workspace_bytes(T, enc.d_in, enc.num_latents, k, self._margin[slot]),
                               dev)
            st = torch.cuda.current_stream().cuda_stream
            self._capi.check(L.saeb_encode_candidates(prep.data_ptr(), T, 0, T, enc.blob.data_ptr(), enc.d_in,
                                                      enc.num_latents, k, self._margin[slot], -1, 0.0, ws.data_ptr(), ws.numel(), st),
                             "saeb_encode_candidates")

    def local_bounds_finish(self, slot=0, coresident=False, pack_m1=None):
        """candidate merge + per-token bound lists of the chunk given to `local_gemm` -> (lb, ub), both [Tc, k]
        descending.  Small kernels: with `coresident` they are shaped to run on another stream INSIDE the next chunk's
        GEMM launches.  `pack_m1` (sharded scans): return the payload of exchange 1 instead, ONE tensor [Tc, 2*pack_m1]
        (the m1 largest lower | upper bounds, unsorted) from the fused register-resident kernel
        (saeb_candidate_bounds_packed); falls back to (lb, ub) when that kernel does not take the shape."""
        eng, L = self.engine, self._capi.lib()
        enc, x2, k = self.enc, self._x[slot], self._k[slot]
        self._merged[slot] = 1
        if enc.planes < 3:
            vals = self._cached[slot][0]
            return vals, vals
        T = x2.shape[0]
        dev = x2.device
        if pack_m1 is not None and enc.planes == 3 and self.scan_value_mode == 2 and self.packed_bounds:
            with torch.cuda.device(dev):
                out = self._scratch(self._lb, slot, T * max(k, 2 * pack_m1) * 4, dev)[: T * 2 * pack_m1 * 4].view(torch.float32)
                out = out.view(T, 2 * pack_m1)
                rc = L.saeb_candidate_bounds_packed(self._prep[slot].data_ptr(), T, 0, T, enc.blob.data_ptr(),
                                                    eng._code(x2), enc.d_in, enc.num_latents, k, self._margin[slot], -1, int(pack_m1),
                                                    out.data_ptr(), self._ws[slot].data_ptr(), self._ws[slot].numel(),
                                                    torch.cuda.current_stream().cuda_stream)
            if rc == 0:
                self._merged[slot] = 2
                return out
            if rc != 1:   # 1 = shape not taken by the fused kernel: two-kernel route below
                self._capi.check(rc, "saeb_candidate_bounds_packed")
        with torch.cuda.device(dev):
            prep, ws = self._prep[slot], self._ws[slot]
            lb = self._scratch(self._lb, slot, T * k * 4, dev)[: T * k * 4].view(torch.float32).view(T, k)
            ub = self._scratch(self._ub, slot, T * k * 4, dev)[: T * k * 4].view(torch.float32).view(T, k)
            st = torch.cuda.current_stream().cuda_stream
            self._capi.check(L.saeb_candidate_bounds(prep.data_ptr(), T, 0, T, enc.blob.data_ptr(), eng._code(x2),
                                                     enc.d_in, enc.num_latents, k, self._margin[slot], -1, lb.data_ptr(), ub.data_ptr(),
                                                     ws.data_ptr(), ws.numel(), 1 if coresident else 0, st),
                             "saeb_candidate_bounds")
        return lb, ub

    def local_bounds(self, x, k, slot=0):
        self.local_gemm(x, k, slot)
        return self.local_bounds_finish(slot)

    def local_topk(self, ext_L=None, ext_U=None, slot=0):
        """exact local TopK entries of the chunk given to local_bounds -> (vals, member, ids), all [Tc, k]; the ids
        are relative to this shard's first feature (`scan_update` knows).
        Refinement in "scan" mode (value_mode 2, include/saeb200.h): only latents that can still enter their feature's
        top-n list (and every latent whose membership in the token's TopK is undecided) are gathered and re-evaluated,
        exactly.  Unsharded call (ext_L = ext_U = None): membership is decided here, `member` is None.  Sharded call:
        `member` carries what decides membership across shards (3e38 = certain, exact value = boundary candidate)."""
        eng, L = self.engine, self._capi.lib()
        enc, x2, k = self.enc, self._x[slot], self._k[slot]
        if enc.planes < 3:
            vals, idx = self._cached[slot]
            return vals, (vals if ext_L is not None else None), idx
        refine = L.saeb_refine_candidates_lo if enc.planes == 4 else L.saeb_refine_candidates
        T = x2.shape[0]
        dev = x2.device
        vals = torch.empty((T, k), dtype=torch.float32, device=dev)
        idx = torch.empty((T, k), dtype=torch.int64, device=dev)
        sharded = ext_L is not None and ext_U is not None
        member = torch.empty((T, k), dtype=torch.float32, device=dev) if sharded else None
        mode = 2 if self.scan_value_mode == 2 and (sharded or ext_L is None) else 0
        with torch.cuda.device(dev):
            st = torch.cuda.current_stream().cuda_stream
            self._capi.check(refine(
                x2.data_ptr(), eng._code(x2), x2.stride(0) if T > 1 else enc.d_in, self._prep[slot].data_ptr(), T, 0, T,
                enc.blob.data_ptr(), enc.W_enc.data_ptr(), enc.d_in, enc.num_latents, k, self._margin[slot], -1, 0.0,
                None if ext_L is None else ext_L.data_ptr(),
                ext_U.data_ptr() if (sharded and mode == 2) else None,
                self.scan.feat_thr.data_ptr() if mode == 2 else None, int(self._merged[slot]), vals.data_ptr(),
                member.data_ptr() if (sharded and mode == 2) else None, idx.data_ptr(), self.status.data_ptr(),
                self._ws[slot].data_ptr(), self._ws[slot].numel(), int(self.refine_max_ctas), mode, st),
                "saeb_refine_candidates")
        if sharded and mode != 2:
            member = vals
        return vals, member, idx   # shard-LOCAL ids: scan_update passes the offset on (no kernel just to add it)

    def kth_of_gathered(self, gathered, kth=None):
        return self.engine.kth_of_gathered(gathered, kth)

    def gathered_bounds(self, gathered, m1, k):
        return self.engine.gathered_bounds(gathered, m1, k)

    def scan_update(self, vals, idx, window_base, tok_thr, member=None):
        same = member is None or member is vals or member.data_ptr() == vals.data_ptr()   # "every member exact" mode
        self.scan.update(vals, idx, window_base, tok_thr, None if same else member, idx_base=self.feat_lo)

    def scan_finalize(self):
        return self.scan.finalize()


@dataclass
class ScanResult:
    top_vals: torch.Tensor  # [N, n_top] f32, every rank holds the full table after the final all-gather
    top_win: torch.Tensor  # [N, n_top] i64, -1 = empty


def _all_gather_cat(t: torch.Tensor, group, sizes=None) -> torch.Tensor:
    world = dist.get_world_size(group)
    if sizes is None:
        out = [torch.empty_like(t) for _ in range(world)]
    else:
        out = [torch.empty((s,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device) for s in sizes]
    dist.all_gather(out, t.contiguous(), group=group)
    return out


class _PhaseTimer:
    """optional CUDA-event phase timing of sharded_scan (diagnostics; synchronises at the end only)"""

    def __init__(self, enabled: bool):
        self.enabled, self.marks = enabled, []

    def mark(self, name: str):
        if self.enabled:
            ev = torch.cuda.Event(enable_timing=True)
            ev.record()
            self.marks.append((name, ev))

    def totals(self):
        out = {}
        if not self.enabled:
            return out
        torch.cuda.synchronize()
        for (n0, e0), (n1, e1) in zip(self.marks, self.marks[1:]):
            out[n1] = out.get(n1, 0.0) + e0.elapsed_time(e1)
        return out


def sharded_scan(chunks: Iterable[torch.Tensor], ops, k: int, ctx_len: int, num_latents: int, *, exact: bool = True,
                 group=None, phase_times: Optional[dict] = None, pipelined: Optional[bool] = None,
                 lb_width: Optional[int] = None, local: bool = False) -> ScanResult:
    """Feature-sharded scan.  `chunks` yields the SAME token chunks ([Tc, d], Tc a multiple of ctx_len) on every rank.

    exact=True, per chunk: (1) every shard computes lower bounds of its k best latents per token and all-gathers
    them -> per-token lower bound of the global k-th value; (2) the shard evaluates exactly only the latents that can
    still reach it and all-gathers its exact local top-k values -> the per-token global k-th value, which filters what
    enters the per-feature lists (the cache keeps a latent only if it is in the token's global TopK,
    features/cache.py:210-218).  Exchange 1 is [Tc, lb_width] fp32 per rank (`bounds_width`: 24 columns at k = 64 on
    8 ranks), exchange 2 [Tc, k] (256 B/token/rank at k = 64).

    Schedules (`pipelined`, default = `ops.pipelined` unless `phase_times` asks for the sequential diagnostic one):
      * "streams" (default): the GEMM of chunk c+1 runs on `ops.stream_gemm` while exchange + refinement + list
        update of chunk c run on `ops.stream_aux`;
      * "lookahead" (SAEB_SCAN_SCHEDULE=lookahead, and what ops without streams get): one compute stream with a
        one-chunk lookahead; both all-gathers are issued asynchronously -- exchange 1 of chunk c right after its
        GEMM, consumed after the GEMM of chunk c+1; exchange 2 of chunk c after its refinement, consumed after the
        refinement of chunk c+1.
    Measured, 1 M tokens, tokens/s (sequential / streams / lookahead): 8 GPUs 4.55 M / 4.76-4.85 M / 4.64 M;
    4 GPUs - / 3.14 M / 2.97 M; 2 GPUs - / 1.78 M / 1.70 M; 1 GPU 0.916 M / 0.951 M / -."""
    distributed = dist.is_available() and dist.is_initialized() and not local   # local: a one-rank scan inside a job
    world = dist.get_world_size(group) if distributed else 1
    k_local = min(k, ops.feat_hi - ops.feat_lo)
    m1 = bounds_width(k, k_local, world) if lb_width is None else max(1, min(k_local, int(lb_width)))
    m1 = min(k_local, max(m1, -(-(k + 1) // max(world, 1))))   # the union of the shards' lists must hold k + 1 entries
    if pipelined is None:
        pipelined = bool(getattr(ops, "pipelined", False)) and phase_times is None
    tm = _PhaseTimer(phase_times is not None and torch.cuda.is_available())
    tm.mark("start")
    exchange = exact and world > 1

    if exchange and hasattr(ops, "push_widths"):
        ops.push_widths = (2 * m1, k_local)

    def finish(x, bounds, slot, window_base):
        """exchange 1 (bounds) -> restricted exact local TopK -> exchange 2 (member values) -> list update, one chunk"""
        ext_L = ext_U = tok_thr = None
        if exchange:
            # one all-gather carries both bound lists: [Tc, m1] lower | [Tc, m1] upper
            payload = bounds if torch.is_tensor(bounds) else torch.cat([_head(bounds[0], m1), _head(bounds[1], m1)], dim=-1)
            g = _exchange(ops, payload, group, 0, slot)
            ext_L, ext_U = _bounds_of_gathered(ops, g, m1, k)
            tm.mark("exchange1")
        args = (ext_L, ext_U, slot) if slot is not None else (ext_L, ext_U)
        vals, member, idx = ops.local_topk(*args)
        tm.mark("refine")
        vals2 = vals.reshape(-1, k_local)
        member2 = None if member is None else member.reshape(-1, k_local)
        if exchange:
            tok_thr = _kth(ops, _exchange(ops, member2, group, 1, slot), k)
            tm.mark("exchange2")
        ops.scan_update(vals2, idx.reshape(-1, k_local), window_base, tok_thr, member2 if exchange else None)
        tm.mark("scan_update")
        return vals2.shape[0] // ctx_len

    window_base = 0
    if not pipelined:
        fused = (exchange and getattr(ops, "packed_bounds", False)
                 and hasattr(ops, "local_gemm") and hasattr(ops, "local_bounds_finish"))
        for x in chunks:
            if fused:   # the same kernels as the pipelined schedules, one after the other
                ops.local_gemm(x, k_local, 0)
                tm.mark("gemm")
                lb = ops.local_bounds_finish(0, False, pack_m1=m1)
                tm.mark("select+bounds")
            else:
                lb = ops.local_bounds(x, k_local)
                tm.mark("gemm+bounds")
            window_base += finish(x, lb, 0 if fused else None, window_base)
    else:
        schedule = scan_schedule(world)
        if not hasattr(ops, "stream_gemm"):
            schedule = "lookahead"
        begin, end = getattr(ops, "begin_pipeline", None), getattr(ops, "end_pipeline", None)
        if begin is not None:
            begin(world)
        try:
            if schedule == "lookahead":
                _lookahead_loop(chunks, ops, k, k_local, ctx_len, exchange, group, m1)
            else:
                _pipelined_loop(chunks, ops, finish, k_local, ctx_len, ops.stream_gemm, ops.stream_aux,
                                torch.cuda.current_stream(),
                                pack_m1=m1 if exchange else None)
        finally:
            if end is not None:
                end()

    top_vals, top_win = ops.scan_finalize()
    tm.mark("scan_finalize")
    if phase_times is not None:
        phase_times.update(tm.totals())
    if world > 1:
        sizes = [shard_range(num_latents, world, r) for r in range(world)]
        sizes = [hi - lo for lo, hi in sizes]
        top_vals = torch.cat(_all_gather_cat(top_vals, group, sizes), 0)   # the single end-of-job all-gather
        top_win = torch.cat(_all_gather_cat(top_win, group, sizes), 0)
    return ScanResult(top_vals, top_win)


def _lookahead_loop(chunks, ops, k, k_local, ctx_len, exchange, group, m1) -> None:
    """world > 1 schedule of sharded_scan: per iteration c
         A(c)    GEMM + bounds of chunk c;            exchange 1 of chunk c   issued (async)
         B1(c-1) wait exchange 1 of c-1, refine c-1;  exchange 2 of chunk c-1 issued (async)
         B2(c-2) wait exchange 2 of c-2, list update of chunk c-2
    Scratch of chunk c lives in slot c & 1: A(c+2) is enqueued after B1(c), which is its last reader.  Every rank
    issues the collectives in the same order."""
    stage1 = None   # chunk waiting for exchange 1: (slot, gathered, work, window_base)
    stage2 = None   # chunk waiting for exchange 2: (vals, idx, gathered, work, window_base)
    window_base = 0

    def b1(item):
        slot, gathered, work, base = item
        ext_L = ext_U = None
        if exchange:
            work.wait()
            ext_L, ext_U = _bounds_of_gathered(ops, gathered, m1, k)
        vals, member, idx = ops.local_topk(ext_L, ext_U, slot)
        vals2, idx2 = vals.reshape(-1, k_local), idx.reshape(-1, k_local)
        mem2 = None if member is None else member.reshape(-1, k_local)
        g2, w2 = _gather_stack(mem2, group, async_op=True) if exchange else (None, None)
        return vals2, idx2, mem2, g2, w2, base

    def b2(item):
        vals2, idx2, mem2, gathered, work, base = item
        tok_thr = None
        if exchange:
            work.wait()
            tok_thr = _kth(ops, gathered, k)
        ops.scan_update(vals2, idx2, base, tok_thr, mem2 if exchange else None)

    for c, x in enumerate(chunks):
        slot = c & 1
        lb, ub = ops.local_bounds(x, k_local, slot)
        g1, w1 = (_gather_stack(torch.cat([_head(lb, m1), _head(ub, m1)], dim=-1), group, async_op=True)
                  if exchange else (None, None))
        if stage1 is not None:
            nxt = b1(stage1)
            if stage2 is not None:
                b2(stage2)
            stage2 = nxt
        stage1 = (slot, g1, w1, window_base)
        n_tok = x.shape[0] if x.dim() == 2 else x.numel() // x.shape[-1]
        window_base += n_tok // ctx_len
    if stage1 is not None:
        nxt = b1(stage1)
        if stage2 is not None:
            b2(stage2)
        b2(nxt)
    elif stage2 is not None:
        b2(stage2)


def _pipelined_loop(chunks, ops, finish, k_local, ctx_len, sg, sa, cur, pack_m1=None) -> int:
    """two-stream software pipeline of sharded_scan; returns the number of windows consumed"""
    sa.wait_stream(cur)
    window_base = 0
    slot_free = [None, None]   # event: the chunk that last used this slot's scratch has left stream_aux
    prev = None

    split = hasattr(ops, "local_gemm") and hasattr(ops, "local_bounds_finish")

    def drain(item):
        x, bounds, ready, slot, base = item
        with torch.cuda.stream(sa):
            sa.wait_event(ready)
            if bounds is None:   # merge + bounds belong to the small-kernel chain, not to the tensor-core stream
                if pack_m1 is not None and getattr(ops, "packed_bounds", False):
                    bounds = ops.local_bounds_finish(slot, coresident=True, pack_m1=pack_m1)
                else:
                    bounds = ops.local_bounds_finish(slot, coresident=True)
            finish(x, bounds, slot, base)
            slot_free[slot] = torch.cuda.Event()
            slot_free[slot].record(sa)

    # activation prep of chunk c+1 on the small-kernel stream, right behind the chain of chunk c-1 (which is the last
    # reader of that slot's scratch): the tensor-core stream then holds nothing but GEMM launches
    prefetch = split and hasattr(ops, "local_prep") and getattr(ops, "prep_ahead", False)
    prep_done = [None, None]

    def prep_ahead(x, slot):
        x.record_stream(sa)
        with torch.cuda.stream(sa):
            sa.wait_stream(cur)
            ops.local_prep(x, k_local, slot)
            prep_done[slot] = torch.cuda.Event()
            prep_done[slot].record(sa)

    it = iter(chunks)
    x_next = next(it, None)
    if prefetch and x_next is not None:
        prep_ahead(x_next, 0)
    c = 0
    while x_next is not None:
        x, x_next = x_next, next(it, None)
        slot = c & 1
        x.record_stream(sg)
        x.record_stream(sa)
        with torch.cuda.stream(sg):
            sg.wait_stream(cur)   # whatever produced this chunk on the caller's stream
            if slot_free[slot] is not None:
                sg.wait_event(slot_free[slot])
            if prefetch:
                sg.wait_event(prep_done[slot])
                ops.local_gemm(x, k_local, slot, prepped=True)
                bounds = None
            elif split:
                ops.local_gemm(x, k_local, slot)
                bounds = None
            else:
                bounds = ops.local_bounds(x, k_local, slot)
            ready = torch.cuda.Event()
            ready.record(sg)
        n_tok = x.shape[0] if x.dim() == 2 else x.numel() // x.shape[-1]
        if prev is not None:
            drain(prev)
        if prefetch and x_next is not None:
            prep_ahead(x_next, (c + 1) & 1)
        prev = (x, bounds, ready, slot, window_base)
        window_base += n_tok // ctx_len
        c += 1
    if prev is not None:
        drain(prev)
    cur.wait_stream(sg)
    cur.wait_stream(sa)
    return window_base


def _exchange(ops, t: torch.Tensor, group, channel: int, slot) -> torch.Tensor:
    """per-chunk exchange of [T, m] lists -> [R, T, m]: the ops' own peer-memory all-gather when it offers one
    (EngineOps.push_gather), else NCCL / gloo"""
    push = getattr(ops, "push_gather", None)
    if push is not None:
        g = push(t, group, channel, slot or 0)
        if g is not None:
            return g
    return _gather_stack(t, group)


def _gather_stack(t: torch.Tensor, group, async_op: bool = False):
    """all-gather equal-shaped [T, k] tensors into one [R, T, k] tensor (no extra copy); with `async_op` returns
    (tensor, work) and the caller waits on `work` before reading"""
    world = dist.get_world_size(group)
    out = torch.empty((world,) + tuple(t.shape), dtype=t.dtype, device=t.device)
    if hasattr(dist, "all_gather_into_tensor") and t.is_cuda:
        work = dist.all_gather_into_tensor(out, t.contiguous(), group=group, async_op=async_op)
    else:
        work = dist.all_gather(list(out.unbind(0)), t.contiguous(), group=group, async_op=async_op)
    return (out, work) if async_op else out


def _head(lb: torch.Tensor, m: int) -> torch.Tensor:
    """first m columns of the (descending) per-token bound lists"""
    return lb if m >= lb.shape[-1] else lb[..., :m]


def _bounds_of_gathered(ops, g: torch.Tensor, m1: int, k: int):
    """exchange 1 -> (ext_L, ext_U): the k-th largest gathered lower bound, and an upper bound of the token's (k+1)-th
    largest upper bound: the (k+1)-th largest of what was sent, or the smallest bound a shard sent if its list was cut
    off (whatever it did not send is no larger than that).  One fused kernel when the ops offer it."""
    fused = getattr(ops, "gathered_bounds", None)
    if fused is not None:
        return fused(g, m1, k)
    ext_L = _kth(ops, g[:, :, :m1], k)
    ext_U = torch.maximum(_kth(ops, g[:, :, m1:], k + 1), g[:, :, m1:].amin(-1).amax(0))
    return ext_L, ext_U


def _kth(ops, gathered: torch.Tensor, k: int) -> torch.Tensor:
    """per-token k-th largest of the R * m gathered values ([R, T, m]); 0 when there are fewer than k values (the
    k-th largest of fewer than k bounds bounds nothing: 0 is the safe answer for lower AND upper bounds of
    non-negative activations -- "no restriction" / "everything positive is a member")"""
    R, T, m = gathered.shape
    if k > R * m:
        return torch.zeros(T, dtype=gathered.dtype, device=gathered.device)
    return ops.kth_of_gathered(gathered, k)


def token_parallel_scan(chunks_local: Iterable[torch.Tensor], ops, k: int, ctx_len: int, num_latents: int,
                        window_base: int, *, n_top: int, group=None) -> ScanResult:
    """The other exact multi-GPU form of the scan: tokens are split across ranks, every rank holds the
    FULL SAE (`ops` covers features [0, num_latents)) and scans its own token slice with no per-chunk exchange at all
    -- a token's TopK is complete on the rank that owns it.  The job ends with one all-gather of the per-rank
    [N, n_top] lists and a per-feature merge.  It needs the whole SAE on every GPU (the feature-sharded form does
    not) and is used as its cross-check: both must produce identical lists.

    `chunks_local` yields this rank's token chunks ([Tc, d], Tc a multiple of ctx_len); `window_base` is the global id
    of its first window (ranks own consecutive, increasing window ranges, so concatenating the per-rank lists in rank
    order keeps equal scores ordered by window id and a stable sort finishes the merge)."""
    distributed = dist.is_available() and dist.is_initialized()
    world = dist.get_world_size(group) if distributed else 1
    base = int(window_base)
    for x in chunks_local:
        ops.local_bounds(x, k)
        vals, _, idx = ops.local_topk(None, None)
        ops.scan_update(vals.reshape(-1, k), idx.reshape(-1, k), base, None)
        n_tok = x.shape[0] if x.dim() == 2 else x.numel() // x.shape[-1]
        base += n_tok // ctx_len
    top_vals, top_win = ops.scan_finalize()
    if world == 1:
        return ScanResult(top_vals, top_win)
    vals_all = torch.cat(_all_gather_cat(top_vals, group), 1)   # [N, R * n_top], rank-major
    wins_all = torch.cat(_all_gather_cat(top_win, group), 1)
    order = torch.sort(vals_all, dim=1, descending=True, stable=True).indices[:, :n_top]
    return ScanResult(torch.gather(vals_all, 1, order), torch.gather(wins_all, 1, order))


def token_parallel_forward(sae, x_local: torch.Tensor):
    """Token-parallel forward: each rank runs the fused forward on its token slice; nothing is communicated."""
    return sae(x_local)
