"""Chunked forward with the tensor-core-bound and the HBM-bound halves on different streams.

Per token chunk the "fp16 + refine" path has two phases with opposite bottlenecks:
  A  activation prep + single-pass tcgen05 GEMM with fused candidate selection             (tensor pipe)
  B  candidate merge + exact fp32 refinement + sparse decode (+ residual sum of squares)    (HBM row gathers)
Phase A of chunk c+1 is issued on the GEMM stream while phase B of chunk c runs on the memory stream.

What makes the two phases actually share the SMs (round 1 measured ~1 % gain from two plain streams):
  * the GEMM is one persistent CTA per SM that owns ~206 KB of its shared memory, so a gather kernel can only run
    beside it if EVERY launch of phase B fits in what is left (~21 KB, 24 K registers per SM).  Phase B therefore runs
    as persistent grids of `ctas_per_sm` x #SM small CTAs that walk the chunk's tokens (`max_ctas` of the C ABI), its
    helper launches (merge, the flagged-row fallback grids that normally exit at once) use small blocks, and nothing
    in it ever has to wait for a GEMM launch boundary;
  * a one-CTA-per-token gather grid would flood every SM that a finishing GEMM CTA frees and keep the next GEMM launch
    out until it has drained -- with a bounded grid the GEMM CTAs always find their SM;
  * the GEMM stream has the higher priority, so at a launch boundary its CTAs are placed first.
The last chunk's phase B has nothing to hide behind and runs with full grids.
"""
from __future__ import annotations

import os
from typing import Optional

import torch

from . import _capi, engine
from ._capi import check


WAVE_TOKENS = 9472  # 37 token tiles of 256 rows: one wave of (tile, split) units on a 148-SM part at S = 2


def _env_int(name: str, default: int) -> int:
    try:
        return int(os.environ.get(name, default))
    except ValueError:
        return default


class OverlappedForward:
    N_WS = 3   # candidate workspaces in rotation: the GEMM may run two chunks ahead of the gathers

    def __init__(self, enc: engine.PackedEncoder, W_dec: torch.Tensor, b_dec: torch.Tensor, k: int,
                 chunk: int = WAVE_TOKENS, ctas_per_sm: Optional[int] = None, priority: Optional[str] = None,
                 value_mode: int = engine.VALUES_EXACT):
        if enc.planes not in (3, 4):
            raise _capi.SaebError("OverlappedForward needs a refine-mode packed encoder (planes = 3 or 4)")
        self.enc, self.W_dec, self.b_dec, self.k, self.chunk = enc, W_dec, b_dec, k, chunk
        self.value_mode = int(value_mode)
        dev = enc.blob.device
        self.dev = dev
        L = _capi.lib()
        with torch.cuda.device(dev):
            self.num_sms = int(L.saeb_query(b"num_sms"))
        # gather CTAs per SM while a GEMM launch is resident (0 = one CTA per token) and depth of the GEMM's shared-memory
        # ring while the forward is pipelined.  Measured (round 2, profiles/r02o_forward_overlap_boundary_mode.log, ms per
        # 65 536-token step): sequential 72.4-79.5, unbounded grids 72.2-74.0, 2 CTAs per SM beside a 5-stage ring
        # 70.7-71.2 (3: 70.4-72.8).  Every kernel of the step runs at the board's power cap on its own (the step is
        # energy-bound), so the gain is small; it only exists since every kernel asks for the largest shared-memory
        # carve-out (common.cuh, SAEB_CARVEOUT).
        self.ctas_per_sm = _env_int("SAEB_OV_CTAS_PER_SM", 2) if ctas_per_sm is None else int(ctas_per_sm)
        self.gemm_stages = _env_int("SAEB_OV_GEMM_STAGES", 5 if self.ctas_per_sm > 0 else 0)
        priority = os.environ.get("SAEB_OV_PRIORITY", "gemm") if priority is None else priority
        self.ws_bytes, self.ws = 0, [None] * self.N_WS
        self._reserve(chunk)
        self.prep = None
        self.status = torch.zeros(self.N_WS, dtype=torch.int32, device=dev)
        lo, hi = 0, -1   # CUDA: a numerically lower priority is the higher one
        pg, pm = {"gemm": (hi, lo), "mem": (lo, hi), "none": (lo, lo)}[priority]
        self.s_gemm = torch.cuda.Stream(dev, priority=pg)
        self.s_mem = torch.cuda.Stream(dev, priority=pm)

    def _reserve(self, *chunk_rows: int) -> None:
        """candidate workspaces large enough for every chunk size of a call (the requirement is not monotonic in the
        row count: fewer token tiles mean more feature-range splits per tile)"""
        L = _capi.lib()
        enc = self.enc
        need = max(L.saeb_candidates_workspace_bytes(r, enc.d_in, enc.num_latents, self.k, 0) for r in chunk_rows if r > 0)
        if need > self.ws_bytes:
            self.ws_bytes = need
            self.ws = [torch.empty(need, dtype=torch.uint8, device=self.dev) for _ in range(self.N_WS)]

    def run(self, x: torch.Tensor, acts: torch.Tensor, idx: torch.Tensor, sae_out: Optional[torch.Tensor] = None,
            sq_err: Optional[torch.Tensor] = None, ready_events=None, done_events=None) -> None:
        """x [T, d] (bf16 / fp16 / fp32, 2-D, unit column stride); acts [T,k] f32, idx [T,k] i64 and sae_out [T,d]
        (optional) are filled; sq_err (0-dim f64, optional) accumulates sum((sae_out - x)^2).
        `ready_events[c]` (optional) gates chunk c's input (H2D copies); `done_events[c]` is recorded when chunk c's
        outputs are complete.  Returns after ENQUEUEING; the caller's current stream waits for completion."""
        L = _capi.lib()
        enc, k = self.enc, self.k
        refine = L.saeb_refine_candidates_lo if enc.planes == 4 else L.saeb_refine_candidates
        T = x.shape[0]
        n_chunks = (T + self.chunk - 1) // self.chunk
        self._reserve(min(T, self.chunk), T - (n_chunks - 1) * self.chunk)
        main = torch.cuda.current_stream()
        self.s_gemm.wait_stream(main)
        self.s_mem.wait_stream(main)
        code = engine._code(x)
        esz = x.element_size()
        ev_a = [torch.cuda.Event() for _ in range(n_chunks)]
        ev_b = [torch.cuda.Event() for _ in range(n_chunks)]
        ldx = x.stride(0) if T > 1 else enc.d_in
        need = L.saeb_prep_bytes(T, enc.d_in)
        if self.prep is None or self.prep.numel() < need:
            self.prep = torch.empty(need, dtype=torch.uint8, device=self.dev)
        prep = self.prep
        nws = self.N_WS
        with torch.cuda.device(self.dev):
            if self.gemm_stages:
                check(L.saeb_set_option(b"gemm_stages", self.gemm_stages), "set_option")
            for c in range(n_chunks):
                a, b = c * self.chunk, min(T, (c + 1) * self.chunk)
                ws = self.ws[c % nws]
                # a GEMM launch follows this chunk's gathers -> bounded persistent grids; last chunk -> full grids
                max_ctas = self.ctas_per_sm * self.num_sms if c + 1 < n_chunks else 0
                with torch.cuda.stream(self.s_gemm):
                    if ready_events is not None:
                        self.s_gemm.wait_event(ready_events[c])
                    if c == 0 and ready_events is None:
                        # whole batch at once: 16 KB of traffic per token, negligible next to the GEMM
                        check(L.saeb_prep_activations(x.data_ptr(), code, T, ldx, enc.d_in, prep.data_ptr(),
                                                      self.s_gemm.cuda_stream), "saeb_prep_activations")
                    elif ready_events is not None:
                        # inputs arrive chunk by chunk (host copies): prepare each chunk when it lands
                        self._prep_rows(L, x, code, ldx, T, a, b, prep)
                    if c >= nws:
                        self.s_gemm.wait_event(ev_b[c - nws])   # scratch reuse
                    check(L.saeb_encode_candidates(prep.data_ptr(), T, a, b - a, enc.blob.data_ptr(), enc.d_in,
                                                   enc.num_latents, k, 0, -1, 0.0, ws.data_ptr(), ws.numel(),
                                                   self.s_gemm.cuda_stream), "saeb_encode_candidates")
                    ev_a[c].record(self.s_gemm)
                with torch.cuda.stream(self.s_mem):
                    self.s_mem.wait_event(ev_a[c])
                    check(refine(x.data_ptr() + a * ldx * esz, code, ldx, prep.data_ptr(), T, a,
                                 b - a, enc.blob.data_ptr(), enc.W_enc.data_ptr(), enc.d_in,
                                 enc.num_latents, k, 0, -1, 0.0, None, None, None, 0, acts[a:b].data_ptr(),
                                 None, idx[a:b].data_ptr(), self.status[c % nws:].data_ptr(), ws.data_ptr(),
                                 ws.numel(), max_ctas, self.value_mode, self.s_mem.cuda_stream),
                          "saeb_refine_candidates")
                    if sae_out is not None:
                        engine.decode(idx[a:b], acts[a:b], self.W_dec, self.b_dec,
                                      x=x[a:b] if sq_err is not None else None, sq_err=sq_err, out=sae_out[a:b],
                                      max_ctas=max_ctas)
                    ev_b[c].record(self.s_mem)
                    if done_events is not None:
                        done_events[c].record(self.s_mem)
            if self.gemm_stages:
                check(L.saeb_set_option(b"gemm_stages", 0), "set_option")
        main.wait_stream(self.s_mem)
        main.wait_stream(self.s_gemm)

    def _prep_rows(self, L, x, code, ldx, T, a, b, prep):
        """Prepare rows [a, b) in place inside the whole-batch `prep` layout (x16 | row_scale | xnorm)."""
        d = self.enc.d_in
        d_pad = (d + 7) // 8 * 8
        # the per-row outputs are independent, so a chunk is prepared through a view of the batch layout: the C entry
        # point lays out its three arrays from the base pointer and T, hence one call per chunk on a temporary and a
        # strided copy would be wasteful -- instead prepare directly with row offsets via three sub-calls' worth of
        # pointer arithmetic (x16 rows are contiguous per row; scales / norms are plain arrays).
        tmp_bytes = L.saeb_prep_bytes(b - a, d)
        if not hasattr(self, "_tmp") or self._tmp.numel() < tmp_bytes:
            self._tmp = torch.empty(tmp_bytes, dtype=torch.uint8, device=self.dev)
        tmp = self._tmp
        check(L.saeb_prep_activations(x.data_ptr() + a * ldx * x.element_size(), code, b - a, ldx, d, tmp.data_ptr(),
                                      self.s_gemm.cuda_stream), "saeb_prep_activations")
        n = b - a

        def layout(rows):  # mirrors prep_layout() of csrc/capi.cu: x16 | row_scale | xnorm | xdnorm
            rs = (rows * d_pad * 2 + 1023) // 1024 * 1024
            step = (rows * 4 + 255) // 256 * 256
            return rs, rs + step, rs + 2 * step

        src, dst = layout(n), layout(T)
        prep[a * d_pad * 2:b * d_pad * 2].copy_(tmp[:n * d_pad * 2], non_blocking=True)
        for so, do in zip(src, dst):
            prep[do + a * 4:do + b * 4].copy_(tmp[so:so + n * 4], non_blocking=True)
