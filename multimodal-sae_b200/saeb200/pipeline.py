"""Reference-facing forward with HOST buffers: x lives in pinned host memory, TopK results and the FVU (and, when
asked for, the reconstruction `sae_out`) come back to the host.  Token chunks are pipelined over three streams (H2D
copy / compute / D2H copy) so that the PCIe traffic hides behind the tensor-core work."""
from __future__ import annotations

import torch

from . import engine


class HostForward:
    def __init__(self, sae, num_tokens: int, chunk: int = 9472, x_dtype=torch.bfloat16):
        dev = sae.device
        self.sae, self.T, self.chunk = sae, num_tokens, min(chunk, num_tokens)
        k, d = sae.cfg.k, sae.d_in
        self.x_dev = torch.empty((num_tokens, d), dtype=x_dtype, device=dev)
        self.acts = torch.empty((num_tokens, k), dtype=torch.float32, device=dev)
        self.idx = torch.empty((num_tokens, k), dtype=torch.int64, device=dev)
        self.sae_out = torch.empty((num_tokens, d), dtype=torch.float32, device=dev)
        self.sq_err = torch.zeros((), dtype=torch.float64, device=dev)
        self.fvu_host = torch.empty((), dtype=torch.float32, pin_memory=True)
        self.s_in, self.s_out = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
        self.n_chunks = (num_tokens + self.chunk - 1) // self.chunk
        self.h2d_bytes = num_tokens * d * self.x_dev.element_size()
        self._d2h_topk = num_tokens * k * 12 + 4
        self._d2h_out = num_tokens * d * 4
        self.overlap = None
        if sae.encoder_planes in (3, 4):
            from .overlap import OverlappedForward

            self.overlap = OverlappedForward(sae.packed_encoder(), sae.W_dec.data, sae.b_dec.data, k, self.chunk,
                                             value_mode=sae._value_mode())

    def d2h_bytes(self, with_sae_out: bool) -> int:
        return self._d2h_topk + (self._d2h_out if with_sae_out else 0)

    def run(self, x_host: torch.Tensor, acts_host: torch.Tensor, idx_host: torch.Tensor,
            out_host: torch.Tensor = None) -> torch.Tensor:
        """x_host [T, d] pinned; acts_host [T, k] f32 / idx_host [T, k] i64 pinned outputs; out_host [T, d] f32 pinned
        (optional): the reconstruction, copied back chunk by chunk on the D2H stream.  Returns the pinned 0-dim FVU
        tensor (valid after the call, which synchronises)."""
        sae, main = self.sae, torch.cuda.current_stream()
        self.sq_err.zero_()
        ev_in = [torch.cuda.Event() for _ in range(self.n_chunks)]
        ev_done = [torch.cuda.Event() for _ in range(self.n_chunks)]
        self.s_in.wait_stream(main)
        with torch.cuda.stream(self.s_in):
            for c in range(self.n_chunks):
                a, b = c * self.chunk, min(self.T, (c + 1) * self.chunk)
                self.x_dev[a:b].copy_(x_host[a:b], non_blocking=True)
                ev_in[c].record(self.s_in)
        if self.overlap is not None:
            self.overlap.run(self.x_dev, self.acts, self.idx, self.sae_out, self.sq_err, ready_events=ev_in,
                             done_events=ev_done)
        else:
            for c in range(self.n_chunks):
                a, b = c * self.chunk, min(self.T, (c + 1) * self.chunk)
                main.wait_event(ev_in[c])
                acts, idx, _ = engine.encode_topk(self.x_dev[a:b], sae.packed_encoder(), sae.cfg.k,
                                                  out_vals=self.acts[a:b], out_idx=self.idx[a:b])
                engine.decode(idx, acts, sae.W_dec.data, sae.b_dec.data, x=self.x_dev[a:b], sq_err=self.sq_err,
                              out=self.sae_out[a:b])
                ev_done[c].record(main)
        with torch.cuda.stream(self.s_out):
            for c in range(self.n_chunks):
                a, b = c * self.chunk, min(self.T, (c + 1) * self.chunk)
                self.s_out.wait_event(ev_done[c])
                acts_host[a:b].copy_(self.acts[a:b], non_blocking=True)
                idx_host[a:b].copy_(self.idx[a:b], non_blocking=True)
                if out_host is not None:
                    out_host[a:b].copy_(self.sae_out[a:b], non_blocking=True)
        tv = engine.total_variance(self.x_dev)
        fvu = (self.sq_err / tv).to(torch.float32)
        self.fvu_host.copy_(fvu, non_blocking=True)
        main.wait_stream(self.s_out)
        main.synchronize()
        return self.fvu_host
