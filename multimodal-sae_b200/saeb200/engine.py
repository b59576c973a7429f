"""Torch-facing wrappers over the C ABI: PyTorch owns device memory and streams, the library does the math.

Every function requires CUDA tensors and raises otherwise -- there is deliberately no CPU / eager fallback.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional, Tuple

import torch

from . import _capi
from ._capi import SaebError, check

_DT = {torch.float32: _capi.F32, torch.bfloat16: _capi.BF16, torch.float16: _capi.F16}
ACT_THRESHOLD = 1e-5  # reference features/cache.py:80-81
# SAEB_DEBUG_CHECKS=1: read the kernels' device-side error flags after every call (synchronises) and raise on an
# out-of-range index; off by default, the flags stay readable as `decode.last_err_flag` / `decode_backward.last_err_flag`
DEBUG_CHECKS = __import__("os").environ.get("SAEB_DEBUG_CHECKS", "0") == "1"
VALUES_EXACT, VALUES_BOUNDARY = 0, 1  # `value_mode` of the refinement (include/saeb200.h)


def _code(t: torch.Tensor) -> int:
    try:
        return _DT[t.dtype]
    except KeyError:
        raise SaebError(f"unsupported dtype {t.dtype}") from None


def _need_cuda(*ts: torch.Tensor) -> None:
    for t in ts:
        if t is not None and not t.is_cuda:
            raise SaebError("saeb200 engine needs CUDA tensors: there is no CPU fallback "
                            f"(got a tensor on {t.device})")


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


_ws_cache = {}


def _workspace(device: torch.device, nbytes: int, tag: str = "enc") -> torch.Tensor:
    """scratch cached per (device, purpose, stream): kernels of one stream reuse it in stream order, calls issued on
    different streams never share it"""
    key = (device.index, tag, torch.cuda.current_stream(device).cuda_stream)
    ws = _ws_cache.get(key)
    if ws is None or ws.numel() < nbytes:
        ws = None
        _ws_cache.pop(key, None)
        ws = torch.empty(max(nbytes, 1), dtype=torch.uint8, device=device)
        _ws_cache[key] = ws
    return ws


def release_workspaces() -> None:
    _ws_cache.clear()


@dataclass
class PackedEncoder:
    """Device-resident repack of (encoder.weight, encoder.bias, b_dec): bf16 planes + folded bias."""

    blob: torch.Tensor  # uint8
    num_latents: int
    d_in: int
    planes: int  # 1 / 2: bf16 planes (2 = hi+lo, parity grade); 3: one fp16 plane + exact refinement (parity grade);
    # 4: mode 3 plus the fp16 residual plane -- refinement by residual correction (half the gather bytes, ~1e-6 values)
    W_enc: Optional[torch.Tensor] = None  # modes 3 / 4 re-evaluate candidates against the fp32 parameter itself

    @staticmethod
    def pack(W_enc: torch.Tensor, b_enc: torch.Tensor, b_dec: torch.Tensor, planes: int = 2) -> "PackedEncoder":
        _need_cuda(W_enc, b_enc, b_dec)
        L = _capi.lib()
        W = W_enc.detach().to(torch.float32).contiguous()
        be = b_enc.detach().to(torch.float32).contiguous()
        bd = b_dec.detach().to(torch.float32).contiguous()
        N, d = W.shape
        nbytes = L.saeb_packed_weights_bytes(N, d, planes)
        blob = torch.empty(nbytes, dtype=torch.uint8, device=W.device)
        with torch.cuda.device(W.device):
            check(L.saeb_pack_weights(W.data_ptr(), be.data_ptr(), bd.data_ptr(), N, d, planes, blob.data_ptr(),
                                      _stream()), "saeb_pack_weights")
        return PackedEncoder(blob, N, d, planes, W if planes >= 3 else None)

    def folded_bias(self) -> torch.Tensor:
        off = _capi.lib().saeb_packed_bias_offset(self.num_latents, self.d_in, self.planes)
        return self.blob[off:off + 4 * self.num_latents].view(torch.float32)

    def plane(self, i: int) -> torch.Tensor:
        d_pad = (self.d_in + 7) // 8 * 8  # rows are padded to 16-byte multiples for TMA
        n = self.num_latents * d_pad * 2
        return self.blob[i * n:(i + 1) * n].view(torch.bfloat16).view(self.num_latents, d_pad)[:, :self.d_in]


def _as_2d(x: torch.Tensor, d: int) -> torch.Tensor:
    if x.shape[-1] != d:
        raise SaebError(f"last dimension {x.shape[-1]} != d_in {d}")
    x2 = x.reshape(-1, d)
    if x2.stride(-1) != 1 or (x2.shape[0] > 1 and x2.stride(0) % 8 != 0) or x2.data_ptr() % 16 != 0:
        x2 = x2.contiguous()
        if x2.data_ptr() % 16 != 0:  # contiguous() can return a view of an odd-offset storage
            x2 = x2.clone()
    return x2


def encode_topk(x: torch.Tensor, enc: PackedEncoder, k: int, *, clamp_feature: int = -1, clamp_value: float = 0.0,
                want_dense: bool = False, want_topk: bool = True, out_vals: Optional[torch.Tensor] = None,
                out_idx: Optional[torch.Tensor] = None, refine_margin: int = 0, value_mode: int = 0
                ) -> Tuple[Optional[torch.Tensor], Optional[torch.Tensor], Optional[torch.Tensor]]:
    """x [..., d] (bf16 / fp16 / fp32) -> (top_acts [..., k] f32, top_indices [..., k] i64, dense [..., N] f32 | None).
    Rows are ordered by (value desc, index asc).  value_mode (refine modes 3 / 4): VALUES_EXACT = every value exact
    fp32; VALUES_BOUNDARY = exact index set, exact values only where they decide the set (include/saeb200.h)."""
    _need_cuda(x, enc.blob)
    L = _capi.lib()
    lead = x.shape[:-1]
    x2 = _as_2d(x, enc.d_in)
    if x2.dtype not in _DT:
        x2 = x2.to(torch.float32)
    T = x2.shape[0]
    dev = x2.device
    if enc.planes >= 3 and want_dense:
        raise SaebError("dense pre-activations need a bf16 hi+lo packed encoder (planes=2), not the refine mode")
    vals = idx = None
    if want_topk:
        vals = out_vals if out_vals is not None else torch.empty((T, k), dtype=torch.float32, device=dev)
        idx = out_idx if out_idx is not None else torch.empty((T, k), dtype=torch.int64, device=dev)
        if (vals.shape != (T, k) or idx.shape != (T, k) or vals.dtype != torch.float32 or idx.dtype != torch.int64
                or not vals.is_contiguous() or not idx.is_contiguous()):
            raise SaebError("out_vals / out_idx must be contiguous [T, k] float32 / int64 tensors")
    dense = torch.empty((T, enc.num_latents), dtype=torch.float32, device=dev) if want_dense else None
    if T > 0 and enc.planes == 3:
        status = torch.zeros(1, dtype=torch.int32, device=dev)
        with torch.cuda.device(dev):
            nbytes = L.saeb_encode_topk_refine_workspace_bytes(T, enc.d_in, enc.num_latents, k, refine_margin)
            ws = _workspace(dev, nbytes)
            check(L.saeb_encode_topk_refine(x2.data_ptr(), _code(x2), T, x2.stride(0) if T > 1 else enc.d_in,
                                            enc.blob.data_ptr(), enc.W_enc.data_ptr(), enc.d_in, enc.num_latents, k,
                                            refine_margin,
                                            clamp_feature, float(clamp_value), vals.data_ptr(), idx.data_ptr(),
                                            status.data_ptr(), ws.data_ptr(), ws.numel(), int(value_mode), _stream()),
                  "saeb_encode_topk_refine")
        encode_topk.last_status = status
    elif T > 0 and enc.planes == 4:
        # the staged entry points, with the residual-correction refinement
        status = torch.zeros(1, dtype=torch.int32, device=dev)
        with torch.cuda.device(dev):
            prep = _workspace(dev, L.saeb_prep_bytes(T, enc.d_in), "prep4")
            ws = _workspace(dev, L.saeb_candidates_workspace_bytes(T, enc.d_in, enc.num_latents, k, refine_margin),
                            "cand4")
            st, code, ldx = _stream(), _code(x2), (x2.stride(0) if T > 1 else enc.d_in)
            check(L.saeb_prep_activations(x2.data_ptr(), code, T, ldx, enc.d_in, prep.data_ptr(), st),
                  "saeb_prep_activations")
            check(L.saeb_encode_candidates(prep.data_ptr(), T, 0, T, enc.blob.data_ptr(), enc.d_in, enc.num_latents, k,
                                           refine_margin, clamp_feature, float(clamp_value), ws.data_ptr(), ws.numel(),
                                           st), "saeb_encode_candidates")
            check(L.saeb_refine_candidates_lo(x2.data_ptr(), code, ldx, prep.data_ptr(), T, 0, T, enc.blob.data_ptr(),
                                              enc.W_enc.data_ptr(), enc.d_in, enc.num_latents, k, refine_margin,
                                              clamp_feature, float(clamp_value), None, None, None, 0, vals.data_ptr(),
                                              None, idx.data_ptr(), status.data_ptr(), ws.data_ptr(), ws.numel(), 0,
                                              int(value_mode), st),
                  "saeb_refine_candidates_lo")
        encode_topk.last_status = status
    elif T > 0:
        with torch.cuda.device(dev):
            nbytes = L.saeb_encode_topk_workspace_bytes(T, enc.d_in, enc.num_latents, k, _code(x2))
            ws = _workspace(dev, nbytes)
            check(L.saeb_encode_topk(x2.data_ptr(), _code(x2), T, x2.stride(0) if T > 1 else enc.d_in,
                                     enc.blob.data_ptr(), enc.planes, enc.d_in, enc.num_latents, k,
                                     clamp_feature, float(clamp_value),
                                     vals.data_ptr() if want_topk else None, idx.data_ptr() if want_topk else None,
                                     dense.data_ptr() if want_dense else None, enc.num_latents,
                                     ws.data_ptr(), ws.numel(), _stream()), "saeb_encode_topk")
    if want_topk:
        vals = vals.view(*lead, k)
        idx = idx.view(*lead, k)
    if want_dense:
        dense = dense.view(*lead, enc.num_latents)
    return vals, idx, dense


encode_topk.last_status = None  # device int: rows that went through the exact dense fallback in the last refine call


def dense_topk(latents: torch.Tensor, k: int) -> Tuple[torch.Tensor, torch.Tensor]:
    """TopK of a dense non-negative latent tensor [..., N] -> (values, int64 indices), (value desc, index asc)."""
    _need_cuda(latents)
    L = _capi.lib()
    N = latents.shape[-1]
    lead = latents.shape[:-1]
    x2 = latents.reshape(-1, N).to(torch.float32)
    if x2.stride(-1) != 1:
        x2 = x2.contiguous()
    T = x2.shape[0]
    vals = torch.empty((T, k), dtype=torch.float32, device=x2.device)
    idx = torch.empty((T, k), dtype=torch.int64, device=x2.device)
    with torch.cuda.device(x2.device):
        check(L.saeb_dense_topk(x2.data_ptr(), T, x2.stride(0) if T > 1 else N, N, k, vals.data_ptr(), idx.data_ptr(),
                                _stream()), "saeb_dense_topk")
    return vals.view(*lead, k), idx.view(*lead, k)


def decode(top_indices: torch.Tensor, top_acts: torch.Tensor, W_dec: torch.Tensor, b_dec: Optional[torch.Tensor],
           *, out_dtype: torch.dtype = torch.float32, x: Optional[torch.Tensor] = None,
           sq_err: Optional[torch.Tensor] = None, out: Optional[torch.Tensor] = None, max_ctas: int = 0
           ) -> torch.Tensor:
    """out[..., :] = sum_j acts[..., j] * W_dec[idx[..., j], :] + b_dec.  W_dec is the [N, d] parameter
    (fp32 parity grade, or a bf16 copy).  If `x` and `sq_err` (0-dim float64) are given, sum((out-x)^2) is added
    to sq_err.  max_ctas > 0: persistent grid of that many CTAs (rides beside a resident GEMM grid; same results)."""
    _need_cuda(top_indices, top_acts, W_dec, b_dec, x, sq_err)
    L = _capi.lib()
    N, d = W_dec.shape
    if not W_dec.is_contiguous():
        raise SaebError("W_dec must be the contiguous [N, d] parameter (not the transposed view)")
    lead = top_indices.shape[:-1]
    k = top_indices.shape[-1]
    idx = top_indices.reshape(-1, k).to(torch.int64).contiguous()
    vals = top_acts.reshape(-1, k).to(torch.float32).contiguous()
    T = idx.shape[0]
    if out is None:
        out = torch.empty((T, d), dtype=out_dtype, device=idx.device)
    elif out.shape != (T, d) or not out.is_contiguous() or out.dtype not in _DT:
        raise SaebError("decode: `out` must be a contiguous [T, d] tensor")
    else:
        out_dtype = out.dtype
    if T == 0:
        return out.view(*lead, d)
    bd = None if b_dec is None else b_dec.detach().to(torch.float32).contiguous()
    x2 = None
    if x is not None and sq_err is not None:
        x2 = _as_2d(x, d)
        if x2.dtype not in _DT:
            x2 = x2.to(torch.float32)
        if x2.stride(0) % 4 != 0 and T > 1:
            x2 = x2.contiguous()
    err_flag = torch.zeros(1, dtype=torch.int32, device=idx.device)
    with torch.cuda.device(idx.device):
        check(L.saeb_decode(idx.data_ptr(), vals.data_ptr(), T, k, W_dec.data_ptr(), _code(W_dec), d, N,
                            None if bd is None else bd.data_ptr(), out.data_ptr(), _DT[out_dtype], d,
                            None if x2 is None else x2.data_ptr(), 0 if x2 is None else _code(x2),
                            0 if x2 is None else (x2.stride(0) if T > 1 else d),
                            None if (sq_err is None or x2 is None) else sq_err.data_ptr(), err_flag.data_ptr(),
                            int(max_ctas), _stream()), "saeb_decode")
    decode.last_err_flag = err_flag
    if DEBUG_CHECKS and int(err_flag.item()) != 0:   # synchronises: debugging aid (SAEB_DEBUG_CHECKS=1)
        raise SaebError("decode: a TopK index is outside [0, num_latents) (the reference's tl.device_assert)")
    return out.view(*lead, d)


decode.last_err_flag = None


def decode_backward(top_indices: torch.Tensor, top_acts: torch.Tensor, W_dec: torch.Tensor, grad_out: torch.Tensor,
                    *, need_acts: bool = True, need_weight: bool = True
                    ) -> Tuple[Optional[torch.Tensor], Optional[torch.Tensor]]:
    """Backward products of `decode` (reference TritonDecoder.backward, sae/kernels.py:411-429), fp32:
    d_acts[..., j] = grad_out[..., :] . W_dec[idx[..., j], :]  and  dW_dec [N, d] = sparse(acts)^T @ grad_out."""
    _need_cuda(top_indices, top_acts, W_dec, grad_out)
    L = _capi.lib()
    N, d = W_dec.shape
    if W_dec.dtype != torch.float32 or not W_dec.is_contiguous():
        raise SaebError("decode_backward: W_dec must be the contiguous fp32 [N, d] parameter")
    k = top_indices.shape[-1]
    idx = top_indices.reshape(-1, k).to(torch.int64).contiguous()
    vals = top_acts.reshape(-1, k).to(torch.float32).contiguous()
    g = grad_out.reshape(-1, d).to(torch.float32).contiguous()
    T = idx.shape[0]
    if g.shape[0] != T:
        raise SaebError(f"decode_backward: grad_out has {g.shape[0]} rows, the TopK tensors {T}")
    err_flag = torch.zeros(1, dtype=torch.int32, device=idx.device)
    d_acts = dW = None
    with torch.cuda.device(idx.device):
        if need_acts:
            d_acts = torch.empty((T, k), dtype=torch.float32, device=idx.device)
            check(L.saeb_decode_backward_acts(g.data_ptr(), d, idx.data_ptr(), T, k, W_dec.data_ptr(), d, N,
                                              d_acts.data_ptr(), err_flag.data_ptr(), _stream()),
                  "saeb_decode_backward_acts")
            d_acts = d_acts.view(*top_indices.shape)
        if need_weight:
            dW = torch.zeros((N, d), dtype=torch.float32, device=idx.device)
            check(L.saeb_decode_backward_weight(g.data_ptr(), d, idx.data_ptr(), vals.data_ptr(), T, k, d, N,
                                                dW.data_ptr(), err_flag.data_ptr(), _stream()),
                  "saeb_decode_backward_weight")
    decode_backward.last_err_flag = err_flag
    return d_acts, dW


decode_backward.last_err_flag = None


def total_variance(x: torch.Tensor) -> torch.Tensor:
    """sum((x - x.mean(0))**2) as a 0-dim float64 tensor (reference sae/sae.py:204)."""
    _need_cuda(x)
    L = _capi.lib()
    d = x.shape[-1]
    x2 = _as_2d(x, d)
    if x2.dtype not in _DT:
        x2 = x2.to(torch.float32)
    T = x2.shape[0]
    scratch = torch.empty(2 * d, dtype=torch.float64, device=x2.device)
    out = torch.zeros((), dtype=torch.float64, device=x2.device)
    with torch.cuda.device(x2.device):
        check(L.saeb_total_variance(x2.data_ptr(), _code(x2), T, d, x2.stride(0) if T > 1 else d,
                                    scratch.data_ptr(), out.data_ptr(), _stream()), "saeb_total_variance")
    return out


def make_filter_bitmap(features: torch.Tensor, num_latents: int) -> torch.Tensor:
    """Feature-id filter (reference `torch.isin(feature, filters[module])`, features/cache.py:89-92) as N bits."""
    words = (num_latents + 31) // 32
    f = features.to(torch.int64).flatten()
    bits = torch.zeros(words * 32, dtype=torch.int64, device=f.device)
    bits[f] = 1
    w = (bits.view(words, 32) << torch.arange(32, device=f.device, dtype=torch.int64)).sum(1)
    return _pack_words(w)


def _pack_words(w64: torch.Tensor) -> torch.Tensor:
    # values in [0, 2^32): store as int32 with wraparound (bit pattern is what the kernel reads)
    w = w64.clone()
    w[w >= 2 ** 31] -= 2 ** 32
    return w.to(torch.int32).contiguous()


def coo_extract(top_acts: torch.Tensor, top_indices: torch.Tensor, seq_len: int, *, row_offset: int = 0,
                threshold: float = ACT_THRESHOLD, filter_bitmap: Optional[torch.Tensor] = None
                ) -> Tuple[torch.Tensor, torch.Tensor]:
    """TopK output of batch*seq_len tokens -> (locations [nnz,3] i64, activations [nnz] f32) in the reference's
    `torch.nonzero` order (features/cache.py:73-92)."""
    _need_cuda(top_acts, top_indices, filter_bitmap)
    L = _capi.lib()
    k = top_acts.shape[-1]
    vals = top_acts.reshape(-1, k).to(torch.float32).contiguous()
    idx = top_indices.reshape(-1, k).to(torch.int64).contiguous()
    T = vals.shape[0]
    dev = vals.device
    loc = torch.empty((T * k, 3), dtype=torch.int64, device=dev)
    act = torch.empty((T * k,), dtype=torch.float32, device=dev)
    if T == 0:
        return loc, act
    nnz = torch.zeros(1, dtype=torch.int64, device=dev)
    with torch.cuda.device(dev):
        ws = _workspace(dev, L.saeb_coo_workspace_bytes(T), "coo")
        check(L.saeb_coo_extract(vals.data_ptr(), idx.data_ptr(), T, k, float(threshold),
                                 None if filter_bitmap is None else filter_bitmap.data_ptr(), seq_len, row_offset,
                                 loc.data_ptr(), act.data_ptr(), nnz.data_ptr(), ws.data_ptr(), ws.numel(),
                                 _stream()), "saeb_coo_extract")
    n = int(nnz.item())  # the reference synchronises here too (.cpu(), features/cache.py:52-53)
    return loc[:n], act[:n]


class TopActivationScan:
    """Per-feature `n_top` best windows by max TopK-masked activation, for the features [feat_lo, feat_hi) owned by
    this GPU.  Lists live on the device (F * n_top * 12 bytes); tokens are fed in chunks with `update`."""

    def __init__(self, feat_lo: int, feat_hi: int, n_top: int, ctx_len: int, device, *, bucket_cap: int = 256,
                 threshold: float = ACT_THRESHOLD):
        self.feat_lo, self.feat_hi = int(feat_lo), int(feat_hi)
        self.F = self.feat_hi - self.feat_lo
        self.n_top, self.ctx_len, self.bucket_cap, self.threshold = n_top, ctx_len, bucket_cap, float(threshold)
        dev = torch.device(device)
        if dev.type != "cuda":
            raise SaebError("TopActivationScan needs a CUDA device: there is no CPU fallback")
        self.device = dev
        self.top_vals = torch.zeros((self.F, n_top), dtype=torch.float32, device=dev)
        self.top_win = torch.full((self.F, n_top), -1, dtype=torch.int64, device=dev)
        self.feat_thr = torch.full((self.F,), self.threshold, dtype=torch.float32, device=dev)
        self.bucket = torch.empty((self.F, bucket_cap, 2), dtype=torch.int32, device=dev)
        self.bucket_cnt = torch.zeros((self.F,), dtype=torch.int32, device=dev)
        self.overflow = torch.zeros(1, dtype=torch.int32, device=dev)
        self._pending = 0
        # True: the pooling kernel keeps its hash tables in global scratch (saeb_scan_pool_ws: no shared memory, runs
        # beside a resident GEMM CTA) -- set by the pipelined scans, whose list update overlaps the next chunk's GEMM
        self.coresident = False
        self._pool_ws, self._pool_ws_k = None, None

    def _pool_workspace(self, k: int) -> torch.Tensor:
        if self._pool_ws is None or self._pool_ws_k != k:
            L = _capi.lib()
            with torch.cuda.device(self.device):
                nbytes = L.saeb_scan_pool_workspace_bytes(k, self.ctx_len)
                self._pool_ws = torch.empty(nbytes, dtype=torch.uint8, device=self.device)
                check(L.saeb_scan_pool_init(self._pool_ws.data_ptr(), nbytes, k, self.ctx_len, _stream()),
                      "saeb_scan_pool_init")
            self._pool_ws_k = k
        return self._pool_ws

    def update(self, top_acts: torch.Tensor, top_indices: torch.Tensor, window_base: int,
               tok_thr: Optional[torch.Tensor] = None, member: Optional[torch.Tensor] = None, idx_base: int = 0) -> None:
        """Feed TopK output of T tokens (T a multiple of ctx_len except for the very last chunk) whose first window
        has global id `window_base`.  tok_thr [T] / member [T, k] (optional): membership threshold per token and the
        values it is compared with (feature-sharded scan, see saeb_scan_pool).  `idx_base`: `top_indices` are relative
        to that feature id (a shard's refinement emits shard-local ids; no kernel just to add the offset)."""
        L = _capi.lib()
        k = top_acts.shape[-1]
        vals = top_acts.reshape(-1, k)
        idx = top_indices.reshape(-1, k)
        mem = None if member is None else member.reshape(-1, k)
        for t_, dt_ in ((vals, torch.float32), (idx, torch.int64), (mem, torch.float32), (tok_thr, torch.float32)):
            if t_ is not None and (t_.dtype != dt_ or not t_.is_contiguous() or t_.device != self.device):
                raise SaebError("scan update needs contiguous float32 values / int64 indices on the scan's device")
        if window_base < 0 or window_base + (vals.shape[0] + self.ctx_len - 1) // self.ctx_len >= 2 ** 32:
            raise SaebError("scan update: window ids must fit 32 bits")
        T = vals.shape[0]
        max_tok = self.bucket_cap * self.ctx_len
        with torch.cuda.device(self.device):
            for t0 in range(0, T, max_tok):
                t1 = min(T, t0 + max_tok)
                n_win = (t1 - t0 + self.ctx_len - 1) // self.ctx_len
                if self._pending + n_win > self.bucket_cap:
                    self.flush()
                v, i = vals[t0:t1], idx[t0:t1]
                args = (v.data_ptr(), i.data_ptr(), t1 - t0, k, self.ctx_len, self.threshold,
                        self.feat_lo - int(idx_base), self.feat_hi - int(idx_base), window_base + t0 // self.ctx_len,
                        None if tok_thr is None else tok_thr[t0:t1].data_ptr(),
                        None if mem is None else mem[t0:t1].data_ptr(),
                        self.feat_thr.data_ptr(), self.bucket.data_ptr(), self.bucket_cnt.data_ptr(),
                        self.bucket_cap, self.overflow.data_ptr())
                if self.coresident:
                    ws = self._pool_workspace(k)
                    check(L.saeb_scan_pool_ws(*args, ws.data_ptr(), ws.numel(), _stream()), "saeb_scan_pool_ws")
                else:
                    check(L.saeb_scan_pool(*args, _stream()), "saeb_scan_pool")
                self._pending += n_win

    def flush(self) -> None:
        if self._pending == 0:
            return
        L = _capi.lib()
        with torch.cuda.device(self.device):
            check(L.saeb_scan_merge(self.bucket.data_ptr(), self.bucket_cnt.data_ptr(), self.bucket_cap, self.F,
                                    self.n_top, self.threshold, self.top_vals.data_ptr(), self.top_win.data_ptr(),
                                    self.feat_thr.data_ptr(), _stream()), "saeb_scan_merge")
        self._pending = 0

    def finalize(self) -> Tuple[torch.Tensor, torch.Tensor]:
        self.flush()
        if int(self.overflow.item()) != 0:
            raise SaebError("top-activation scan: a feature's bucket overflowed (more than bucket_cap windows between "
                            "two flushes)")
        return self.top_vals, self.top_win


class TopImageScan(TopActivationScan):
    """Image form of the scan: per feature the `n_top` images with the largest MEAN TopK-masked activation over the first
    `n_base` positions of their token row (reference pool_max_activations_windows_image, features/constructors.py:
    88-148, for every feature of the shard at once).  `finalize()` -> (scores [F, n_top], image ids [F, n_top], -1 =
    empty), ordered (score desc, image id asc)."""

    def __init__(self, feat_lo: int, feat_hi: int, n_top: int, tokens_per_image: int, n_base: int, device, *,
                 bucket_cap: int = 256, threshold: float = ACT_THRESHOLD):
        super().__init__(feat_lo, feat_hi, n_top, tokens_per_image, device, bucket_cap=bucket_cap, threshold=threshold)
        if not 1 <= n_base <= tokens_per_image:
            raise SaebError("TopImageScan: need 1 <= n_base <= tokens_per_image")
        self.tokens_per_image, self.n_base = int(tokens_per_image), int(n_base)
        self._ws, self._ws_k = None, None

    def _workspace(self, k: int) -> torch.Tensor:
        if self._ws is None or self._ws_k != k:
            L = _capi.lib()
            with torch.cuda.device(self.device):
                nbytes = L.saeb_image_pool_workspace_bytes(k, self.n_base, self.F)
                self._ws = torch.empty(nbytes, dtype=torch.uint8, device=self.device)
                check(L.saeb_image_pool_init(self._ws.data_ptr(), nbytes, k, self.n_base, self.F, _stream()),
                      "saeb_image_pool_init")
            self._ws_k = k
        return self._ws

    def update(self, top_acts: torch.Tensor, top_indices: torch.Tensor, image_base: int,
               tok_thr: Optional[torch.Tensor] = None, member: Optional[torch.Tensor] = None, idx_base: int = 0) -> None:
        """Feed the TopK output of whole image rows ([n_images * tokens_per_image, k]); the first row is image
        `image_base`.  tok_thr [T] (feature-sharded scan): per-token membership threshold, compared with the activations
        themselves -- a MEAN needs every member's exact value, so the shards must refine with every member exact
        (`EngineOps.scan_value_mode = 0`); separate member values are refused.  `idx_base` as in TopActivationScan."""
        if member is not None and member.data_ptr() != top_acts.data_ptr():
            raise SaebError("TopImageScan averages activations: run the shards with every member exact "
                            "(scan_value_mode = 0), not with the scan mode's member values")
        L = _capi.lib()
        k = top_acts.shape[-1]
        vals = top_acts.reshape(-1, k)
        idx = top_indices.reshape(-1, k)
        tpi = self.tokens_per_image
        if vals.shape[0] % tpi != 0:
            raise SaebError("TopImageScan.update needs whole image rows")
        n_img = vals.shape[0] // tpi
        ws = self._workspace(k)
        with torch.cuda.device(self.device):
            for i0 in range(0, n_img, self.bucket_cap):
                i1 = min(n_img, i0 + self.bucket_cap)
                if self._pending + (i1 - i0) > self.bucket_cap:
                    self.flush()
                v, i = vals[i0 * tpi:i1 * tpi], idx[i0 * tpi:i1 * tpi]
                check(L.saeb_image_pool(v.data_ptr(), i.data_ptr(), i1 - i0, tpi, k, self.n_base, self.threshold,
                                        self.feat_lo - int(idx_base), self.feat_hi - int(idx_base), image_base + i0,
                                        None if tok_thr is None else tok_thr[i0 * tpi:i1 * tpi].data_ptr(),
                                        self.feat_thr.data_ptr(), self.bucket.data_ptr(), self.bucket_cnt.data_ptr(),
                                        self.bucket_cap, self.overflow.data_ptr(), ws.data_ptr(), ws.numel(),
                                        _stream()), "saeb_image_pool")
                self._pending += i1 - i0


def kth_of_gathered(gathered: torch.Tensor, kth: Optional[int] = None) -> torch.Tensor:
    """gathered [R, T, m] f32 (all-gathered per-shard value lists, m values per shard and token) -> per-token
    kth-largest of the R*m values [T] (values <= 0 count as 0).  kth defaults to m: the per-token global k-th value from
    the shards' local top-k values."""
    _need_cuda(gathered)
    L = _capi.lib()
    R, T, m = gathered.shape
    kth = m if kth is None else int(kth)
    g = gathered.contiguous()
    if g.dtype != torch.float32:
        raise SaebError("kth_of_gathered needs float32 values")
    out = torch.empty((T,), dtype=torch.float32, device=g.device)
    with torch.cuda.device(g.device):
        check(L.saeb_kth_largest_gathered(g.data_ptr(), R, T, m, kth, out.data_ptr(), _stream()),
              "saeb_kth_largest_gathered")
    return out


def gathered_bounds(gathered: torch.Tensor, m1: int, k: int) -> Tuple[torch.Tensor, torch.Tensor]:
    """gathered [R, T, 2*m1] f32 (exchange 1 of the feature-sharded scan: per shard and token the m1 largest lower
    bounds | the m1 largest upper bounds) -> (ext_L [T], ext_U [T]): the k-th largest lower bound and an upper bound of
    the (k+1)-th largest upper bound over all latents (saeb_gathered_bounds)."""
    _need_cuda(gathered)
    L = _capi.lib()
    R, T, w = gathered.shape
    if w != 2 * m1 or gathered.dtype != torch.float32 or not gathered.is_contiguous():
        raise SaebError("gathered_bounds needs a contiguous float32 [R, T, 2*m1] tensor")
    ext_L = torch.empty((T,), dtype=torch.float32, device=gathered.device)
    ext_U = torch.empty((T,), dtype=torch.float32, device=gathered.device)
    with torch.cuda.device(gathered.device):
        check(L.saeb_gathered_bounds(gathered.data_ptr(), R, T, m1, int(k), ext_L.data_ptr(), ext_U.data_ptr(),
                                     _stream()), "saeb_gathered_bounds")
    return ext_L, ext_U


def mean_activations(x: torch.Tensor, enc_dense: PackedEncoder, *, chunk_tokens: int = 2048) -> torch.Tensor:
    """mean over tokens of the dense latents relu(W_enc (x - b_dec) + b_enc), [N] fp32 -- `sae.pre_acts(x).mean(0)` of
    the reference's probing tool (tools/probe_activations.py:109-121) without the [T, N] tensor: token chunks go through
    the fused GEMM's dense store into one scratch buffer and are reduced by saeb_column_sums (fp64 accumulation)."""
    _need_cuda(x, enc_dense.blob)
    if enc_dense.planes != 2:
        raise SaebError("mean_activations needs the bf16 hi+lo packed encoder (planes=2): it is the one with a dense store")
    L = _capi.lib()
    x2 = _as_2d(x, enc_dense.d_in)
    T, N = x2.shape[0], enc_dense.num_latents
    colsum = torch.zeros(N, dtype=torch.float64, device=x2.device)
    with torch.cuda.device(x2.device):
        for t0 in range(0, T, chunk_tokens):
            xc = x2[t0:t0 + chunk_tokens]
            _, _, dense = encode_topk(xc, enc_dense, 1, want_dense=True, want_topk=False)
            check(L.saeb_column_sums(dense.data_ptr(), xc.shape[0], N, N, colsum.data_ptr(), _stream()),
                  "saeb_column_sums")
    return (colsum / max(T, 1)).to(torch.float32)


def feature_maps(x: torch.Tensor, W_enc: torch.Tensor, b_enc: torch.Tensor, b_dec: torch.Tensor,
                 features: torch.Tensor) -> torch.Tensor:
    """maps[j, t] = relu((x_t - b_dec) . W_enc[features[j]] + b_enc[features[j]]), exact fp32: the columns
    `latents[:, features]` of the dense latents (tools/probe_activations.py:122) for a few selected features."""
    _need_cuda(x, W_enc, b_enc, b_dec, features)
    L = _capi.lib()
    N, d = W_enc.shape
    x2 = _as_2d(x, d)
    if x2.dtype not in _DT:
        x2 = x2.to(torch.float32)
    T = x2.shape[0]
    W = W_enc.detach().to(torch.float32).contiguous()
    sel = features.to(torch.int64).contiguous()
    out = torch.empty((sel.numel(), T), dtype=torch.float32, device=x2.device)
    err = torch.zeros(1, dtype=torch.int32, device=x2.device)
    with torch.cuda.device(x2.device):
        check(L.saeb_feature_maps(x2.data_ptr(), _code(x2), T, x2.stride(0) if T > 1 else d, W.data_ptr(),
                                  b_enc.detach().to(torch.float32).contiguous().data_ptr(),
                                  b_dec.detach().to(torch.float32).contiguous().data_ptr(), d, N, sel.data_ptr(),
                                  sel.numel(), out.data_ptr(), err.data_ptr(), _stream()), "saeb_feature_maps")
    if int(err.item()) != 0:
        raise SaebError("feature_maps: feature id out of range")
    return out


def coo_top_windows(locations: torch.Tensor, activations: torch.Tensor, n_top: int, *, ctx_len: Optional[int] = None,
                    seq_len: Optional[int] = None, n_base: Optional[int] = None, device=None
                    ) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor, torch.Tensor]:
    """Ranking step of the example constructors for EVERY feature of a split file in one pass on the device.

    locations [nnz, 3] int64 (row, pos, feature) / activations [nnz] f32 as stored in the cache's split files.
      * text windows (`ctx_len`, `seq_len` given; reference pool_max_activation_windows, features/constructors.py:
        11-85): score(feature, row, w) = max activation over positions [w * ctx_len, (w + 1) * ctx_len), windows beyond
        seq_len // ctx_len are dropped like the reference's max_pool1d does; window id = row * n_win + w;
      * images (`n_base` given; pool_max_activations_windows_image, :104-122): score(feature, row) = sum of the
        activations at positions < n_base divided by n_base (avg_pool1d over the base image tokens); window id = row.
    Returns (features [F'] sorted ascending, offsets [F' + 1], scores [M], window ids [M]): for features[i] the entries
    offsets[i]:offsets[i+1] are its best <= n_top windows ordered (score desc, window id asc).  Tensors live on the
    device.  The triples are grouped with one stable device sort, the pooled scores come from saeb_coo_window_scores
    (file-order reduction, so image means equal the reference's sequential CPU sums bit for bit), the ranking is one
    more device sort of (feature, score) keys."""
    text = ctx_len is not None
    if text == (n_base is not None):
        raise SaebError("coo_top_windows: give either ctx_len + seq_len (text windows) or n_base (images)")
    dev = torch.device(device) if device is not None else locations.device
    if dev.type != "cuda":
        raise SaebError("coo_top_windows needs a CUDA device: there is no CPU fallback")
    L = _capi.lib()
    loc = locations.to(dev, non_blocking=True)
    act = activations.to(dev, torch.float32, non_blocking=True)
    nnz = loc.shape[0]
    empty = torch.zeros(0, dtype=torch.int64, device=dev)
    if nnz == 0:
        return empty, torch.zeros(1, dtype=torch.int64, device=dev), torch.zeros(0, device=dev), empty
    row, pos, feat = loc[:, 0], loc[:, 1], loc[:, 2]
    if text:
        n_win = int(seq_len) // int(ctx_len)
        key = row * n_win + pos // ctx_len
        valid = pos < n_win * ctx_len
        mode, divisor = 0, 1.0
    else:
        key = row
        valid = pos < n_base
        mode, divisor = 1, float(n_base)
    feat, key, acts = feat[valid], key[valid], act[valid]
    nnz = feat.numel()
    if nnz == 0:
        return empty, torch.zeros(1, dtype=torch.int64, device=dev), torch.zeros(0, device=dev), empty
    # group by (feature, window) with ONE stable sort: file order is kept inside a window, so the sums below add in
    # the order the reference's CPU ops do
    span = int(key.max().item()) + 1
    order = torch.sort(feat * span + key, stable=True).indices
    feat, key, acts = feat[order].contiguous(), key[order].contiguous(), acts[order].contiguous()
    score = torch.empty(nnz, dtype=torch.float32, device=dev)
    head = torch.empty(nnz, dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        check(L.saeb_coo_window_scores(feat.data_ptr(), key.data_ptr(), acts.data_ptr(), nnz, mode, divisor,
                                       score.data_ptr(), head.data_ptr(), _stream()), "saeb_coo_window_scores")
    sel = head.nonzero().squeeze(1)
    f_h, k_h, s_h = feat[sel], key[sel], score[sel]
    keep = s_h > 0   # the constructors rank non-zero pooled values only (constructors.py:58-60)
    f_h, k_h, s_h = f_h[keep], k_h[keep], s_h[keep]
    # (feature asc, score desc) in one stable sort; window ids are already ascending inside a feature
    comp = (f_h << 32) | (0xFFFFFFFF - s_h.view(torch.int32).to(torch.int64))
    o2 = torch.sort(comp, stable=True).indices
    f_s, k_s, s_s = f_h[o2], k_h[o2], s_h[o2]
    feats, counts = torch.unique_consecutive(f_s, return_counts=True)
    starts = torch.cumsum(counts, 0) - counts
    rank = torch.arange(f_s.numel(), device=dev) - torch.repeat_interleave(starts, counts)
    top = rank < n_top
    kept = torch.clamp(counts, max=n_top)
    offsets = torch.cat([torch.zeros(1, dtype=torch.int64, device=dev), torch.cumsum(kept, 0)])
    return feats, offsets, s_s[top], k_s[top]


class CooArena:
    """Device-resident accumulation of the activation cache of one hooked module (reference Cache.add / Cache.save,
    features/cache.py:42-71, which copies every batch to the host): `append` enqueues one extraction that writes behind
    a device-side cursor, nothing comes back to the host until `tensors()`.

    The host only tracks an upper bound of the fill (T * k per batch); when that bound reaches the capacity it reads
    the true cursor once (the only synchronisation besides the final one) and grows the arena geometrically if the data
    really do not fit."""

    def __init__(self, device, capacity: int = 1 << 22):
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise SaebError("CooArena needs a CUDA device: there is no CPU fallback")
        self.capacity = int(capacity)
        self.loc = torch.empty((self.capacity, 3), dtype=torch.int64, device=self.device)
        self.act = torch.empty((self.capacity,), dtype=torch.float32, device=self.device)
        self.cursor = torch.zeros(1, dtype=torch.int64, device=self.device)
        self.overflow = torch.zeros(1, dtype=torch.int32, device=self.device)
        self.upper = 0        # host-side upper bound of the cursor
        self.syncs = 0        # host round trips so far (diagnostics / tests)

    def _reserve(self, extra: int) -> None:
        if self.upper + extra <= self.capacity:
            return
        n = int(self.cursor.item())
        self.syncs += 1
        self.upper = n
        if n + extra > self.capacity:
            cap = max(2 * self.capacity, n + extra)
            loc = torch.empty((cap, 3), dtype=torch.int64, device=self.device)
            act = torch.empty((cap,), dtype=torch.float32, device=self.device)
            loc[:n].copy_(self.loc[:n])
            act[:n].copy_(self.act[:n])
            self.loc, self.act, self.capacity = loc, act, cap

    def append(self, top_acts: torch.Tensor, top_indices: torch.Tensor, seq_len: int, *, row_offset: int = 0,
               threshold: float = ACT_THRESHOLD, filter_bitmap: Optional[torch.Tensor] = None) -> None:
        _need_cuda(top_acts, top_indices, filter_bitmap)
        L = _capi.lib()
        k = top_acts.shape[-1]
        vals = top_acts.reshape(-1, k).to(torch.float32).contiguous()
        idx = top_indices.reshape(-1, k).to(torch.int64).contiguous()
        T = vals.shape[0]
        if T == 0:
            return
        self._reserve(T * k)
        with torch.cuda.device(self.device):
            ws = _workspace(self.device, L.saeb_coo_workspace_bytes(T) + 256, "coo_append")
            check(L.saeb_coo_append(vals.data_ptr(), idx.data_ptr(), T, k, float(threshold),
                                    None if filter_bitmap is None else filter_bitmap.data_ptr(), seq_len, row_offset,
                                    self.loc.data_ptr(), self.act.data_ptr(), self.capacity, self.cursor.data_ptr(),
                                    self.overflow.data_ptr(), ws.data_ptr(), ws.numel(), _stream()), "saeb_coo_append")
        self.upper += T * k

    def tensors(self) -> Tuple[torch.Tensor, torch.Tensor]:
        """(locations [nnz, 3], activations [nnz]) views of the arena on the device (one synchronisation)"""
        n = int(self.cursor.item())
        self.syncs += 1
        if int(self.overflow.item()) != 0:
            raise SaebError("CooArena overflowed (internal error: the reserve logic must prevent this)")
        return self.loc[:n], self.act[:n]
