"""The reference's decoder seam (sae/utils.py:107-129): a module-level `decoder_impl(top_indices, top_acts, W_dec_T)`.

Here the seam is bound to the CUDA gather-decode of `saeb200` -- there is no Triton and no eager fallback; CPU tensors
raise.  Like the reference's `TritonDecoder` (sae/kernels.py:403-429) it is a `torch.autograd.Function`: gradients
reach `top_acts` (gathered dot products) and `W_dec` (sparse^T @ dense) through the library's backward kernels."""
from __future__ import annotations

import torch
from torch import Tensor

from saeb200 import engine


def _weight(W_dec_T: Tensor) -> Tensor:
    """`W_dec.mT` ([d, N] view) -> the contiguous [N, d] parameter the kernels gather rows from"""
    W = W_dec_T.mT
    return W if W.is_contiguous() else W.contiguous()


def _out_dtype(top_acts: Tensor) -> torch.dtype:
    return top_acts.dtype if top_acts.dtype in (torch.float32, torch.float16, torch.bfloat16) else torch.float32


class SparseDecode(torch.autograd.Function):
    """forward(ctx, top_indices, top_acts, W_dec_T) / backward, the contract of the reference's TritonDecoder."""

    @staticmethod
    def forward(ctx, top_indices: Tensor, top_acts: Tensor, W_dec_T: Tensor) -> Tensor:
        ctx.save_for_backward(top_indices, top_acts, W_dec_T)
        return engine.decode(top_indices, top_acts, _weight(W_dec_T.detach()), None, out_dtype=_out_dtype(top_acts))

    @staticmethod
    def backward(ctx, grad_output: Tensor):
        top_indices, top_acts, W_dec_T = ctx.saved_tensors
        need_acts, need_w = ctx.needs_input_grad[1], ctx.needs_input_grad[2]
        W = _weight(W_dec_T.detach()).to(torch.float32)
        d_acts, dW = engine.decode_backward(top_indices, top_acts.detach(), W, grad_output, need_acts=need_acts,
                                            need_weight=need_w)
        if d_acts is not None:
            d_acts = d_acts.to(top_acts.dtype)
        # W_dec_T is the transposed view of the [N, d] parameter: the matching gradient layout is dW.mT
        return None, d_acts, None if dW is None else dW.to(W_dec_T.dtype).mT


def cuda_decode(top_indices: Tensor, top_acts: Tensor, W_dec_T: Tensor) -> Tensor:
    """Same contract as the reference's `triton_decode` / `eager_decode`: `W_dec_T` is `W_dec.mT`, a [d, N] view of
    the contiguous [N, d] parameter; returns sum_j acts[..., j] * W_dec[idx[..., j], :] (no bias)."""
    if torch.is_grad_enabled() and (top_acts.requires_grad or W_dec_T.requires_grad):
        return SparseDecode.apply(top_indices, top_acts, W_dec_T)
    return engine.decode(top_indices, top_acts, _weight(W_dec_T), None, out_dtype=_out_dtype(top_acts))


decoder_impl = cuda_decode
