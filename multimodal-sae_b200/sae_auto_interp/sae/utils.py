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


class SparseEncode(torch.autograd.Function):
    """Fused encode + TopK with the gradient the reference gets from `nn.Linear -> relu -> topk` (sae/sae.py:172-185):
    the selection is a constant of the backward pass, gradients flow through the k selected latents only,
        d x      = sum_j g_j W_enc[idx_j]                 (the decode gather with W_enc as the row table)
        d W_enc  = sparse(g)^T (x - b_dec),  d b_enc[n] = sum of g over the entries that selected n,  d b_dec = -sum_t dx
    with g = grad(top_acts) masked where the latent is not positive (relu) or was clamped (a constant).  Without this
    an SAE spliced into the host model blocks the gradient to everything upstream of it (attribution patching over
    several layers, features/patching/attribution.py)."""

    @staticmethod
    def forward(ctx, x: Tensor, W_enc: Tensor, b_enc: Tensor, b_dec: Tensor, enc, k: int, clamp_feature: int,
                clamp_value: float, value_mode: int):
        acts, idx, _ = engine.encode_topk(x.detach(), enc, k, clamp_feature=clamp_feature, clamp_value=clamp_value,
                                          value_mode=value_mode)
        ctx.save_for_backward(x, idx, acts, W_enc, b_enc, b_dec)
        ctx.clamp_feature = clamp_feature
        ctx.mark_non_differentiable(idx)
        return acts, idx

    @staticmethod
    def backward(ctx, grad_acts: Tensor, _grad_idx):
        x, idx, acts, W_enc, b_enc, b_dec = ctx.saved_tensors
        k = idx.shape[-1]
        g = grad_acts.reshape(-1, k).to(torch.float32)
        live = acts.reshape(-1, k) > 0
        if ctx.clamp_feature >= 0:
            live = live & (idx.reshape(-1, k) != ctx.clamp_feature)
        g = torch.where(live, g, torch.zeros_like(g)).contiguous()
        idx2 = idx.reshape(-1, k)
        W = W_enc.detach().to(torch.float32).contiguous()
        need_x, need_w, need_be, need_bd = ctx.needs_input_grad[:4]
        dx = dW = dbe = dbd = None
        if need_x or need_bd:
            dx_flat = engine.decode(idx2, g, W, None)                       # [T, d] fp32
            if need_bd:
                dbd = (-dx_flat.sum(0)).to(b_dec.dtype)
            if need_x:
                dx = dx_flat.view(x.shape).to(x.dtype)
        if need_w:
            centred = (x.detach().reshape(-1, x.shape[-1]).to(torch.float32) - b_dec.detach().to(torch.float32))
            _, dW = engine.decode_backward(idx2, g, W, centred, need_acts=False, need_weight=True)
            dW = dW.to(W_enc.dtype)
        if need_be:
            dbe = torch.zeros(W.shape[0], dtype=torch.float32, device=g.device).index_add_(0, idx2.reshape(-1), g.reshape(-1))
            dbe = dbe.to(b_enc.dtype)
        return dx, dW, dbe, dbd, None, None, None, None, None


def cuda_decode(top_indices: Tensor, top_acts: Tensor, W_dec_T: Tensor) -> Tensor:
    """Same contract as the reference's `triton_decode` / `eager_decode`: `W_dec_T` is `W_dec.mT`, a [d, N] view of
    the contiguous [N, d] parameter; returns sum_j acts[..., j] * W_dec[idx[..., j], :] (no bias)."""
    if torch.is_grad_enabled() and (top_acts.requires_grad or W_dec_T.requires_grad):
        return SparseDecode.apply(top_indices, top_acts, W_dec_T)
    return engine.decode(top_indices, top_acts, _weight(W_dec_T), None, out_dtype=_out_dtype(top_acts))


decoder_impl = cuda_decode
