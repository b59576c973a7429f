"""The reference's decoder seam (sae/utils.py:107-129): a module-level `decoder_impl(top_indices, top_acts, W_dec_T)`.

Here the seam is bound to the CUDA gather-decode of `saeb200` -- there is no Triton and no eager fallback; CPU tensors
raise."""
from __future__ import annotations

import torch
from torch import Tensor

from saeb200 import engine


def cuda_decode(top_indices: Tensor, top_acts: Tensor, W_dec_T: Tensor) -> Tensor:
    """Same contract as the reference's `triton_decode` / `eager_decode`: `W_dec_T` is `W_dec.mT`, a [d, N] view of
    the contiguous [N, d] parameter; returns sum_j acts[..., j] * W_dec[idx[..., j], :] (no bias)."""
    W = W_dec_T.mT
    if not W.is_contiguous():
        W = W.contiguous()
    out_dtype = top_acts.dtype if top_acts.dtype in (torch.float32, torch.float16, torch.bfloat16) else torch.float32
    return engine.decode(top_indices, top_acts, W, None, out_dtype=out_dtype)


decoder_impl = cuda_decode
