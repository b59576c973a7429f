"""Forward / backward passes of the host model with SAE reconstructions spliced in (reference
features/patching/utils.py:9-80).

The reference's hook materialises the dense latents to switch features off (`pre_acts -> latents * mask -> select_topk
-> decode`, utils.py:42-49).  Switching ONE latent off is the steering clamp with value 0, which the fused
encode + TopK kernel takes as an argument (the column is overridden in the GEMM epilogue and, being 0, can never pass
the running threshold) -- no dense [tokens, num_latents] tensor.  Several latents at once go through the dense route
exactly as in the reference.  The decode runs through the autograd seam (`sae/utils.py::SparseDecode`), so the
reconstruction carries `requires_grad` and `tensor.retain_grad()` works as `Attribution.get_attribution` expects.
"""
from __future__ import annotations

from typing import Any, Dict, Optional, Sequence, Tuple, Union

import torch

from ...sae import Sae


def get_logit_diff(logits: torch.Tensor, answer_token_indices: torch.Tensor) -> torch.Tensor:
    """mean over the batch of logit[correct] - logit[baseline] at the last position (reference utils.py:9-19);
    `answer_token_indices` is [batch, 2] = (correct id, baseline id)."""
    last = logits[:, -1, :] if logits.dim() == 3 else logits
    picked = last.gather(1, answer_token_indices[:, :2])
    return (picked[:, 0] - picked[:, 1]).mean()


def _sae_reconstruction(sae: Sae, hidden: torch.Tensor, off_features) -> torch.Tensor:
    """[tokens, d] -> fp16 reconstruction with `off_features` (None, one id, or several ids) forced to 0"""
    if off_features is None:
        top_acts, top_indices = sae.encode(hidden)
    elif isinstance(off_features, int) or (torch.is_tensor(off_features) and off_features.dim() == 0):
        top_acts, top_indices = sae.encode(hidden, clamp_feature=int(off_features), clamp_value=0.0)
    else:
        latents = sae.pre_acts(hidden)
        keep = torch.ones(latents.shape[-1], dtype=latents.dtype, device=latents.device)
        keep[torch.as_tensor(off_features, device=latents.device)] = 0
        top_acts, top_indices = sae.select_topk(latents * keep)
    return sae.decode(top_acts, top_indices).to(torch.float16)


def get_model_forward_cache_with_sae(model: torch.nn.Module, inputs: Dict[str, Any], sae_dict: Dict[str, Sae],
                                     module_to_name: Dict[torch.nn.Module, str],
                                     off_features: Optional[Union[int, Sequence[int], torch.Tensor]] = None
                                     ) -> Tuple[torch.Tensor, Dict[str, torch.Tensor]]:
    """Run `model(**inputs)` with every hooked module's output replaced by its SAE reconstruction (fp16, same
    [batch, seq, dim] shape); returns (logits, {module name: reconstruction}) (reference utils.py:22-70)."""
    cache: Dict[str, torch.Tensor] = {}

    def splice(module, _inputs, outputs):
        is_tuple = isinstance(outputs, tuple)
        parts = list(outputs)
        name = module_to_name[module]
        hidden = parts[0]
        rec = _sae_reconstruction(sae_dict[name], hidden.flatten(0, 1), off_features).view(hidden.shape)
        cache[name] = rec
        return (rec, *parts[1:]) if is_tuple else rec

    handles = [m.register_forward_hook(splice) for m in module_to_name]
    try:
        logits = model(**inputs)["logits"]
    finally:
        for h in handles:
            h.remove()
    return logits, cache


def get_model_backward_cache_with_sae(logits: torch.Tensor, metrics) -> torch.Tensor:
    """metric(logits).backward(); gradients land on the cached reconstructions that called retain_grad()
    (reference utils.py:73-80)."""
    value = metrics(logits)
    value.backward()
    return value
