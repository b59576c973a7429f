"""Steering (reference features/steering.py): generate with one SAE latent clamped.

The hook body of the reference materialises the dense [1, T, num_latents] latents, writes one column, runs torch.topk
and a Triton decode (features/steering.py:105-124).  Here the clamp is an argument of the fused encode+TopK kernel
(the clamped column is overridden in the GEMM epilogue) and the decode writes fp16 directly; like the reference the
layer output is *replaced* by the reconstruction and the clamp only applies at prefill (sequence length != 1)."""
from __future__ import annotations

import os
from typing import List

import torch
import torch.nn as nn

from saeb200 import engine

from ..sae import Sae


def steering_hook_output(sae: Sae, hidden: torch.Tensor, feature: int, clamp_value: float) -> torch.Tensor:
    """hidden [1, T, d] -> SAE reconstruction with `feature` clamped (prefill only), fp16 [1, T, d]."""
    clamp = feature if hidden.shape[1] != 1 else -1
    top_acts, top_indices = sae.encode(hidden, clamp_feature=clamp, clamp_value=float(clamp_value))
    out = engine.decode(top_indices[0], top_acts[0], sae.W_dec.data, sae.b_dec.data, out_dtype=torch.float16)
    return out.unsqueeze(0)


class SteeringController:
    def __init__(self, sae: Sae, module_name: str, feature_idx: List[int], model: nn.Module, processor, prompt: str,
                 image_path: str = None, k: float = 50):
        self.sae, self.feature_idx, self.model, self.k = sae, feature_idx, model, k
        self.module_name = module_name
        self.hooked_module = model.language_model.get_submodule(module_name)
        self.processor = processor
        local_rank = os.environ.get("LOCAL_RANK")
        self.ddp = local_rank is not None
        self.rank = int(local_rank) if local_rank is not None else 0
        self.conversation = [{"role": "user", "content": [{"type": "text", "text": prompt}]}]
        self.image = None
        if image_path is not None:
            from PIL import Image

            self.image = Image.open(image_path)
            self.conversation[0]["content"].append({"type": "image"})
        self.prompt = processor.apply_chat_template(self.conversation, add_generation_prompt=True)
        self.inputs = processor(images=self.image, text=self.prompt, return_tensors="pt").to(model.device)

    def _generate(self) -> str:
        with torch.no_grad():
            output = self.model.generate(**self.inputs, max_new_tokens=512)
        cont = output[:, self.inputs["input_ids"].shape[-1]:]
        return self.processor.batch_decode(cont, skip_special_tokens=True)[0]

    def run(self):
        original = self._generate()
        results = {}
        for idx in self.feature_idx:
            handles = self.clamp_features_max(self.sae, idx, self.hooked_module, k=self.k)
            try:
                clamped = self._generate()
            finally:
                for h in handles:
                    h.remove()
            results[f"{self.module_name}_feature{idx}"] = {"original_resps": original, "clamped_resps": clamped,
                                                           "idx": idx}
        return results

    def clamp_features_max(self, sae: Sae, feature: int, hooked_module: torch.nn.Module, k: float = 10):
        def hook(module, _inputs, outputs):
            parts = list(outputs)
            parts[0] = steering_hook_output(sae, parts[0], feature, k)
            return tuple(parts) if isinstance(outputs, tuple) else parts[0]

        return [hooked_module.register_forward_hook(hook)]
