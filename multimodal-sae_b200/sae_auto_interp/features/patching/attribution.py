"""`Attribution` (reference features/patching/attribution.py): per-latent attribution of a logit difference,
    attribution[module][i] = sum_dim (clean_rec - corrupted_rec_i) * d metric / d corrupted_rec_i     -> [batch, seq]
where corrupted_rec_i is the SAE reconstruction with latent i switched off."""
from __future__ import annotations

import collections
import json
import os
from functools import partial
from typing import Dict, List, Optional, Sequence, Union

import torch
import torch.distributed as dist

from ...sae import Sae
from .utils import get_logit_diff, get_model_backward_cache_with_sae, get_model_forward_cache_with_sae

os.environ.setdefault("TOKENIZERS_PARALLELISM", "false")


def attribution_for_feature(model, inputs: dict, sae_dict: Dict[str, Sae], module_to_name, metric, feature,
                            clean_cache: Optional[Dict[str, torch.Tensor]] = None) -> Dict[str, torch.Tensor]:
    """One iteration of the reference's loop (attribution.py:131-184) for latent `feature`: {module: [batch, seq]} on
    the CPU.  The clean pass does not depend on the latent; pass `clean_cache` to reuse it."""
    if clean_cache is None:
        _, clean_cache = get_model_forward_cache_with_sae(model, inputs, sae_dict, module_to_name)
    logits, corrupted = get_model_forward_cache_with_sae(model, inputs, sae_dict, module_to_name,
                                                         off_features=feature)
    for rec in corrupted.values():
        rec.retain_grad()
    get_model_backward_cache_with_sae(logits=logits, metrics=metric)
    out = {}
    for name in sae_dict:
        delta = (clean_cache[name].detach() - corrupted[name].detach()) * corrupted[name].grad
        out[name] = delta.sum(dim=-1).cpu()
    return out


class Attribution:
    def __init__(self, model, tokenizer, sae_path: str, data_path: str, selected_sae: str = None,
                 image_processor=None) -> None:
        self.model, self.image_processor = model, image_processor
        local = os.path.exists(sae_path)
        if selected_sae is not None:
            sae = (Sae.load_from_disk(os.path.join(sae_path, selected_sae), device=model.device) if local
                   else Sae.load_from_hub(sae_path, hookpoint=selected_sae, device=model.device))
            self.sae_dict = {selected_sae: sae}
        else:
            self.sae_dict = Sae.load_many(sae_path, local=local, device=model.device)
        for sae in self.sae_dict.values():
            sae.eval()
        self.data_path = data_path
        with open(data_path) as fh:
            self.data = json.load(fh)   # [{"prompt", "answer", "baseline", "image"}, ...]
        self.prompt = [item["prompt"] for item in self.data]
        self.answer = [[str(item["answer"]), str(item["baseline"])] for item in self.data]
        self.image, self.image_sizes = [], []
        for item in self.data:
            from PIL import Image

            img = Image.open(item["image"])
            self.image.append(img)
            self.image_sizes.append([img.size[0], img.size[1]])
        self.pixel_values = (self.image_processor(list(self.image), do_pad=True, return_tensors="pt")["pixel_values"]
                             .to(model.device).to(model.dtype))
        self.prompt_ids = tokenizer(self.prompt, return_tensors="pt")["input_ids"].to(model.device)[:, 1:]
        ids = [[tokenizer.convert_tokens_to_ids(a), tokenizer.convert_tokens_to_ids(b)] for a, b in self.answer]
        self.answer_ids = torch.tensor(ids).to(model.device)
        self.attention_mask = self.prompt_ids.ne(tokenizer.pad_token_id)
        self.name_to_module = {name: model.language_model.get_submodule(name) for name in self.sae_dict}
        self.module_to_name = {mod: name for name, mod in self.name_to_module.items()}
        self.metric = partial(get_logit_diff, answer_token_indices=self.answer_ids)

    def _inputs(self) -> dict:
        return {"input_ids": self.prompt_ids, "pixel_values": self.pixel_values, "image_sizes": self.image_sizes,
                "attention_mask": self.attention_mask}

    def get_attribution(self, indices: Union[List[int], torch.Tensor, None] = None) -> Dict[str, List[torch.Tensor]]:
        ddp = os.environ.get("LOCAL_RANK") is not None
        if indices is None:
            # the reference looks up a misspelt config field here (attribution.py:121) and therefore always sweeps
            # d_in * expansion_factor latents; kept, so that both produce the same number of rows
            first = next(iter(self.sae_dict.values()))
            indices = torch.arange(first.d_in * first.cfg.expansion_factor)
        show = (not dist.is_initialized()) or dist.get_rank() == 0
        try:
            from tqdm import tqdm

            bar = tqdm(total=len(indices), desc="Calculating attribution", disable=not show)
        except Exception:
            bar = None
        result: Dict[str, List[torch.Tensor]] = collections.defaultdict(list)
        inputs = self._inputs()
        # the clean pass is the same for every latent: run it once (the reference repeats it per latent)
        with torch.no_grad():
            _, clean = get_model_forward_cache_with_sae(self.model, inputs, self.sae_dict, self.module_to_name)
        for idx in indices:
            per_module = attribution_for_feature(self.model, inputs, self.sae_dict, self.module_to_name, self.metric,
                                                 int(idx), clean_cache=clean)
            for name, att in per_module.items():
                result[name].append(att)
            if bar is not None:
                bar.update(1)
        if bar is not None:
            bar.close()
        if ddp and dist.is_initialized():
            dist.barrier()
        return result
