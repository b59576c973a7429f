"""Loader helpers every launcher calls (reference sae_auto_interp/utils.py:44-48,106-135); signatures unchanged."""
from __future__ import annotations

import json
import os
from typing import Dict, Optional

import torch

from .sae import Sae


def load_filter(path: str, device: str = "cuda:0") -> Dict[str, torch.Tensor]:
    with open(path) as fh:
        raw = json.load(fh)
    return {name: torch.tensor(ids, device=device) for name, ids in raw.items()}


def load_saes(sae_path: str, filters: Optional[Dict[str, torch.Tensor]] = None, device="cuda:0") -> Dict[str, Sae]:
    local = os.path.exists(sae_path)
    if filters is None:
        return Sae.load_many(sae_path, local=local, device=device)
    out = {}
    for module_name in filters:
        out[module_name] = (Sae.load_from_disk(os.path.join(sae_path, module_name), device=device) if local
                            else Sae.load_from_hub(sae_path, module_name, device=device))
    return out


def load_single_sae(sae_path: str, module_name: str, device="cuda:0") -> Sae:
    if os.path.exists(sae_path):
        return Sae.load_from_disk(os.path.join(sae_path, module_name), device=device)
    return Sae.load_from_hub(sae_path, module_name, device=device)


def load_explanation(explanation_dir: str) -> Dict[str, str]:
    """{feature name: explanation text} from every json file of a directory; each file is a list of one-entry records
    `{"<module>_feature<i>": text, "prompt": ...}` written by the explain launchers (reference utils.py:51-65)."""
    merged: Dict[str, str] = {}
    for name in os.listdir(explanation_dir):
        path = os.path.join(explanation_dir, name)
        if not os.path.isfile(path):
            continue
        with open(path) as fh:
            for record in json.load(fh):
                merged.update({key: text for key, text in record.items() if key != "prompt"})
    return merged


def maybe_load_llava_model(model_name, rank, dtype, hf_token):
    """(model, processor): LLaVA-NeXT with its processor when the name says so, else a plain `AutoModel` and None
    (reference utils.py:68-88).  The model goes to `cuda:{rank}` in `dtype`."""
    from transformers import AutoModel, LlavaNextForConditionalGeneration, LlavaNextProcessor

    place = dict(device_map={"": f"cuda:{rank}"}, torch_dtype=dtype, token=hf_token)
    if "llava" in model_name:
        return (LlavaNextForConditionalGeneration.from_pretrained(model_name, **place),
                LlavaNextProcessor.from_pretrained(model_name))
    return AutoModel.from_pretrained(model_name, **place), None


def load_llava_quantized(model_name, rank):
    """LLaVA-NeXT with float8 quanto weights, fp16 compute (reference utils.py:91-103)."""
    from transformers import LlavaNextForConditionalGeneration, LlavaNextProcessor, QuantoConfig

    model = LlavaNextForConditionalGeneration.from_pretrained(
        model_name, device_map={"": f"cuda:{rank}"}, quantization_config=QuantoConfig(weights="float8"),
        torch_dtype=torch.float16)
    return model, LlavaNextProcessor.from_pretrained(model_name)


def get_anyres_padded_images(image, image_grid_pinpoints):
    """the image resized to the any-resolution grid cell LLaVA-NeXT picks for it (reference utils.py:138-146)"""
    from transformers.image_processing_utils import select_best_resolution

    best = select_best_resolution([image.size[0], image.size[1]], image_grid_pinpoints)
    return image.resize((best[0], best[1]))


def get_anyres_unpadded_size(orig_height: int, orig_width: int, height: int, width: int, image_grid_pinpoints,
                             patch_size: int):
    """(rows, cols) of image tokens left after LLaVA-NeXT removes the padding of its any-resolution grid (an image
    newline token follows every row); integer arithmetic as in the reference (utils.py:149-184)."""
    from transformers.image_processing_utils import select_best_resolution

    best_h, best_w = select_best_resolution([orig_height, orig_width], image_grid_pinpoints)
    rows = (height // patch_size) * (best_h // height)
    cols = (width // patch_size) * (best_w // width)
    if width / height > cols / rows:      # wider than the grid: rows were padded
        rows -= 2 * ((rows - (height * cols) // width) // 2)
    else:                                 # taller than the grid: columns were padded
        cols -= 2 * ((cols - (width * rows) // height) // 2)
    return rows, cols


def get_llava_image_pos(input_ids, image_tok: int):
    """(start, stop) such that embeddings[start:stop] are the image tokens of a single-image prompt
    (reference utils.py:187-198): start = position of the image placeholder, stop = minus the number of text tokens
    after it."""
    at = input_ids.index(image_tok)
    return at, at + 1 - len(input_ids)


def load_tokenizer(model):
    """left-padding tokenizer whose pad token is its eos token (reference utils.py:237-245)"""
    from transformers import AutoTokenizer

    tok = AutoTokenizer.from_pretrained(model, padding_side="left")
    tok._pad_token = tok._eos_token
    return tok
