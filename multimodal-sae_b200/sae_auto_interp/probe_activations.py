"""Probe which SAE features fire on one image / text (reference tools/probe_activations.py).

The reference's hook materialises the dense latents `sae.pre_acts(hidden)` [1, T, 131072], averages them over the
tokens, takes the `top_k` features by mean activation and slices their columns out (tools/probe_activations.py:
109-126).  `probe_hidden` returns the same two tensors from the fused engine: the token mean comes from the GEMM's
dense store reduced chunk by chunk (saeb_column_sums), the per-token maps of the selected features are re-evaluated
exactly (saeb_feature_maps); `make_hook` wraps it as the forward hook the tool registers.  The rest of the tool (model
loading, mask upsampling, PNG output) is host glue around these two tensors and unchanged in spirit.
"""
from __future__ import annotations

from typing import Optional, Sequence, Tuple

import torch

from saeb200 import engine


def probe_hidden(sae, hidden: torch.Tensor, interval: Sequence[int], *, drop_first: bool = False
                 ) -> Tuple[torch.Tensor, torch.Tensor]:
    """hidden [1, T, d] -> (topk_indices [interval[1] - interval[0]] int64 on the CPU, topk_acts
    [n_features, T'] fp32 on the CPU), T' = T - 1 when `drop_first` (the tool skips the BOS position for image-only
    llama inputs, tools/probe_activations.py:115-116).  Features are ordered by mean activation, descending."""
    h = hidden[0] if hidden.dim() == 3 else hidden
    if drop_first:
        h = h[1:]
    mean = engine.mean_activations(h, sae.packed_encoder(2))
    _, order = engine.dense_topk(mean[None], int(interval[1]))
    top = order[0, int(interval[0]):]
    maps = engine.feature_maps(h, sae.encoder.weight.data, sae.encoder.bias.data, sae.b_dec.data, top)
    return top.cpu(), maps.cpu()


def make_hook(sae, interval: Sequence[int], sink: dict, *, drop_first: bool = False):
    """forward hook storing `topk_indices` / `topk_acts` in `sink` (the tool keeps them in module globals)"""

    def hook(module, _inputs, outputs):
        hidden = outputs[0] if isinstance(outputs, (tuple, list)) else outputs
        sink["topk_indices"], sink["topk_acts"] = probe_hidden(sae, hidden, interval, drop_first=drop_first)

    return hook


def base_image_maps(topk_acts: torch.Tensor, base_img_tokens: int = 576, patch_size: int = 24):
    """per-feature activation map over the base image tokens (tools/probe_activations.py:143-149)"""
    return [acts[:base_img_tokens].view(patch_size, patch_size) for acts in topk_acts]


def main(argv: Optional[Sequence[str]] = None) -> None:   # pragma: no cover - needs the LLaVA checkpoint
    import argparse
    import json
    import os

    from PIL import Image
    from transformers import AutoTokenizer

    from .features.features import upsample_mask
    from .utils import load_single_sae, maybe_load_llava_model

    ap = argparse.ArgumentParser()
    ap.add_argument("--model", default="llava-hf/llama3-llava-next-8b-hf")
    ap.add_argument("--sae-path", required=True)
    ap.add_argument("--module-name", default="model.layers.24")
    ap.add_argument("--image-path")
    ap.add_argument("--text")
    ap.add_argument("--top-k", type=int, default=20)
    ap.add_argument("--interval")
    ap.add_argument("--save-to", default="probe_out")
    args = ap.parse_args(argv)
    sae = load_single_sae(args.sae_path, args.module_name)
    model, processor = maybe_load_llava_model(args.model, rank=0, dtype=torch.float16, hf_token=None)
    tokenizer = AutoTokenizer.from_pretrained(args.model)
    image = Image.open(args.image_path) if args.image_path is not None else None
    assert image is not None or args.text is not None, "Image and text can no both be None"
    interval = [int(i) for i in args.interval.split("-")] if args.interval else [0, args.top_k]
    if args.text is not None:
        content = [{"type": "text", "text": args.text}] + ([{"type": "image"}] if image is not None else [])
        prompt = processor.apply_chat_template([{"role": "user", "content": content}], add_generation_prompt=True)
    else:
        prompt = "<image>"
    inputs = processor(images=image, text=prompt, return_tensors="pt").to(model.device)
    sink: dict = {}
    hooked = model.language_model.get_submodule(args.module_name)
    handle = hooked.register_forward_hook(make_hook(
        sae, interval, sink, drop_first="llama" in tokenizer.name_or_path and args.text is None))
    try:
        with torch.no_grad():
            model(**{k: v for k, v in inputs.items()})
    finally:
        handle.remove()
    os.makedirs(args.save_to, exist_ok=True)
    if image is not None:
        image_dir = os.path.join(args.save_to, "images")
        os.makedirs(image_dir, exist_ok=True)
        background = Image.new("L", (336, 336), 0).convert("RGB")
        for idx, amap in zip(sink["topk_indices"], base_image_maps(sink["topk_acts"])):
            mask = upsample_mask(amap, (336, 336))
            Image.composite(background, image.resize((336, 336)), mask).convert("RGB").save(
                os.path.join(image_dir, f"feat_{idx}.png"))
    with open(os.path.join(args.save_to, "filters.json"), "w") as fh:
        json.dump({args.module_name: sink["topk_indices"].tolist()}, fh)


if __name__ == "__main__":   # pragma: no cover
    main()
