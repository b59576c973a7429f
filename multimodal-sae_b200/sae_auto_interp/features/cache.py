"""Feature-activation cache writer (reference features/cache.py) on the fused B200 path.

Per batch the reference runs  pre_acts -> torch.topk -> zeros_like + scatter_ -> nonzero / mask-gather / isin
(features/cache.py:206-218, :73-92), i.e. three dense [batch, seq, num_latents] fp32 tensors.  Here a batch is one
call to the fused encode+TopK kernel and one COO-extraction kernel; the emitted `locations` (int64 [nnz, 3] =
(row, pos, feature), row-major order) and `activations` (fp32 [nnz]) are the same tensors the reference caches, and
the split-file format consumed by `FeatureDataset` is unchanged.
"""
from __future__ import annotations

import os
import re
from collections import defaultdict
from typing import Dict, Optional

import torch
import torch.distributed as dist
from safetensors.torch import load_file, save_file
from torch.utils.data import DataLoader

from saeb200 import engine

from ..sae import Sae

try:  # progress bars are cosmetic
    from tqdm import tqdm
except Exception:  # pragma: no cover
    def tqdm(it=None, **_kw):
        return it


class Cache:
    """Accumulates per-module COO triples (same fields as the reference's `Cache`).

    On the fused path (`add_topk` with CUDA tensors) the triples of a whole run accumulate in a device arena
    (saeb200.engine.CooArena, saeb_coo_append): no `.cpu()` per batch like the reference (features/cache.py:52-53);
    `save()` does the single device-to-host copy and leaves CPU tensors in `feature_locations` /
    `feature_activations`, as the reference does.  `device_tensors(module)` hands the arena to `save_splits`, which
    then sorts by split on the device."""

    def __init__(self, shard_size: int, filters: Optional[Dict[str, torch.Tensor]] = None, batch_size: int = 64):
        self.feature_locations = defaultdict(list)
        self.feature_activations = defaultdict(list)
        self._arenas: Dict[str, "engine.CooArena"] = {}
        self.filters = filters
        self.batch_size = batch_size
        self.shard_size = shard_size  # global row offset of this rank's dataset shard
        self.rank = dist.get_rank() if dist.is_initialized() else 0
        self._bitmaps: Dict[str, torch.Tensor] = {}

    def _bitmap(self, module_path: str, num_latents: int, device) -> Optional[torch.Tensor]:
        if self.filters is None:
            return None
        key = (module_path, str(device))
        if key not in self._bitmaps:
            self._bitmaps[key] = engine.make_filter_bitmap(self.filters[module_path].to(device), num_latents)
        return self._bitmaps[key]

    def add_topk(self, top_acts: torch.Tensor, top_indices: torch.Tensor, batch_number: int, module_path: str,
                 num_latents: int) -> None:
        """Fused path: TopK output [batch, seq, k] of one batch -> cached triples."""
        seq_len = top_acts.shape[-2]
        row_offset = batch_number * self.batch_size + self.shard_size
        bitmap = self._bitmap(module_path, num_latents, top_acts.device)
        if top_acts.is_cuda:
            arena = self._arenas.get(module_path)
            if arena is None:
                arena = self._arenas[module_path] = engine.CooArena(top_acts.device)
                self.feature_locations[module_path]   # creates the keys the launchers iterate over
                self.feature_activations[module_path]
            arena.append(top_acts, top_indices, seq_len, row_offset=row_offset, filter_bitmap=bitmap)
            return
        loc, act = engine.coo_extract(top_acts, top_indices, seq_len, row_offset=row_offset, filter_bitmap=bitmap)
        self.feature_locations[module_path].append(loc.cpu())
        self.feature_activations[module_path].append(act.cpu())

    def device_tensors(self, module_path: str):
        """(locations, activations) still on the device, or None when this module was cached through the host path"""
        arena = self._arenas.get(module_path)
        return None if arena is None else arena.tensors()

    def get_nonzeros(self, latents: torch.Tensor, module_path: str):
        """Legacy entry for callers holding a dense (already TopK-masked) [batch, seq, feature] tensor: dense ->
        per-token top entries -> the same extraction kernel."""
        k = int((latents.abs() > engine.ACT_THRESHOLD).sum(-1).max().clamp_min(1).item())
        vals, idx = engine.dense_topk(latents, min(k, latents.shape[-1]))
        return engine.coo_extract(vals, idx, latents.shape[-2],
                                  filter_bitmap=self._bitmap(module_path, latents.shape[-1], latents.device))

    def add(self, latents: torch.Tensor, batch_number: int, module_path: str) -> None:
        loc, act = self.get_nonzeros(latents, module_path)
        loc = loc.cpu()
        loc[:, 0] += batch_number * self.batch_size + self.shard_size
        self.feature_locations[module_path].append(loc)
        self.feature_activations[module_path].append(act.cpu())

    def save(self) -> None:
        for module_path, arena in self._arenas.items():
            loc, act = arena.tensors()
            parts_l = self.feature_locations[module_path] if isinstance(self.feature_locations[module_path], list) else []
            parts_a = self.feature_activations[module_path] if isinstance(self.feature_activations[module_path], list) else []
            self.feature_locations[module_path] = parts_l + [loc.cpu()]     # the single device-to-host copy
            self.feature_activations[module_path] = parts_a + [act.cpu()]
        for module_path in list(self.feature_locations.keys()):
            if isinstance(self.feature_locations[module_path], list):
                self.feature_locations[module_path] = torch.cat(self.feature_locations[module_path], dim=0)
                self.feature_activations[module_path] = torch.cat(self.feature_activations[module_path], dim=0)


class FeatureCache:
    def __init__(self, model, tokenizer, submodule_dict: Dict[str, Sae], batch_size: int, shard_size: int,
                 filters: Optional[Dict[str, torch.Tensor]] = None):
        inner = getattr(model, "language_model", None)
        if inner is not None and hasattr(model, "generate") and type(model).__name__.startswith("LlavaNext"):
            self.llava_model, self.model = model, inner
        else:
            self.llava_model, self.model = None, model
        self.tokenizer = tokenizer
        self.name_to_module = {name: self.model.get_submodule(name) for name in submodule_dict}
        self.module_to_name = {mod: name for name, mod in self.name_to_module.items()}
        self.submodule_dict = submodule_dict
        self.batch_size = batch_size
        first = next(iter(submodule_dict.values()))
        self.width = first.cfg.num_latents if first.cfg.num_latents else first.d_in * first.cfg.expansion_factor
        self.cache = Cache(shard_size, filters, batch_size=batch_size)
        # the reference silently drops the last feature id of every split (features/cache.py:247,294); keep that
        # for byte parity of the split files unless asked otherwise
        self.fix_split_bounds = False
        if filters is not None:
            self.filter_submodules(filters)

    def load_token_batches(self, n_tokens: int, tokens: torch.Tensor):
        tokens = tokens[: n_tokens // tokens.shape[1]]
        n = len(tokens) // self.batch_size
        return [tokens[self.batch_size * i: self.batch_size * (i + 1), :] for i in range(n)]

    def filter_submodules(self, filters: Dict[str, torch.Tensor]) -> None:
        self.submodule_dict = {name: sae for name, sae in self.submodule_dict.items() if name in filters}

    # ---- one batch: hidden states of every hooked module -> cached triples
    def _capture(self, run_model):
        captured: Dict[str, torch.Tensor] = {}

        def hook(module, _inputs, outputs):
            captured[self.module_to_name[module]] = outputs[0] if isinstance(outputs, tuple) else outputs

        handles = [mod.register_forward_hook(hook) for mod in self.name_to_module.values()]
        try:
            with torch.no_grad():
                run_model()
        finally:
            for h in handles:
                h.remove()
        return captured

    def _encode_and_cache(self, captured: Dict[str, torch.Tensor], batch_number: int, drop_first: bool) -> None:
        for module_path, hidden in captured.items():
            if module_path not in self.submodule_dict:
                continue
            sae = self.submodule_dict[module_path]
            if drop_first:  # the image path drops the BOS position (reference features/cache.py:407-409)
                hidden = hidden[:, 1:, :]
            top_acts, top_indices = sae.encode(hidden, exact_values=True)   # the cache ranks by these values
            self.cache.add_topk(top_acts, top_indices, batch_number, module_path, sae.num_latents)

    def run(self, n_tokens: int, tokens):
        loader = DataLoader(tokens, batch_size=self.batch_size, drop_last=True, shuffle=False)
        rank_zero = not dist.is_initialized() or dist.get_rank() == 0
        total_tokens = 0
        device = self.model.device
        for batch_number, batch in enumerate(tqdm(loader, desc="Caching features", disable=not rank_zero)):
            total_tokens += n_tokens
            ids = batch["input_ids"].to(device)
            runner = (lambda: self.llava_model(ids)) if self.llava_model is not None else (lambda: self.model(ids))
            captured = self._capture(runner)
            self._encode_and_cache(captured, batch_number, drop_first=False)
        print(f"Total tokens processed: {total_tokens:,}")
        self.cache.save()
        if dist.is_initialized():
            dist.barrier()

    # ---- persistence (formats unchanged)
    def save(self, save_dir):
        for module_path in self.cache.feature_locations.keys():
            save_file({"locations": self.cache.feature_locations[module_path],
                       "activations": self.cache.feature_activations[module_path]},
                      f"{save_dir}/{module_path}.safetensors")

    def _generate_split_indices(self, n_splits):
        bounds = torch.linspace(0, self.width, steps=n_splits + 1).long()
        return list(zip(bounds[:-1], bounds[1:] - 1))

    def save_splits(self, n_splits: int, save_dir, rank: int):
        """One file per feature range: `{module}/Rank{rank}_{start}_{end}.safetensors` (reference :282-309).
        Entries are routed with a single stable sort by split id instead of n_splits boolean passes."""
        split_indices = self._generate_split_indices(n_splits)
        starts = torch.tensor([int(s) for s, _ in split_indices])
        ends = torch.tensor([int(e) for _, e in split_indices])
        for module_path in self.cache.feature_locations.keys():
            dev_pair = self.cache.device_tensors(module_path)
            if dev_pair is not None:   # the arena is still in HBM: bucket + stable sort by split there, one copy back
                loc, act = dev_pair
                starts, ends = starts.to(loc.device), ends.to(loc.device)
            else:
                loc = self.cache.feature_locations[module_path]
                act = self.cache.feature_activations[module_path]
                starts, ends = starts.cpu(), ends.cpu()
            feat = loc[:, 2].contiguous()
            sid = torch.bucketize(feat, starts, right=True) - 1
            upper = ends[sid] + (1 if self.fix_split_bounds else 0)
            keep = feat < upper
            order = torch.sort(sid[keep], stable=True).indices
            loc_k, act_k, sid_k = loc[keep][order].cpu(), act[keep][order].cpu(), sid[keep][order].cpu()
            cuts = torch.searchsorted(sid_k, torch.arange(n_splits + 1))
            module_dir = f"{save_dir}/{module_path}"
            os.makedirs(module_dir, exist_ok=True)
            for i, (start, end) in enumerate(split_indices):
                a, b = int(cuts[i]), int(cuts[i + 1])
                save_file({"locations": loc_k[a:b].contiguous(), "activations": act_k[a:b].contiguous()},
                          f"{module_dir}/Rank{rank}_{start}_{end}.safetensors")

    def concate_safetensors(self, n_splits: int, save_dir):
        """Rank 0: merge every rank's split file into `{module}/{start}_{end}.safetensors` (reference :249-280)."""
        for module_path in self.cache.feature_locations.keys():
            module_dir = f"{save_dir}/{module_path}"
            for start, end in self._generate_split_indices(n_splits):
                pat = re.compile(r"Rank[0-9]+_{}_{}\.safetensors".format(start, end))
                parts = sorted(f for f in os.listdir(module_dir) if pat.search(f))
                acts, locs = [], []
                for name in parts:
                    data = load_file(os.path.join(module_dir, name))
                    acts.append(data["activations"])
                    locs.append(data["locations"])
                    os.remove(os.path.join(module_dir, name))
                save_file({"locations": torch.cat(locs, dim=0), "activations": torch.cat(acts, dim=0)},
                          f"{module_dir}/{start}_{end}.safetensors")


class FeatureImageCache(FeatureCache):
    def __init__(self, model, tokenizer, submodule_dict: Dict[str, Sae], batch_size: int, shard_size: int,
                 filters: Optional[Dict[str, torch.Tensor]] = None, processor=None):
        super().__init__(model, tokenizer, submodule_dict, batch_size, shard_size, filters)
        if processor is None:  # the reference resolves this at import time; do it lazily and only if needed
            from transformers import LlavaNextProcessor

            processor = LlavaNextProcessor.from_pretrained("llava-hf/llama3-llava-next-8b-hf")
        self.processor = processor
        self.prompt = "<image>"

    def run(self, n_tokens: int, tokens):
        def collate(instances):
            images = [inst["image"].convert("RGB") for inst in instances]
            return dict(images=images, image_sizes=torch.tensor([im.size for im in images]).to(torch.long))

        loader = DataLoader(tokens, batch_size=self.batch_size, drop_last=True, shuffle=False, collate_fn=collate,
                            num_workers=0)
        rank_zero = not dist.is_initialized() or dist.get_rank() == 0
        total_images = 0
        device = self.model.device
        for batch_number, batch in enumerate(tqdm(loader, desc="Caching features", disable=not rank_zero)):
            inputs = self.processor(text=[self.prompt] * self.batch_size, images=batch["images"], return_tensors="pt")
            total_images += self.batch_size
            captured = self._capture(lambda: self.llava_model(
                input_ids=inputs["input_ids"].to(device), pixel_values=inputs["pixel_values"].to(device),
                image_sizes=inputs["image_sizes"].to(device), attention_mask=inputs["attention_mask"].to(device)))
            self._encode_and_cache(captured, batch_number, drop_first=True)
        print(f"Total Images processed: {total_images:,}")
        self.cache.save()
        if dist.is_initialized():
            dist.barrier()
