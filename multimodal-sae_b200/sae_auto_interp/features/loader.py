"""Cache reader (reference features/loader.py): split files -> per-feature (locations[:, :2], activations).

Host-side and format-compatible with the reference's `{module}/{start}_{end}.safetensors` files.  A split file is
grouped by feature ONCE (stable sort + searchsorted) instead of one boolean mask over all entries per requested
feature; the order of a feature's rows is unchanged.
"""
from __future__ import annotations

import os
from typing import Callable, Dict, Iterator, List, NamedTuple, Optional

import torch
from safetensors.torch import load_file
from torch.utils.data import Dataset

from ..config import FeatureConfig
from .features import Feature, FeatureRecord


class BufferOutput(NamedTuple):
    feature: Feature
    locations: torch.Tensor  # [n, 2] = (row, pos)
    activations: torch.Tensor  # [n]


class _FeatureGroups:
    """rows of a split file grouped by feature id"""

    def __init__(self, locations: torch.Tensor):
        ids = locations[:, 2].contiguous()
        self.order = torch.sort(ids, stable=True).indices
        self.sorted_ids = ids[self.order]
        self.present = torch.unique_consecutive(self.sorted_ids)

    def rows_of(self, feature: int) -> torch.Tensor:
        a = int(torch.searchsorted(self.sorted_ids, feature))
        b = int(torch.searchsorted(self.sorted_ids, feature, right=True))
        return self.order[a:b]


class TensorBuffer(Dataset):
    """Lazy view of one split file; indexable (`buffer[i]["buffer"]`) and iterable (with `min_examples` filtering,
    yielding `None` for skipped features) like the reference's."""

    def __init__(self, path: str, module_path: str, features: Optional[torch.Tensor] = None, min_examples: int = 120):
        super().__init__()
        self.tensor_path, self.module_path = path, module_path
        self.features, self.min_examples = features, min_examples
        self.start = 0
        self.activations = self.locations = None
        self._groups: Optional[_FeatureGroups] = None

    # -- loading
    def _load(self) -> None:
        blob = load_file(self.tensor_path)
        self.locations, self.activations = blob["locations"], blob["activations"]
        self._groups = _FeatureGroups(self.locations)
        if self.features is None:
            self.features = self._groups.present

    def _ensure(self) -> None:
        if self._groups is None:
            self._load()

    def _emit(self, feature: int) -> BufferOutput:
        rows = self._groups.rows_of(feature)
        return BufferOutput(Feature(self.module_path, feature), self.locations[rows][:, :2], self.activations[rows])

    # -- Dataset protocol
    def __len__(self) -> int:
        if self.features is None:
            self._ensure()
        return len(self.features)

    def __getitem__(self, index):
        self._ensure()
        self.start += 1
        return {"buffer": self._emit(int(self.features[index]))}

    # -- iterator protocol (the only place `min_examples` is applied, as in the reference)
    def __iter__(self) -> Iterator:
        self._ensure()
        self.start, self.end = 0, len(self.features)
        return self

    def __next__(self):
        if self.start >= self.end:
            self.activations = self.locations = self._groups = None
            raise StopIteration
        feature = int(self.features[self.start])
        self.start += 1
        if self._groups.rows_of(feature).numel() < self.min_examples:
            return None
        return self._emit(feature)


def _split_path(raw_dir: str, module: str, lo, hi) -> str:
    """file of the features [lo, hi) -- the name carries the inclusive upper end"""
    return f"{raw_dir}/{module}/{lo}_{hi - 1}.safetensors"


class FeatureDataset:
    """One `TensorBuffer` per (module, split file): every feature, or only the selected ones."""

    def __init__(self, raw_dir: str, cfg: FeatureConfig, modules: Optional[List[str]] = None,
                 features: Optional[Dict[str, torch.Tensor]] = None):
        self.cfg = cfg
        self.buffers: List[TensorBuffer] = []
        (self._build if features is None else self._build_selected)(*(
            (raw_dir, modules) if features is None else (raw_dir, modules, features)))

    def _edges(self) -> torch.Tensor:
        return torch.linspace(0, self.cfg.width, steps=self.cfg.n_splits + 1).long()

    def _build(self, raw_dir: str, modules: Optional[List[str]] = None) -> None:
        edges = self._edges()
        for module in (modules if modules is not None else os.listdir(raw_dir)):
            self.buffers += [TensorBuffer(_split_path(raw_dir, module, lo, hi), module,
                                          min_examples=self.cfg.min_examples)
                             for lo, hi in zip(edges[:-1], edges[1:])]

    def _build_selected(self, raw_dir: str, modules: List[str], features: Dict[str, torch.Tensor]) -> None:
        edges = self._edges()
        for module in modules:
            wanted = features[module]
            which = torch.bucketize(wanted, edges, right=True)  # split number (1-based) of every wanted feature
            for b in torch.unique(which):
                self.buffers.append(TensorBuffer(_split_path(raw_dir, module, edges[b - 1], edges[b]), module,
                                                 wanted[which == b], min_examples=self.cfg.min_examples))

    def __len__(self) -> int:
        return len(self.buffers)

    def load(self, collate: bool = False, constructor: Optional[Callable] = None, sampler: Optional[Callable] = None,
             transform: Optional[Callable] = None, device="auto"):
        """Per buffer, per feature: record -> `constructor(record=, buffer_output=)` -> `sampler(record)` ->
        `transform(record)` (the reference's callback protocol, features/loader.py:201-248).  `collate=True` returns
        one flat list, otherwise a generator of per-buffer lists.

        Device route: when `constructor` is a `functools.partial` of this package's `pool_max_activation_windows` /
        `default_constructor` / `pool_max_activations_windows_image` and a CUDA device is available (`device="auto"`,
        or an explicit device; `None` forces the host route), the RANKING step of the constructor -- the densify +
        pool + topk the reference repeats for every feature -- runs once per split file on the device for all its
        features (saeb200.engine.coo_top_windows) and each constructor call only materialises its selected examples.
        This is what `launch/explain/explain_images.py:56-65` drives."""
        plan = _device_plan(constructor, self.cfg) if device is not None else None
        dev = None
        if plan is not None:
            if device == "auto":
                dev = torch.device("cuda") if torch.cuda.is_available() else None
            else:
                dev = torch.device(device)

        def build(out: BufferOutput, ranked=None) -> FeatureRecord:
            record = FeatureRecord(out.feature)
            if constructor is not None:
                kwargs = {"record": record, "buffer_output": out}
                if ranked is not None:
                    kwargs["ranked"] = ranked
                constructor(**kwargs)
            for step in (sampler, transform):
                if step is not None:
                    step(record)
            return record

        def records_of(buffer: TensorBuffer) -> List[FeatureRecord]:
            buffer._load()
            ranking = _rank_split_on_device(buffer, plan, dev) if dev is not None else None
            out = []
            for i in range(len(buffer)):
                bo = buffer[i]["buffer"]
                out.append(build(bo, None if ranking is None else ranking(bo.feature.feature_index)))
            return out

        if collate:
            return [rec for buffer in self.buffers for rec in records_of(buffer)]
        return (records_of(buffer) for buffer in self.buffers)


def _device_plan(constructor, cfg):
    """what the device has to rank for this constructor, or None if it is not one of the window constructors"""
    import functools

    from . import constructors as C

    if not isinstance(constructor, functools.partial):
        return None
    kw = constructor.keywords
    if constructor.func is C.pool_max_activations_windows_image:
        c = kw.get("cfg", cfg)
        return {"kind": "image", "n_base": getattr(kw.get("processor"), "num_image_tokens", 576),
                "n_top": c.max_examples + 50}
    if constructor.func in (C.pool_max_activation_windows, C.default_constructor):
        c = kw.get("cfg", cfg)
        tokens = kw.get("tokens")
        if tokens is None or not hasattr(tokens, "shape"):
            return None
        ctx_len = kw.get("ctx_len") or (c.example_ctx_len if c is not None else None)
        n_top = kw.get("max_examples") or (c.max_examples if c is not None else None)
        if ctx_len is None or n_top is None:
            return None
        return {"kind": "text", "ctx_len": int(ctx_len), "seq_len": int(tokens.shape[1]), "n_top": int(n_top)}
    return None


def _rank_split_on_device(buffer: "TensorBuffer", plan, dev):
    """one device pass over a split file -> lookup `feature id -> (scores, window ids)` (CPU tensors)"""
    from saeb200 import engine

    if plan["kind"] == "image":
        feats, offs, scores, wins = engine.coo_top_windows(buffer.locations, buffer.activations, plan["n_top"],
                                                           n_base=plan["n_base"], device=dev)
    else:
        feats, offs, scores, wins = engine.coo_top_windows(buffer.locations, buffer.activations, plan["n_top"],
                                                           ctx_len=plan["ctx_len"], seq_len=plan["seq_len"], device=dev)
    feats, offs, scores, wins = feats.cpu(), offs.cpu(), scores.cpu(), wins.cpu()   # one copy back per split file
    index = {int(f): i for i, f in enumerate(feats.tolist())}
    empty = (scores[:0], wins[:0])

    def lookup(feature: int):
        i = index.get(int(feature))
        if i is None:
            return empty
        a, b = int(offs[i]), int(offs[i + 1])
        return scores[a:b], wins[a:b]

    return lookup
