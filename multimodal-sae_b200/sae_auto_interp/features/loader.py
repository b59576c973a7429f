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
             transform: Optional[Callable] = None):
        """Per buffer, per feature: record -> `constructor(record=, buffer_output=)` -> `sampler(record)` ->
        `transform(record)` (the reference's callback protocol, features/loader.py:201-248).  `collate=True` returns
        one flat list, otherwise a generator of per-buffer lists."""

        def build(out: BufferOutput) -> FeatureRecord:
            record = FeatureRecord(out.feature)
            for step, kwargs in ((constructor, {"record": record, "buffer_output": out}), (sampler, None),
                                 (transform, None)):
                if step is not None:
                    step(**kwargs) if kwargs else step(record)
            return record

        def records_of(buffer: TensorBuffer) -> List[FeatureRecord]:
            buffer._load()
            return [build(buffer[i]["buffer"]) for i in range(len(buffer))]

        if collate:
            return [rec for buffer in self.buffers for rec in records_of(buffer)]
        return (records_of(buffer) for buffer in self.buffers)
