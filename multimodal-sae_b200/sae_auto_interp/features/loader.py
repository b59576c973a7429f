"""Cache reader (reference features/loader.py): split files -> per-feature (locations[:, :2], activations).

Host-side and format-compatible.  A split file is grouped by feature once (stable sort + searchsorted) instead of a
full boolean mask per requested feature; the per-feature row order is unchanged."""
from __future__ import annotations

import os
from typing import Callable, Dict, List, NamedTuple, Optional

import torch
from safetensors.torch import load_file
from torch.utils.data import Dataset

from ..config import FeatureConfig
from .features import Feature, FeatureRecord


class BufferOutput(NamedTuple):
    feature: Feature
    locations: torch.Tensor  # [n, 2] (row, pos)
    activations: torch.Tensor  # [n]


class TensorBuffer(Dataset):
    """Lazy view of one split file."""

    def __init__(self, path: str, module_path: str, features: Optional[torch.Tensor] = None, min_examples: int = 120):
        super().__init__()
        self.tensor_path = path
        self.module_path = module_path
        self.features = features
        self.min_examples = min_examples
        self.start = 0
        self.activations = None
        self.locations = None
        self._order = self._keys = None

    def _load(self):
        data = load_file(self.tensor_path)
        self.activations, self.locations = data["activations"], data["locations"]
        feat = self.locations[:, 2].contiguous()
        self._order = torch.sort(feat, stable=True).indices
        self._keys = feat[self._order]
        if self.features is None:
            self.features = torch.unique(feat)

    def __len__(self):
        return len(self.features) if self.features is not None else len(torch.unique(self.locations[:, 2]))

    def _rows(self, feature: int) -> torch.Tensor:
        lo = int(torch.searchsorted(self._keys, feature))
        hi = int(torch.searchsorted(self._keys, feature, right=True))
        return self._order[lo:hi]

    def _output(self, feature: int) -> BufferOutput:
        rows = self._rows(feature)
        return BufferOutput(Feature(self.module_path, feature), self.locations[rows][:, :2], self.activations[rows])

    def __getitem__(self, index):
        if self.locations is None:
            self._load()
        self.start += 1
        return {"buffer": self._output(int(self.features[index]))}

    def __iter__(self):
        if self.locations is None:
            self._load()
        self.start = 0
        self.end = len(self.features)
        return self

    def __next__(self):
        if self.start >= self.end:
            self.activations = self.locations = self._order = self._keys = None
            raise StopIteration
        feature = int(self.features[self.start])
        self.start += 1
        if self._rows(feature).numel() < self.min_examples:  # min_examples only applies on the iterator path
            return None
        return self._output(feature)


class FeatureDataset:
    """One `TensorBuffer` per (module, split file) -- all features, or only the selected ones."""

    def __init__(self, raw_dir: str, cfg: FeatureConfig, modules: Optional[List[str]] = None,
                 features: Optional[Dict[str, torch.Tensor]] = None):
        self.cfg = cfg
        self.buffers: List[TensorBuffer] = []
        if features is None:
            self._build(raw_dir, modules)
        else:
            self._build_selected(raw_dir, modules, features)

    def _edges(self):
        return torch.linspace(0, self.cfg.width, steps=self.cfg.n_splits + 1).long()

    def _build(self, raw_dir: str, modules: Optional[List[str]] = None):
        edges = self._edges()
        for module in (os.listdir(raw_dir) if modules is None else modules):
            for start, end in zip(edges[:-1], edges[1:]):
                self.buffers.append(TensorBuffer(f"{raw_dir}/{module}/{start}_{end - 1}.safetensors", module,
                                                 min_examples=self.cfg.min_examples))

    def _build_selected(self, raw_dir: str, modules: List[str], features: Dict[str, torch.Tensor]):
        edges = self._edges()
        for module in modules:
            wanted = features[module]
            bucket = torch.bucketize(wanted, edges, right=True)
            for b in torch.unique(bucket):
                start, end = edges[b - 1], edges[b]
                self.buffers.append(TensorBuffer(f"{raw_dir}/{module}/{start}_{end - 1}.safetensors", module,
                                                 wanted[bucket == b], min_examples=self.cfg.min_examples))

    def __len__(self):
        return len(self.buffers)

    def load(self, collate: bool = False, constructor: Optional[Callable] = None, sampler: Optional[Callable] = None,
             transform: Optional[Callable] = None):
        """For every buffer and feature: build the record, let `constructor` pick the top examples, `sampler` the
        train split and `transform` post-process (same callback protocol as the reference, :201-248)."""

        def process(out: BufferOutput) -> FeatureRecord:
            record = FeatureRecord(out.feature)
            if constructor is not None:
                constructor(record=record, buffer_output=out)
            if sampler is not None:
                sampler(record)
            if transform is not None:
                transform(record)
            return record

        def worker(buffer: TensorBuffer):
            buffer._load()
            return [process(buffer[i]["buffer"]) for i in range(len(buffer))]

        if collate:
            records = []
            for buffer in self.buffers:
                records.extend(worker(buffer))
            return records
        return (worker(buffer) for buffer in self.buffers)
