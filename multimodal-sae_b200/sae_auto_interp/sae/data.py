"""Text dataset helpers the cache launchers import (reference sae/data.py): GPT-style chunking and a memory-mapped
token dataset.  Host-side, no device math."""
from __future__ import annotations

import math
from multiprocessing import cpu_count
from typing import List, Union

import numpy as np
import torch
from torch.utils.data import Dataset as TorchDataset


def get_columns_all_equal(dataset) -> List[str]:
    """column names of a `Dataset`, or of a `DatasetDict` whose splits must all agree (reference data.py:103-121)"""
    names = dataset.column_names
    if isinstance(names, dict):
        per_split = list(names.values())
        if any(cols != per_split[0] for cols in per_split):
            raise ValueError("All splits must have the same columns")
        return per_split[0]
    return names


def chunk_and_tokenize(data, tokenizer, *, format: str = "torch", num_proc: int = cpu_count() // 2,
                       text_key: str = "text", max_seq_len: int = 2048, return_final_batch: bool = False,
                       load_from_cache_file: bool = True):
    """Tokenise a text dataset into rows of exactly `max_seq_len` tokens (reference data.py:16-100): documents are
    joined with the eos token (the stream starts with one), long documents spill into the following rows, and the last,
    incomplete row of every map batch is dropped unless `return_final_batch`."""
    chunk = min(tokenizer.model_max_length, max_seq_len)
    sep = tokenizer.eos_token or "<|endoftext|>"

    def tokenize(batch):
        out = tokenizer(sep.join([""] + batch[text_key]), max_length=chunk, return_attention_mask=False,
                        return_overflowing_tokens=True, truncation=True)
        spill = out.pop("overflowing_tokens", None)
        if spill:   # slow tokenizers return one flat list: cut it into rows ourselves
            assert isinstance(out.input_ids[0], int)
            rows = [out["input_ids"]] + [spill[i * chunk:(i + 1) * chunk] for i in range(math.ceil(len(spill) / chunk))]
            out = {"input_ids": rows}
        if not return_final_batch:
            out = {key: rows[:-1] for key, rows in out.items()}
        if len(out["input_ids"]) == 0:
            raise ValueError("Not enough data to create a single complete batch. Either allow the final batch to be "
                             "returned, or supply more data.")
        return out

    data = data.map(tokenize, batched=True, batch_size=2048, num_proc=num_proc,
                    remove_columns=get_columns_all_equal(data), load_from_cache_file=load_from_cache_file)
    return data.with_format(format, columns=["input_ids"])


class MemmapDataset(TorchDataset):
    """rows of `ctx_len` token ids in a flat binary file, memory mapped (reference data.py:124-157)"""

    def __init__(self, data_path: str, ctx_len: int, max_examples: Union[int, None] = None, dtype=np.uint16):
        self.mmap = np.memmap(data_path, dtype=dtype, mode="r").reshape(-1, ctx_len)[:max_examples]

    def __len__(self) -> int:
        return len(self.mmap)

    def __getitem__(self, idx):
        return {"input_ids": torch.from_numpy(self.mmap[idx].astype(np.int64))}

    def _view(self, rows) -> "MemmapDataset":
        other = MemmapDataset.__new__(MemmapDataset)
        other.mmap = rows
        return other

    def select(self, rng: range) -> "MemmapDataset":
        return self._view(self.mmap[rng.start:rng.stop])

    def shard(self, num_shards: int, shard_id: int) -> "MemmapDataset":
        return self._view(np.array_split(self.mmap, num_shards)[shard_id])
