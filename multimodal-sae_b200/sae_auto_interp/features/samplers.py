"""Example samplers (reference features/samplers.py): pure host-side list handling."""
from __future__ import annotations

import random
from typing import Dict, List

from .features import Example, FeatureRecord


def split_quantiles(examples: List[Example], n_quantiles: int, n_samples: int, seed: int = 22):
    random.seed(seed)
    size = len(examples) // n_quantiles
    picked = []
    for q in range(n_quantiles):
        bucket = examples[q * size:(q + 1) * size]
        picked.extend(random.sample(bucket, min(len(bucket), n_samples)))
    return picked


def train(examples: List[Example], n_train: int, train_type: str, seed: int = 22, n_quantiles: int = 10):
    if train_type == "top":
        return examples[:n_train]
    if train_type == "random":
        random.seed(seed)
        return random.sample(examples, n_train)
    if train_type == "quantile":
        return split_quantiles(examples, n_quantiles, n_train)
    raise ValueError(f"Invalid train_type: {train_type}")


def sample(record: FeatureRecord, cfg) -> None:
    record.train = train(record.examples, n_train=cfg.n_examples_train, train_type=cfg.train_type,
                         n_quantiles=cfg.n_quantiles)


def sample_with_explanation(record: FeatureRecord, cfg, explanations: Dict[str, str]) -> None:
    sample(record, cfg)
    record.explanation = explanations[f"{record.feature}"]
