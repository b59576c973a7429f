"""Picking which examples of a feature go to the explainer (reference features/samplers.py); host-side list work."""
from __future__ import annotations

import random
from typing import Dict, List, Sequence

from .features import FeatureRecord


def split_quantiles(examples: Sequence, n_quantiles: int, n_samples: int, seed: int = 22) -> List:
    """Up to `n_samples` random examples from each of `n_quantiles` equal slices of the (sorted) example list."""
    rng_state = random.getstate()
    random.seed(seed)
    try:
        width = len(examples) // n_quantiles
        out: List = []
        for q in range(n_quantiles):
            part = list(examples[q * width:(q + 1) * width])
            out += random.sample(part, min(n_samples, len(part)))
        return out
    finally:
        random.setstate(rng_state)


def _top(examples, n, **_):
    return examples[:n]


def _random(examples, n, seed=22, **_):
    random.seed(seed)
    return random.sample(examples, n)


def _quantile(examples, n, n_quantiles=10, **_):
    return split_quantiles(examples, n_quantiles, n)


_STRATEGIES = {"top": _top, "random": _random, "quantile": _quantile}


def train(examples: List, n_train: int, train_type: str, seed: int = 22, n_quantiles: int = 10) -> List:
    try:
        pick = _STRATEGIES[train_type]
    except KeyError:
        raise ValueError(f"Invalid train_type: {train_type}") from None
    return pick(examples, n_train, seed=seed, n_quantiles=n_quantiles)


def sample(record: FeatureRecord, cfg) -> None:
    """Fill `record.train` according to cfg.train_type / n_examples_train / n_quantiles."""
    record.train = train(record.examples, cfg.n_examples_train, cfg.train_type, n_quantiles=cfg.n_quantiles)


def sample_with_explanation(record: FeatureRecord, cfg, explanations: Dict[str, str]) -> None:
    sample(record, cfg)
    record.explanation = explanations[str(record.feature)]
