"""Launcher configuration dataclasses (reference sae_auto_interp/config.py).  Field names / defaults are the CLI and
on-disk contract of the cache -> explain pipeline; `FeatureConfig` and `CacheConfig` drive the hot path.

The classes are built from one table per class (name, type, default); `positional` marks the fields the reference's
command lines take without a flag."""
from __future__ import annotations

from dataclasses import MISSING
from typing import Literal, Union

from ._compat import table_dataclass

_PYTHIA, _REDPAJAMA = "EleutherAI/pythia-160m", "togethercomputer/RedPajama-Data-1T-Sample"


def _config(name: str, doc: str, *rows, positional=()):
    return table_dataclass(name, doc, __name__, *rows, positional=positional)


ExperimentConfig = _config(
    "ExperimentConfig", "explain / score launchers: which model, which SAEs, how many examples per feature",
    ("model", str, _PYTHIA), ("dataset", str, (_REDPAJAMA,)), ("sae_path", Union[str, None], None),
    ("n_examples_train", int, 10), ("n_examples_test", int, 7), ("n_quantiles", int, 10), ("n_random", int, 5),
    ("train_type", Literal["top", "random", "quantile"], "top"),
    ("explainer", str, "meta-llama/Meta-Llama-3.1-405B-Instruct-FP8"), ("explanation_dir", str, "./explanation_dir"),
    ("scores_dir", str, "./scores_dir"), ("selected_layers", list, []), ("split", str, "train"),
    ("save_dir", str, "./features_cache"), ("filters_path", str, None))

FeatureConfig = _config(
    "FeatureConfig", "cache reader / example constructors: SAE width, window length, examples kept, split files",
    ("width", int, MISSING), ("example_ctx_len", int, 64), ("min_examples", int, 200), ("max_examples", int, 10000),
    ("n_splits", int, 2))

CacheConfig = _config(
    "CacheConfig", "cache launchers: host model, dataset, batch / context sizes, where the split files go",
    ("model", str, _PYTHIA), ("dataset", str, _REDPAJAMA), ("sae_path", Union[str, None], None),
    ("batch_size", int, 32), ("load_in_8bit", bool, False), ("split", str, "train"), ("n_splits", int, 2),
    ("ctx_len", int, 2048), ("hf_token", Union[str, None], None), ("save_dir", str, "./features_cache"),
    ("verbosity", str, "INFO"), ("filters_path", str, None), positional=("model", "dataset"))

AttributionConfig = _config(
    "AttributionConfig", "attribution-patching launcher; data_path is a list of {prompt, answer, baseline, image}",
    ("model", str, _PYTHIA), ("data_path", str, "./data/digit.json"), ("sae_path", Union[str, None], None),
    ("selected_sae", str, "layers.24"), ("save_dir", str, "./attribution_cache"), positional=("model",))
