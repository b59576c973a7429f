"""Launcher configuration dataclasses (reference sae_auto_interp/config.py).  Field names / defaults are the CLI and
on-disk contract of the cache -> explain pipeline; `FeatureConfig` and `CacheConfig` drive the hot path."""
from __future__ import annotations

from dataclasses import dataclass
from typing import Literal, Union

from ._compat import Serializable, field, list_field


@dataclass
class ExperimentConfig(Serializable):
    model: str = "EleutherAI/pythia-160m"
    dataset: str = ("togethercomputer/RedPajama-Data-1T-Sample",)
    sae_path: Union[str, None] = None
    n_examples_train: int = 10
    n_examples_test: int = 7
    n_quantiles: int = 10
    n_random: int = 5
    train_type: Literal["top", "random", "quantile"] = "top"
    explainer: str = "meta-llama/Meta-Llama-3.1-405B-Instruct-FP8"
    explanation_dir: str = "./explanation_dir"
    scores_dir: str = "./scores_dir"
    selected_layers: list = list_field()
    split: str = "train"
    save_dir: str = "./features_cache"
    filters_path: str = None


@dataclass
class FeatureConfig(Serializable):
    width: int  # number of SAE latents
    example_ctx_len: int = 64  # tokens per example window
    min_examples: int = 200
    max_examples: int = 10000
    n_splits: int = 2  # feature-range split files per module


@dataclass
class CacheConfig(Serializable):
    model: str = field(default="EleutherAI/pythia-160m", positional=True)
    dataset: str = field(default="togethercomputer/RedPajama-Data-1T-Sample", positional=True)
    sae_path: Union[str, None] = None
    batch_size: int = 32
    load_in_8bit: bool = False
    split: str = "train"
    n_splits: int = 2
    ctx_len: int = 2048
    hf_token: Union[str, None] = None
    save_dir: str = "./features_cache"
    verbosity: str = "INFO"
    filters_path: str = None


@dataclass
class AttributionConfig(Serializable):
    model: str = field(default="EleutherAI/pythia-160m", positional=True)
    data_path: str = "./data/digit.json"  # list of {"prompt", "answer", "baseline", "image"}
    sae_path: Union[str, None] = None
    selected_sae: str = "layers.24"
    save_dir: str = "./attribution_cache"
