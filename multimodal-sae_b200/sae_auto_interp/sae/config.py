"""SAE configuration.  Field names and defaults are the checkpoint contract: `cfg.json` written by
`Sae.save_to_disk` holds exactly these keys plus `d_in` (reference sae/config.py:7-29, sae/sae.py:135-162)."""
from __future__ import annotations

from dataclasses import dataclass
from typing import Union

from .._compat import Serializable, list_field


@dataclass
class SaeConfig(Serializable):
    expansion_factor: int = 32  # num_latents = d_in * expansion_factor when num_latents == 0
    normalize_decoder: bool = True  # unit-norm decoder rows at construction
    num_latents: int = 0
    k: int = 32  # non-zeros kept per token
    multi_topk: bool = False  # training-only loss; not evaluated by the inference engine
    signed: bool = False  # accepted so that old checkpoints load


@dataclass
class TrainConfig(Serializable):
    """Kept for import compatibility (`from sae_auto_interp.sae import TrainConfig`); training is out of scope."""

    sae: SaeConfig
    batch_size: int = 8
    grad_acc_steps: int = 1
    micro_acc_steps: int = 1
    lr: Union[float, None] = None
    lr_warmup_steps: int = 1000
    auxk_alpha: float = 0.0
    dead_feature_threshold: int = 10_000_000
    hookpoints: list = list_field()
    layers: list = list_field()
    layer_stride: int = 1
    distribute_modules: bool = False
    save_every: int = 1000
    log_to_wandb: bool = True
    run_name: Union[str, None] = None
    wandb_log_frequency: int = 1

    def __post_init__(self):
        if self.layers and self.layer_stride != 1:
            raise ValueError("Cannot specify both `layers` and `layer_stride`.")
