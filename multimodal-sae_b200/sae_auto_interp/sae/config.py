"""SAE configuration.  Field names and defaults are the checkpoint contract: `cfg.json` written by
`Sae.save_to_disk` holds exactly the `SaeConfig` keys plus `d_in` (reference sae/config.py:7-29, sae/sae.py:135-162)."""
from __future__ import annotations

from dataclasses import MISSING
from typing import Union

from .._compat import table_dataclass

SaeConfig = table_dataclass(
    "SaeConfig", "TopK SAE hyper-parameters: num_latents = d_in * expansion_factor when num_latents == 0; k non-zeros "
    "per token; unit-norm decoder rows at construction; multi_topk is a training-only loss (not evaluated by the "
    "inference engine); `signed` is accepted so that old checkpoints load", __name__,
    ("expansion_factor", int, 32), ("normalize_decoder", bool, True), ("num_latents", int, 0), ("k", int, 32),
    ("multi_topk", bool, False), ("signed", bool, False))


def _check_layers(self):
    if self.layers and self.layer_stride != 1:
        raise ValueError("Cannot specify both `layers` and `layer_stride`.")


TrainConfig = table_dataclass(
    "TrainConfig", "Kept for import compatibility (`from sae_auto_interp.sae import TrainConfig`); training is out of "
    "scope of this engine", __name__,
    ("sae", SaeConfig, MISSING), ("batch_size", int, 8), ("grad_acc_steps", int, 1), ("micro_acc_steps", int, 1),
    ("lr", Union[float, None], None), ("lr_warmup_steps", int, 1000), ("auxk_alpha", float, 0.0),
    ("dead_feature_threshold", int, 10_000_000), ("hookpoints", list, []), ("layers", list, []),
    ("layer_stride", int, 1), ("distribute_modules", bool, False), ("save_every", int, 1000),
    ("log_to_wandb", bool, True), ("run_name", Union[str, None], None), ("wandb_log_frequency", int, 1),
    post_init=_check_layers)
