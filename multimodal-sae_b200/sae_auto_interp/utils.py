"""Loader helpers every launcher calls (reference sae_auto_interp/utils.py:44-48,106-135); signatures unchanged."""
from __future__ import annotations

import json
import os
from typing import Dict, Optional

import torch

from .sae import Sae


def load_filter(path: str, device: str = "cuda:0") -> Dict[str, torch.Tensor]:
    with open(path) as fh:
        raw = json.load(fh)
    return {name: torch.tensor(ids, device=device) for name, ids in raw.items()}


def load_saes(sae_path: str, filters: Optional[Dict[str, torch.Tensor]] = None, device="cuda:0") -> Dict[str, Sae]:
    local = os.path.exists(sae_path)
    if filters is None:
        return Sae.load_many(sae_path, local=local, device=device)
    out = {}
    for module_name in filters:
        out[module_name] = (Sae.load_from_disk(os.path.join(sae_path, module_name), device=device) if local
                            else Sae.load_from_hub(sae_path, module_name, device=device))
    return out


def load_single_sae(sae_path: str, module_name: str, device="cuda:0") -> Sae:
    if os.path.exists(sae_path):
        return Sae.load_from_disk(os.path.join(sae_path, module_name), device=device)
    return Sae.load_from_hub(sae_path, module_name, device=device)
