"""Deterministic synthetic SAE weights / activations generated on the device (bench + large-shape tests).

Semantics follow the reference constructor (sae/sae.py:59-66): encoder.weight ~ U(-1/sqrt(d), 1/sqrt(d)) (nn.Linear
default init), W_dec = row-normalised clone; the biases are small Gaussians so the bias path is exercised (trained
checkpoints have non-zero biases).  Activations are N(0,1) rounded to bf16, like the bf16 residual stream the cache
launcher feeds the SAE (launch/cache/cache_image.py:36-39)."""
from __future__ import annotations

import torch


def make_sae(d_in: int, num_latents: int, k: int, device, seed: int = 1234):
    from sae_auto_interp.sae import Sae, SaeConfig

    g = torch.Generator(device=device).manual_seed(seed)
    sae = Sae.__new__(Sae)
    torch.nn.Module.__init__(sae)
    sae.cfg = SaeConfig(num_latents=num_latents, k=k)
    sae.d_in, sae.num_latents = d_in, num_latents
    sae.encoder = torch.nn.Linear(d_in, num_latents, device="meta")
    bound = 1.0 / d_in ** 0.5
    W = (torch.rand(num_latents, d_in, device=device, generator=g) * 2 - 1) * bound
    sae.encoder.weight = torch.nn.Parameter(W)
    sae.encoder.bias = torch.nn.Parameter(torch.randn(num_latents, device=device, generator=g) * 0.01)
    Wd = W.clone()
    Wd /= torch.norm(Wd, dim=1, keepdim=True) + torch.finfo(torch.float32).eps
    sae.W_dec = torch.nn.Parameter(Wd)
    sae.b_dec = torch.nn.Parameter(torch.randn(d_in, device=device, generator=g) * 0.1)
    sae.encoder_planes = 3
    sae.refine_values = "boundary"
    sae._packed = {}
    sae._overlap = None
    sae.overlap_chunk = 9472
    sae.requires_grad_(False)
    return sae


def make_activations(T: int, d_in: int, device, seed: int = 1, dtype=torch.bfloat16, pinned_host: bool = False):
    g = torch.Generator(device=device).manual_seed(seed)
    x = torch.empty((T, d_in), dtype=dtype, device=device)
    step = 1 << 16
    for t0 in range(0, T, step):
        t1 = min(T, t0 + step)
        x[t0:t1] = torch.randn((t1 - t0, d_in), device=device, generator=g).to(dtype)
    if pinned_host:
        h = torch.empty((T, d_in), dtype=dtype, pin_memory=True)
        h.copy_(x)
        return h
    return x
