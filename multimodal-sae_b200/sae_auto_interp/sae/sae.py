"""`Sae`: the reference's TopK sparse autoencoder module (sae/sae.py:44-271) backed by the B200 engine.

Attribute names (`encoder`, `W_dec`, `b_dec`, `cfg`, `d_in`, `num_latents`) and the checkpoint layout
(`cfg.json` + `sae.safetensors` with keys encoder.weight / encoder.bias / W_dec / b_dec) are unchanged, so existing
checkpoints and callers work as they are.  What changes is where the math runs:

  encode / forward : one fused tcgen05 GEMM + TopK kernel, the dense [tokens, num_latents] tensor is never written
  decode           : gather of the k selected W_dec rows
  pre_acts         : same GEMM kernel with a dense store (for callers that really want the dense tensor)

Everything needs CUDA tensors; there is no CPU path.
"""
from __future__ import annotations

import json
from fnmatch import fnmatch
from pathlib import Path
from typing import NamedTuple, Optional, Union

import torch
from safetensors.torch import load_model, save_model
from torch import Tensor, nn

from saeb200 import engine
from saeb200._capi import SaebError

from .._compat import natsorted
from .config import SaeConfig
from .utils import decoder_impl


class EncoderOutput(NamedTuple):
    top_acts: Tensor  # [..., k] activations of the selected latents
    top_indices: Tensor  # [..., k] int64 latent ids


class ForwardOutput(NamedTuple):
    sae_out: Tensor
    latent_acts: Tensor
    latent_indices: Tensor
    fvu: Tensor  # fraction of variance unexplained
    auxk_loss: Tensor  # always 0 at inference
    multi_topk_fvu: Tensor  # always 0 at inference


class Sae(nn.Module):
    def __init__(self, d_in: int, cfg: SaeConfig, device: Union[str, torch.device] = "cpu",
                 dtype: Optional[torch.dtype] = None, *, decoder: bool = True):
        super().__init__()
        self.cfg = cfg
        self.d_in = d_in
        self.num_latents = cfg.num_latents or d_in * cfg.expansion_factor
        self.encoder = nn.Linear(d_in, self.num_latents, device=device, dtype=dtype)
        with torch.no_grad():
            self.encoder.bias.zero_()
        self.W_dec = nn.Parameter(self.encoder.weight.detach().clone()) if decoder else None
        if decoder and cfg.normalize_decoder:
            self.set_decoder_norm_to_unit_norm()
        self.b_dec = nn.Parameter(torch.zeros(d_in, dtype=dtype, device=device))
        # device-side repack of the encoder, rebuilt when parameters change.  encoder_planes selects the parity-grade
        # mode of `encode`: 3 = one fp16 tensor-core pass + exact fp32 refinement (default), 2 = bf16 hi+lo (two passes),
        # 4 = mode 3 with the refinement done by residual correction (half the gather bytes, values to ~1e-6)
        self.encoder_planes = 3
        # which TopK values the refinement re-evaluates exactly (modes 3 / 4): "boundary" = the index set is decided
        # rigorously, exact fp32 values only where they decide it, the other members keep the tensor-core value (fp16
        # W_enc: ~5e-5 relative at d = 4096, inside the 1e-3 bar; ~7x fewer gathered rows); "all" = every value exact
        self.refine_values = "boundary"
        self._packed = {}
        self._overlap = None
        self.overlap_chunk = 9472  # one wave of 37 token tiles: one GEMM launch per pipeline chunk

    # ------------------------------------------------------------------ loading / saving
    @staticmethod
    def load_many(name: str, local: bool = False, layers: Optional[list] = None,
                  device: Union[str, torch.device] = "cpu", *, decoder: bool = True,
                  pattern: Optional[str] = None) -> "dict[str, Sae]":
        glob = pattern + "/*" if pattern is not None else None
        if local:
            root = Path(name)
        else:
            from huggingface_hub import snapshot_download

            root = Path(snapshot_download(name, allow_patterns=glob))
        if layers is not None:
            return {layer: Sae.load_from_disk(root / layer, device=device, decoder=decoder)
                    for layer in natsorted(layers)}
        dirs = [p for p in root.iterdir() if p.is_dir() and (glob is None or fnmatch(p.name, glob))]
        return {p.name: Sae.load_from_disk(p, device=device, decoder=decoder)
                for p in natsorted(dirs, key=lambda p: p.name)}

    @staticmethod
    def load_from_hub(name: str, hookpoint: Optional[str] = None, device: Union[str, torch.device] = "cpu", *,
                      decoder: bool = True) -> "Sae":
        from huggingface_hub import snapshot_download

        root = Path(snapshot_download(name, allow_patterns=f"{hookpoint}/*" if hookpoint is not None else None))
        if hookpoint is not None:
            root = root / hookpoint
        elif not (root / "cfg.json").exists():
            raise FileNotFoundError("No config file found; try specifying a layer.")
        return Sae.load_from_disk(root, device=device, decoder=decoder)

    @staticmethod
    def load_from_disk(path: Union[Path, str], device: Union[str, torch.device] = "cpu", *,
                       decoder: bool = True) -> "Sae":
        path = Path(path)
        with open(path / "cfg.json") as fh:
            raw = json.load(fh)
        d_in = raw.pop("d_in")
        known = {f for f in SaeConfig.__dataclass_fields__}
        cfg = SaeConfig(**{k: v for k, v in raw.items() if k in known})
        sae = Sae(d_in, cfg, device=device, decoder=decoder)
        load_model(model=sae, filename=str(path / "sae.safetensors"), device=str(device), strict=decoder)
        return sae

    def save_to_disk(self, path: Union[Path, str]) -> None:
        path = Path(path)
        path.mkdir(parents=True, exist_ok=True)
        save_model(self, str(path / "sae.safetensors"))
        with open(path / "cfg.json", "w") as fh:
            json.dump({**self.cfg.to_dict(), "d_in": self.d_in}, fh)

    @property
    def device(self):
        return self.encoder.weight.device

    @property
    def dtype(self):
        return self.encoder.weight.dtype

    # ------------------------------------------------------------------ engine plumbing
    def invalidate_packed(self) -> None:
        """Drop the device-side repacks of the encoder.  They are rebuilt automatically when a parameter is replaced or
        modified in place through autograd-visible ops (the cache key holds data_ptr and `_version`); writes through
        `.data` (e.g. `sae.encoder.weight.data.mul_(2)`) bump neither -- call this after such edits.  Loading a state
        dict does it for you."""
        self._packed.clear()
        self._overlap = None

    def _load_from_state_dict(self, *args, **kwargs):
        super()._load_from_state_dict(*args, **kwargs)
        self.invalidate_packed()

    def packed_encoder(self, planes: Optional[int] = None) -> engine.PackedEncoder:
        planes = self.encoder_planes if planes is None else planes
        w, b, bd = self.encoder.weight, self.encoder.bias, self.b_dec
        key = (w.data_ptr(), w._version, b.data_ptr(), b._version, bd.data_ptr(), bd._version)
        hit = self._packed.get(planes)
        if hit is None or hit[0] != key:
            if not w.is_cuda:
                raise SaebError("Sae lives on the CPU: the B200 engine has no CPU path, move it to a CUDA device")
            self._packed[planes] = (key, engine.PackedEncoder.pack(w.data, b.data, bd.data, planes))
        return self._packed[planes][1]

    # ------------------------------------------------------------------ reference API
    def pre_acts(self, x: Tensor) -> Tensor:
        """Dense relu(W_enc (x - b_dec) + b_enc), shape [..., num_latents] fp32 (reference sae/sae.py:172-177).
        Kept for callers that need the dense tensor; `encode` / `forward` never materialise it."""
        _, _, dense = engine.encode_topk(x, self.packed_encoder(2), self.cfg.k, want_dense=True, want_topk=False)
        return dense

    def select_topk(self, latents: Tensor) -> EncoderOutput:
        """Top-k of an already dense latent tensor (reference sae/sae.py:179-181).  Only reached by callers that
        built the dense tensor themselves (e.g. after editing it); the fused path is `encode`."""
        return EncoderOutput(*engine.dense_topk(latents, self.cfg.k))

    def encode(self, x: Tensor, *, clamp_feature: int = -1, clamp_value: float = 0.0,
               exact_values: Optional[bool] = None) -> EncoderOutput:
        """Fused encoder GEMM + TopK (reference sae/sae.py:183-185).  Rows come back ordered by
        (activation desc, index asc); the reference's order is unspecified (`sorted=False`).  `exact_values` overrides
        `self.refine_values` for this call (the activation cache asks for exact values: it ranks by them)."""
        vm = self._value_mode(exact_values)
        # differentiable when the INPUT carries a gradient (attribution patching through several spliced SAEs); the
        # encoder's own parameter gradients are a training matter: opt in with `self.train_encoder = True`
        if torch.is_grad_enabled() and (x.requires_grad or getattr(self, "train_encoder", False)):
            from .utils import SparseEncode   # differentiable like the reference's nn.Linear -> relu -> topk

            acts, idx = SparseEncode.apply(x, self.encoder.weight, self.encoder.bias, self.b_dec,
                                           self.packed_encoder(), self.cfg.k, int(clamp_feature), float(clamp_value),
                                           vm)
            return EncoderOutput(acts, idx)
        acts, idx, _ = engine.encode_topk(x, self.packed_encoder(), self.cfg.k, clamp_feature=clamp_feature,
                                          clamp_value=clamp_value, value_mode=vm)
        return EncoderOutput(acts, idx)

    def _value_mode(self, exact_values: Optional[bool] = None) -> int:
        if exact_values is None:
            exact_values = getattr(self, "refine_values", "boundary") == "all"
        return engine.VALUES_EXACT if exact_values else engine.VALUES_BOUNDARY

    def decode(self, top_acts: Tensor, top_indices: Tensor) -> Tensor:
        """sum_j acts_j W_dec[idx_j] + b_dec (reference sae/sae.py:187-191)."""
        assert self.W_dec is not None, "Decoder weight was not initialized."
        y = decoder_impl(top_indices, top_acts.to(self.dtype), self.W_dec.mT)
        return y + self.b_dec

    def forward(self, x: Tensor, dead_mask: Optional[Tensor] = None) -> ForwardOutput:
        """Inference forward (reference sae/sae.py:193-247): sae_out, top-k latents and FVU.  The AuxK
        (`dead_mask`) and Multi-TopK branches are training losses and are not part of this engine."""
        if dead_mask is not None or self.cfg.multi_topk:
            raise NotImplementedError("AuxK / Multi-TopK losses are training-only and outside the inference engine")
        assert self.W_dec is not None, "Decoder weight was not initialized."
        x2 = x.reshape(-1, self.d_in)
        T = x2.shape[0]
        if self.encoder_planes in (3, 4) and T >= 2 * self.overlap_chunk and x2.is_cuda:
            # large batches: GEMM of chunk c+1 overlaps the HBM-bound refinement + decode of chunk c (two streams)
            from saeb200.overlap import OverlappedForward

            enc = self.packed_encoder()
            if (self._overlap is None or self._overlap.enc is not enc
                    or self._overlap.value_mode != self._value_mode()):
                self._overlap = OverlappedForward(enc, self.W_dec.data, self.b_dec.data, self.cfg.k, self.overlap_chunk,
                                                  value_mode=self._value_mode())
            xin = engine._as_2d(x2, self.d_in)
            top_acts = torch.empty((T, self.cfg.k), dtype=torch.float32, device=x.device)
            top_indices = torch.empty((T, self.cfg.k), dtype=torch.int64, device=x.device)
            sae_out = torch.empty((T, self.d_in), dtype=torch.float32, device=x.device)
            sq_err = torch.zeros((), dtype=torch.float64, device=x.device)
            self._overlap.run(xin, top_acts, top_indices, sae_out, sq_err)
            top_acts = top_acts.view(*x.shape[:-1], self.cfg.k)
            top_indices = top_indices.view(*x.shape[:-1], self.cfg.k)
            sae_out = sae_out.view(*x.shape)
        else:
            top_acts, top_indices = self.encode(x)
            sq_err = torch.zeros((), dtype=torch.float64, device=top_acts.device)
            sae_out = engine.decode(top_indices, top_acts, self.W_dec.data, self.b_dec.data, out_dtype=torch.float32,
                                    x=x, sq_err=sq_err)
        # reference: (x - x.mean(0)).pow(2).sum() -- the mean is over the FIRST dimension only (sae/sae.py:204), so a
        # [batch, seq, d] input is centred per (position, channel), i.e. as a [batch, seq * d] matrix
        total_variance = engine.total_variance(x if x.dim() <= 2 else x.reshape(x.shape[0], -1))
        fvu = (sq_err / total_variance).to(torch.float32)
        if sae_out.dtype != self.dtype:   # the reference computes in the SAE's dtype (sae/sae.py:172, :187-191)
            sae_out = sae_out.to(self.dtype)
        zero = sae_out.new_tensor(0.0)
        return ForwardOutput(sae_out, top_acts, top_indices, fvu, zero, zero)

    # ------------------------------------------------------------------ parameter utilities (host-side, unchanged)
    @torch.no_grad()
    def set_decoder_norm_to_unit_norm(self) -> None:
        assert self.W_dec is not None, "Decoder weight was not initialized."
        eps = torch.finfo(self.W_dec.dtype).eps
        self.W_dec.data /= torch.norm(self.W_dec.data, dim=1, keepdim=True) + eps

    @torch.no_grad()
    def remove_gradient_parallel_to_decoder_directions(self) -> None:
        assert self.W_dec is not None and self.W_dec.grad is not None
        along = (self.W_dec.grad * self.W_dec.data).sum(dim=1, keepdim=True)
        self.W_dec.grad -= along * self.W_dec.data
