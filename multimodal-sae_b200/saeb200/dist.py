"""Multi-GPU forms of the path, one process per GPU (torchrun), `torch.distributed` for the plumbing.

* token-parallel forward: tokens are independent, so ranks take disjoint token slices with the full SAE replicated and
  no data-path collective (the reference's own strategy, `dataset.shard`, launch/cache/cache_image.py:60-71);
* feature-sharded top-activation scan (the north-star layout): rank r owns SAE features [r*N/R, (r+1)*N/R); every rank
  sees every token; per-feature top lists are rank-local and disjoint, so the job ends with ONE all-gather of the
  per-shard lists.  Because the cache keeps a latent only if it is in the token's *global* top-k
  (features/cache.py:210-218), exact mode additionally all-gathers each chunk's local top-k VALUES ([Tc, k] fp32 per
  rank, 256 B/token/rank) and derives the per-token global k-th value; `exact=False` skips that exchange and ranks by
  the shard-local top-k instead (not reference semantics, reported separately).

The choreography is written against a small `ops` object so that the same code runs on CUDA (`EngineOps`, NCCL) and in
the CPU `gloo` tests (where `tests/` inject oracle-backed ops).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Iterable, Optional, Tuple

import torch
import torch.distributed as dist


def shard_range(num_latents: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous, balanced feature range of `rank` (sizes differ by at most one)."""
    base, rem = divmod(num_latents, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def token_slice(num_tokens: int, world: int, rank: int) -> Tuple[int, int]:
    base, rem = divmod(num_tokens, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


class EngineOps:
    """CUDA implementation of the per-rank steps (thin calls into saeb200.engine)."""

    def __init__(self, W_enc_shard, b_enc_shard, b_dec, feat_lo, feat_hi, n_top, ctx_len, device, planes=2,
                 bucket_cap=256):
        from . import engine

        self.engine = engine
        self.enc = engine.PackedEncoder.pack(W_enc_shard, b_enc_shard, b_dec, planes)
        self.feat_lo, self.feat_hi = feat_lo, feat_hi
        self.scan = engine.TopActivationScan(feat_lo, feat_hi, n_top, ctx_len, device, bucket_cap=bucket_cap)

    def encode_topk(self, x, k):
        vals, idx, _ = self.engine.encode_topk(x, self.enc, k)
        return vals, idx + self.feat_lo  # global feature ids

    def kth_of_gathered(self, gathered):
        return self.engine.kth_of_gathered(gathered)

    def scan_update(self, vals, idx, window_base, tok_thr):
        self.scan.update(vals, idx, window_base, tok_thr)

    def scan_finalize(self):
        return self.scan.finalize()


@dataclass
class ScanResult:
    top_vals: torch.Tensor  # [N, n_top] f32, every rank holds the full table after the final all-gather
    top_win: torch.Tensor  # [N, n_top] i64, -1 = empty


def _all_gather_cat(t: torch.Tensor, group, sizes=None) -> torch.Tensor:
    world = dist.get_world_size(group)
    if sizes is None:
        out = [torch.empty_like(t) for _ in range(world)]
    else:
        out = [torch.empty((s,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device) for s in sizes]
    dist.all_gather(out, t.contiguous(), group=group)
    return out


def sharded_scan(chunks: Iterable[torch.Tensor], ops, k: int, ctx_len: int, num_latents: int, *, exact: bool = True,
                 group=None) -> ScanResult:
    """Feature-sharded scan.  `chunks` yields the SAME token chunks ([Tc, d], Tc a multiple of ctx_len) on every rank."""
    distributed = dist.is_available() and dist.is_initialized()
    world = dist.get_world_size(group) if distributed else 1
    k_local = min(k, ops.feat_hi - ops.feat_lo)
    window_base = 0
    for x in chunks:
        vals, idx = ops.encode_topk(x, k_local)
        vals2 = vals.reshape(-1, k_local)
        tok_thr = None
        if exact and world > 1:
            gathered = torch.stack(_all_gather_cat(vals2, group), 0)  # [R, Tc, k_local]
            # per-token global k-th value among the R*k_local shard-local leaders
            tok_thr = ops.kth_of_gathered(gathered) if k_local == k else _kth_host(gathered, k)
        ops.scan_update(vals2, idx.reshape(-1, k_local), window_base, tok_thr)
        window_base += vals2.shape[0] // ctx_len
    top_vals, top_win = ops.scan_finalize()
    if world > 1:
        sizes = [shard_range(num_latents, world, r) for r in range(world)]
        sizes = [hi - lo for lo, hi in sizes]
        top_vals = torch.cat(_all_gather_cat(top_vals, group, sizes), 0)   # the single end-of-job all-gather
        top_win = torch.cat(_all_gather_cat(top_win, group, sizes), 0)
    return ScanResult(top_vals, top_win)


def _kth_host(gathered: torch.Tensor, k: int) -> torch.Tensor:
    R, T, kl = gathered.shape
    flat = gathered.permute(1, 0, 2).reshape(T, R * kl)
    return flat.topk(k, dim=-1).values[:, -1].contiguous()


def token_parallel_forward(sae, x_local: torch.Tensor):
    """Token-parallel forward: each rank runs the fused forward on its token slice; nothing is communicated."""
    return sae(x_local)
