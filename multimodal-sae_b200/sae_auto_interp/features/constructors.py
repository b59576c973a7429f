"""Example constructors (reference features/constructors.py): per-feature "top activating examples".

Two routes produce the same ranking:
  * `pool_max_activation_windows` / `random_activation_windows` -- the reference's per-feature host route driven by
    `FeatureDataset.load` (kept for drop-in use; tiny tensors, host-side);
  * `top_windows_all_features` -- the B200 route: one device-side scan over the TopK stream that yields every
    feature's best windows at once (saeb200.engine.TopActivationScan), with no per-feature dense tensors.
"""
from __future__ import annotations

import torch

from ..config import FeatureConfig
from .features import FeatureRecord, prepare_examples
from .loader import BufferOutput


def _feature_window_scores(locations: torch.Tensor, activations: torch.Tensor, seq_len: int, ctx_len: int):
    """Sparse equivalent of densify + max_pool1d (constructors.py:11-33): returns (rows, dense rows [n_rows, seq],
    pooled [n_rows, n_win]) for the rows on which the feature fired."""
    rows, inv = torch.unique(locations[:, 0], return_inverse=True)
    dense = torch.zeros((rows.numel(), seq_len), dtype=activations.dtype)
    dense.index_put_((inv, locations[:, 1]), activations, accumulate=True)
    n_win = seq_len // ctx_len
    pooled = dense[:, : n_win * ctx_len].reshape(rows.numel(), n_win, ctx_len).amax(-1)
    return rows, dense, pooled


def pool_max_activation_windows(record: FeatureRecord, buffer_output: BufferOutput, tokens: torch.Tensor,
                                cfg: FeatureConfig = None, *, ctx_len: int = None, max_examples: int = None):
    """Top `max_examples` windows of `ctx_len` tokens by max activation, descending (reference :70-85)."""
    ctx_len = ctx_len if ctx_len is not None else cfg.example_ctx_len
    max_examples = max_examples if max_examples is not None else cfg.max_examples
    seq_len = tokens.shape[1]
    rows, dense, pooled = _feature_window_scores(buffer_output.locations, buffer_output.activations, seq_len, ctx_len)
    n_win = pooled.shape[1]
    flat = pooled.flatten()
    k = min(max_examples, int((flat != 0).sum()))
    top = torch.topk(flat, k).indices
    r, w = top // n_win, top % n_win
    offs = torch.arange(ctx_len)
    cols = (w * ctx_len)[:, None] + offs[None, :]
    record.examples = prepare_examples(tokens[rows[r]][torch.arange(k)[:, None], cols],
                                       dense[r][torch.arange(k)[:, None], cols])


def random_activation_windows(record, tokens: torch.Tensor, buffer_output: BufferOutput, ctx_len: int,
                              n_random: int):
    """`n_random` windows from rows on which the feature never fired (reference :184-209)."""
    torch.manual_seed(22)
    free = torch.ones(tokens.shape[0], dtype=torch.bool)
    free[buffer_output.locations[:, 0].unique()] = False
    avail = free.nonzero().squeeze(-1)
    pick = avail[torch.randperm(len(avail))[:n_random]]
    toks = tokens[pick, 10: 10 + ctx_len]
    record.random_examples = prepare_examples(toks, torch.zeros_like(toks))


def default_constructor(record: FeatureRecord, tokens: torch.Tensor, buffer_output: BufferOutput, n_random: int,
                        ctx_len: int, max_examples: int):
    pool_max_activation_windows(record, buffer_output=buffer_output, tokens=tokens, ctx_len=ctx_len,
                                max_examples=max_examples)
    random_activation_windows(record, tokens=tokens, buffer_output=buffer_output, n_random=n_random, ctx_len=ctx_len)


def pool_max_activations_windows_image(record: FeatureRecord, buffer_output: BufferOutput, tokens, cfg: FeatureConfig,
                                       processor=None):
    """Image variant (reference :88-148): rank images by the mean activation over the first `num_image_tokens`
    (576) positions.  Returns the ranked image rows in `record.examples` as (row, score) pairs; rendering the
    activation masks onto PIL images is presentation code outside this engine."""
    n_base = getattr(processor, "num_image_tokens", 576)
    loc, act = buffer_output.locations, buffer_output.activations
    n_images = len(tokens)
    score = torch.zeros(n_images, dtype=act.dtype)
    inside = loc[:, 1] < n_base
    score.index_add_(0, loc[inside, 0], act[inside])
    score /= n_base
    k = min(cfg.max_examples, n_images)
    top = torch.topk(score, k)
    record.examples = [(int(i), float(s)) for i, s in zip(top.indices, top.values)]


def random_activations_image(record: FeatureRecord, buffer_output: BufferOutput, tokens, cfg: FeatureConfig,
                             processor=None):
    pick = torch.randint(0, len(tokens), (cfg.max_examples,))
    record.examples = [(int(i), 0.0) for i in pick]


def top_windows_all_features(top_acts: torch.Tensor, top_indices: torch.Tensor, num_latents: int, ctx_len: int,
                             n_top: int, *, feat_lo: int = 0, feat_hi: int = None, window_base: int = 0):
    """Device route: TopK stream [T, k] -> (scores [F, n_top], window ids [F, n_top]) for features
    [feat_lo, feat_hi), ordered (score desc, window asc); -1 marks empty slots."""
    from saeb200.engine import TopActivationScan

    feat_hi = num_latents if feat_hi is None else feat_hi
    scan = TopActivationScan(feat_lo, feat_hi, n_top, ctx_len, top_acts.device)
    scan.update(top_acts, top_indices, window_base)
    return scan.finalize()
