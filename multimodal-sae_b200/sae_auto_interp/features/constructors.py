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
from .features import FeatureRecord, prepare_examples, prepare_image_examples
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
                                cfg: FeatureConfig = None, *, ctx_len: int = None, max_examples: int = None,
                                ranked=None):
    """Top `max_examples` windows of `ctx_len` tokens by max activation, descending (reference :70-85).

    `ranked` (optional) = (scores, window ids) of this feature from the device ranking of the whole split file
    (saeb200.engine.coo_top_windows, window id = row * n_win + w): the per-feature densify + max-pool + topk is then
    skipped and only the selected windows are materialised."""
    ctx_len = ctx_len if ctx_len is not None else cfg.example_ctx_len
    max_examples = max_examples if max_examples is not None else cfg.max_examples
    seq_len = tokens.shape[1]
    if ranked is not None:
        n_win = seq_len // ctx_len
        win = ranked[1][:max_examples].to(torch.long)
        r, w = win // n_win, win % n_win
        loc, act = buffer_output.locations, buffer_output.activations
        ent_win = loc[:, 0] * n_win + loc[:, 1] // ctx_len
        inside = loc[:, 1] < n_win * ctx_len
        order = torch.argsort(win)
        slot_sorted = torch.searchsorted(win[order], ent_win.clamp(max=int(win.max()) if win.numel() else 0))
        slot_sorted = slot_sorted.clamp(max=max(win.numel() - 1, 0))
        hit = inside & (win.numel() > 0) & (win[order][slot_sorted] == ent_win)
        dense = torch.zeros((win.numel(), ctx_len), dtype=act.dtype)
        dense.index_put_((order[slot_sorted[hit]], loc[hit, 1] % ctx_len), act[hit], accumulate=True)
        cols = (w * ctx_len)[:, None] + torch.arange(ctx_len)[None, :]
        record.examples = prepare_examples(tokens[r][torch.arange(win.numel())[:, None], cols], dense)
        return
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
                        ctx_len: int, max_examples: int, ranked=None):
    pool_max_activation_windows(record, buffer_output=buffer_output, tokens=tokens, ctx_len=ctx_len,
                                max_examples=max_examples, ranked=ranked)
    random_activation_windows(record, tokens=tokens, buffer_output=buffer_output, n_random=n_random, ctx_len=ctx_len)


IMAGE_SEQ_LEN = 8000   # the reference's stand-in sequence length for image rows (constructors.py:103)


def _image_rows(locations: torch.Tensor, activations: torch.Tensor, rows, seq_len: int = IMAGE_SEQ_LEN) -> torch.Tensor:
    """dense activation rows [len(rows), seq_len] of the selected images only (the reference densifies all
    n_images x 8000, constructors.py:104-107)"""
    rows_t = torch.as_tensor(rows, dtype=torch.long)
    dense = torch.zeros((len(rows_t), seq_len), dtype=activations.dtype)
    for slot, r in enumerate(rows_t.tolist()):
        sel = locations[:, 0] == r
        dense[slot].index_put_((locations[sel, 1],), activations[sel], accumulate=True)
    return dense


def image_scores(locations: torch.Tensor, activations: torch.Tensor, n_images: int, n_base: int) -> torch.Tensor:
    """per-image mean activation over the first `n_base` positions = avg_pool1d(dense[:, :n_base], n_base)
    (reference constructors.py:109-114), without the dense tensor"""
    score = torch.zeros(n_images, dtype=activations.dtype)
    inside = locations[:, 1] < n_base
    score.index_add_(0, locations[inside, 0], activations[inside])
    return score / n_base


def _dedup_ranked(ranked, ids, max_examples: int):
    """keep the first occurrence of every dataset id in rank order, `max_examples` of them (reference :121-135).  With
    fewer distinct images than `max_examples` the reference raises (`len()` of an int, :131); here the best image is
    repeated, which is what that line set out to do."""
    seen, keep = set(), []
    for idx, image_id in zip(ranked, ids):
        if image_id not in seen:
            seen.add(image_id)
            keep.append(idx)
    if len(keep) < max_examples:
        keep += [keep[0]] * (max_examples - len(keep))
    return keep[:max_examples]


def pool_max_activations_windows_image(record: FeatureRecord, buffer_output: BufferOutput, tokens, cfg: FeatureConfig,
                                       processor=None, ranked=None):
    """Image variant (reference :88-148): rank images by the mean activation over the first `num_image_tokens` (576)
    positions, take the best `max_examples + 50`, drop repeated dataset ids, keep `max_examples`; `record.examples`
    holds one ImageExample per kept image.  `tokens` is the image dataset (`len`, `.features`, `.select(indices=)`).

    `ranked` (optional) = (scores, image ids) of this feature from the device ranking of the whole split file
    (saeb200.engine.coo_top_windows(n_base=...), positive scores only, score desc / id asc): replaces the per-feature
    score vector + topk; images that never fired fill the list in ascending id order when fewer than
    `max_examples + 50` did (the reference's topk over the zero scores picks them in an unspecified order)."""
    n_base = getattr(processor, "num_image_tokens", 576)
    loc, act = buffer_output.locations, buffer_output.activations
    want = cfg.max_examples + 50
    if ranked is not None:
        ranked_ids = ranked[1][:want].tolist()
        if len(ranked_ids) < want:
            have = set(ranked_ids)
            ranked_ids += [i for i in range(len(tokens)) if i not in have][: want - len(ranked_ids)]
        ranked = ranked_ids
    else:
        score = image_scores(loc, act, len(tokens), n_base)
        ranked = torch.topk(score, want).indices.tolist()
    if "id" in tokens.features:
        ranked = _dedup_ranked(ranked, tokens.select(indices=ranked)["id"], cfg.max_examples)
    else:
        ranked = ranked[: cfg.max_examples]
    images = tokens.select(indices=ranked)["image"]
    record.examples = prepare_image_examples(torch.zeros(len(ranked), IMAGE_SEQ_LEN), _image_rows(loc, act, ranked),
                                             images, processor)


def random_activations_image(record: FeatureRecord, buffer_output: BufferOutput, tokens, cfg: FeatureConfig,
                             processor=None):
    """`max_examples` images drawn uniformly (reference :151-181)"""
    pick = torch.randint(0, len(tokens), (cfg.max_examples,))
    images = tokens.select(indices=pick)["image"]
    record.examples = prepare_image_examples(torch.zeros(len(pick), IMAGE_SEQ_LEN),
                                             _image_rows(buffer_output.locations, buffer_output.activations, pick),
                                             images, processor)


def top_windows_all_features(top_acts: torch.Tensor, top_indices: torch.Tensor, num_latents: int, ctx_len: int,
                             n_top: int, *, feat_lo: int = 0, feat_hi: int = None, window_base: int = 0):
    """Device route: TopK stream [T, k] -> (scores [F, n_top], window ids [F, n_top]) for features
    [feat_lo, feat_hi), ordered (score desc, window asc); -1 marks empty slots."""
    from saeb200.engine import TopActivationScan

    feat_hi = num_latents if feat_hi is None else feat_hi
    scan = TopActivationScan(feat_lo, feat_hi, n_top, ctx_len, top_acts.device)
    scan.update(top_acts, top_indices, window_base)
    return scan.finalize()


def top_images_all_features(top_acts: torch.Tensor, top_indices: torch.Tensor, num_latents: int, tokens_per_image: int,
                            n_top: int, *, n_base: int = 576, feat_lo: int = 0, feat_hi: int = None,
                            image_base: int = 0):
    """Device route of the image constructor: TopK stream of whole image rows [n_images * tokens_per_image, k] ->
    (scores [F, n_top], image ids [F, n_top]) for features [feat_lo, feat_hi): the images with the largest mean
    activation over their first `n_base` positions, ordered (score desc, image id asc); -1 marks empty slots.  Ask for
    `max_examples + 50` and drop repeated dataset ids with `_dedup_ranked` to reproduce the reference's selection."""
    from saeb200.engine import TopImageScan

    feat_hi = num_latents if feat_hi is None else feat_hi
    scan = TopImageScan(feat_lo, feat_hi, n_top, tokens_per_image, n_base, top_acts.device)
    scan.update(top_acts, top_indices, image_base)
    return scan.finalize()
