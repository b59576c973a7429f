"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference hot path.

This is the *oracle* the CUDA engine is checked against.  It restates, on CPU
(torch fp32 tensors / numpy integer arrays), the algorithm of
EvolvingLMMs-Lab/multimodal-sae for the SAE encode -> TopK -> sparse decode
path, the activation-cache extraction and the per-feature top-activation scan.
Every function cites the reference file:line it follows (paths relative to the
reference root).

Parity status: PINNED.  `tests/test_oracle_golden.py` checks every function
below against fixtures under `tests/golden/` that were produced by running the
unmodified reference in the build container (`oracle/gen_golden.py`, using the
import shims of `oracle/ref_shims.py`).  The reference itself ships exactly one
test (train/sae/tests/test_decode.py:6-20: sparse decode == scatter + dense
matmul); that property is restated in `tests/test_oracle_golden.py` too.

Third-party arithmetic: the reference's GEMM / TopK / nonzero / isin live in
PyTorch (`torch>=2.1.0`, pyproject.toml:29; torch 2.11.0+cu128 is what is
installed here and on the GPU box).  The oracle calls the same torch CPU ops
for fp32 arithmetic (so summation order is MKL's, like the reference on CPU)
and offers an fp64 evaluation (`pre_acts_f64`) for tie audits.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline /
`--impl reference` leg may import this module.  The product never does.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, List, NamedTuple, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F


# ---------------------------------------------------------------------------
# SAE parameters (reference sae/sae.py:44-66)
# ---------------------------------------------------------------------------
@dataclass
class SaeParams:
    """`encoder.weight [N,d]`, `encoder.bias [N]`, `W_dec [N,d]`, `b_dec [d]`, all fp32."""

    W_enc: torch.Tensor
    b_enc: torch.Tensor
    W_dec: torch.Tensor
    b_dec: torch.Tensor
    k: int

    @property
    def d_in(self) -> int:
        return self.W_enc.shape[1]

    @property
    def num_latents(self) -> int:
        return self.W_enc.shape[0]


def init_params(d_in: int, num_latents: int, k: int, seed: int, *, bias_std: float = 0.01,
                b_dec_std: float = 0.1) -> SaeParams:
    """Synthetic weights with the reference constructor's semantics (sae/sae.py:59-66):
    `encoder.weight` ~ nn.Linear default init U(-1/sqrt(d), 1/sqrt(d)); `W_dec` = clone of the
    encoder weight with rows normalised to unit norm (sae/sae.py:62-64, :249-255).  The reference
    zero-inits both biases (:60, :66); trained checkpoints have non-zero ones, so the synthetic
    biases are small Gaussians to exercise the bias path."""
    g = torch.Generator(device="cpu").manual_seed(seed)
    bound = 1.0 / (d_in ** 0.5)
    W_enc = (torch.rand(num_latents, d_in, generator=g, dtype=torch.float32) * 2 - 1) * bound
    b_enc = torch.randn(num_latents, generator=g, dtype=torch.float32) * bias_std
    b_dec = torch.randn(d_in, generator=g, dtype=torch.float32) * b_dec_std
    W_dec = W_enc.clone()
    eps = torch.finfo(W_dec.dtype).eps
    W_dec /= torch.norm(W_dec, dim=1, keepdim=True) + eps  # sae/sae.py:253-255
    return SaeParams(W_enc, b_enc, W_dec, b_dec, k)


# ---------------------------------------------------------------------------
# encode / TopK / decode / forward (reference sae/sae.py:172-247)
# ---------------------------------------------------------------------------
def pre_acts(p: SaeParams, x: torch.Tensor) -> torch.Tensor:
    """reference sae/sae.py:172-177: relu(Linear(x.to(fp32) - b_dec))."""
    sae_in = x.to(torch.float32) - p.b_dec
    return F.relu(F.linear(sae_in, p.W_enc, p.b_enc))


def pre_acts_f64(p: SaeParams, x: torch.Tensor) -> torch.Tensor:
    """Same formula evaluated in fp64 (tie audit: which rows have a k/(k+1) gap below fp32 noise)."""
    sae_in = x.to(torch.float64) - p.b_dec.double()
    return F.relu(F.linear(sae_in, p.W_enc.double(), p.b_enc.double()))


class EncoderOutput(NamedTuple):
    top_acts: torch.Tensor
    top_indices: torch.Tensor


def select_topk(latents: torch.Tensor, k: int) -> EncoderOutput:
    """reference sae/sae.py:179-181: `latents.topk(k, sorted=False)`; int64 indices; order unspecified."""
    return EncoderOutput(*latents.topk(k, sorted=False))


def encode(p: SaeParams, x: torch.Tensor) -> EncoderOutput:
    """reference sae/sae.py:183-185."""
    return select_topk(pre_acts(p, x), p.k)


def eager_decode(top_indices: torch.Tensor, top_acts: torch.Tensor, W_dec_T: torch.Tensor) -> torch.Tensor:
    """reference sae/utils.py:108-111 (`W_dec_T` is `W_dec.mT`, shape [d, N])."""
    buf = top_acts.new_zeros(top_acts.shape[:-1] + (W_dec_T.shape[-1],))
    acts = buf.scatter_(dim=-1, index=top_indices, src=top_acts)
    return acts @ W_dec_T.mT


def sparse_decode(top_indices: torch.Tensor, top_acts: torch.Tensor, W_dec: torch.Tensor) -> torch.Tensor:
    """Restatement of the Triton forward decode kernel (reference sae/kernels.py:222-284):
    out[a,:] = sum_{j=0..K-1, v!=0} vals[a,j] * W_dec[idx[a,j], :], fp32 accumulate in j order,
    zero values skipped (:277), result cast to vals.dtype (:282-284)."""
    A, K = top_indices.shape
    out = torch.zeros(A, W_dec.shape[1], dtype=torch.float32)
    for j in range(K):
        v = top_acts[:, j].to(torch.float32)
        rows = W_dec[top_indices[:, j]].to(torch.float32)
        out += torch.where(v[:, None] != 0, v[:, None] * rows, torch.zeros_like(rows))
    return out.to(top_acts.dtype)


def decode_backward(top_indices: torch.Tensor, top_acts: torch.Tensor, W_dec: torch.Tensor, grad_out: torch.Tensor
                    ) -> Tuple[torch.Tensor, torch.Tensor]:
    """Backward of the decoder seam, reference `TritonDecoder.backward` (sae/kernels.py:411-429):
      d_acts[a, j] = grad_out[a, :] . W_dec[idx[a, j], :]          (`triton_dense_dense_sparseout_matmul`, :287-400:
                                                                    "(dense1 @ dense2).gather(1, at_indices)")
      dW_dec[n, :] = sum_{(a, j): idx[a, j] == n} acts[a, j] * grad_out[a, :]   (`triton_sparse_transpose_dense_matmul`,
                                                                    :10-175: sparse.T @ dense, zero values skipped)
    Both in fp32.  The same numbers come out of autograd through `eager_decode` (sae/utils.py:108-111), which is what
    the golden fixture was generated with."""
    A, K = top_indices.shape
    g = grad_out.to(torch.float32)
    d_acts = (g @ W_dec.to(torch.float32).T).gather(1, top_indices)
    dW = torch.zeros(W_dec.shape, dtype=torch.float32)
    dW.index_add_(0, top_indices.reshape(-1),
                  (top_acts.to(torch.float32)[:, :, None] * g[:, None, :]).reshape(A * K, -1))
    return d_acts, dW


def decode(p: SaeParams, top_acts: torch.Tensor, top_indices: torch.Tensor) -> torch.Tensor:
    """reference sae/sae.py:187-191: decoder_impl(idx, acts.to(dtype), W_dec.mT) + b_dec."""
    y = eager_decode(top_indices, top_acts.to(torch.float32), p.W_dec.mT)
    return y + p.b_dec


class ForwardOutput(NamedTuple):
    sae_out: torch.Tensor
    latent_acts: torch.Tensor
    latent_indices: torch.Tensor
    fvu: torch.Tensor
    auxk_loss: torch.Tensor
    multi_topk_fvu: torch.Tensor


def forward(p: SaeParams, x: torch.Tensor) -> ForwardOutput:
    """reference sae/sae.py:193-247 with dead_mask=None, multi_topk=False (inference)."""
    pa = pre_acts(p, x)
    top_acts, top_indices = select_topk(pa, p.k)
    sae_out = decode(p, top_acts, top_indices)
    e = sae_out - x
    total_variance = (x - x.mean(0)).pow(2).sum()  # :204
    fvu = e.pow(2).sum() / total_variance  # :229-230
    zero = sae_out.new_tensor(0.0)
    return ForwardOutput(sae_out, top_acts, top_indices, fvu, zero, zero)


# ---------------------------------------------------------------------------
# canonical forms used for comparisons
# ---------------------------------------------------------------------------
def canonical_topk(top_acts: torch.Tensor, top_indices: torch.Tensor) -> Tuple[np.ndarray, np.ndarray]:
    """Sort each row's (idx, val) pairs by index; `topk(sorted=False)` order is unspecified
    (reference sae/sae.py:181), so index *sets* are what parity means."""
    idx = top_indices.cpu().numpy().astype(np.int64)
    val = top_acts.detach().cpu().to(torch.float32).numpy()
    order = np.argsort(idx, axis=-1, kind="stable")
    return np.take_along_axis(idx, order, -1), np.take_along_axis(val, order, -1)


def tie_audit(p: SaeParams, x: torch.Tensor, k: int, rel_tol: float) -> np.ndarray:
    """Rows whose fp64 k-th / (k+1)-th pre-activation gap is below `rel_tol` * k-th value: on those rows
    two correct fp32 implementations may legitimately pick different boundary elements."""
    pa = pre_acts_f64(p, x)
    top = pa.topk(k + 1, sorted=True).values
    kth, nxt = top[:, k - 1], top[:, k]
    return ((kth - nxt) <= rel_tol * kth.abs().clamp_min(1e-30)).numpy()


# ---------------------------------------------------------------------------
# cache path (reference features/cache.py)
# ---------------------------------------------------------------------------
def topk_masked_latents(p: SaeParams, hidden: torch.Tensor) -> torch.Tensor:
    """reference features/cache.py:206-218 (and :399-417): pre_acts -> torch.topk(k) -> zeros_like +
    scatter_: dense TopK-masked latents [bs, seq, N]."""
    latents = pre_acts(p, hidden)
    topk = torch.topk(latents, k=p.k, dim=-1)
    result = torch.zeros_like(latents)
    result.scatter_(-1, topk.indices, topk.values)
    return result


def get_nonzeros(latents: torch.Tensor, selected_features: Optional[torch.Tensor] = None
                 ) -> Tuple[torch.Tensor, torch.Tensor]:
    """reference features/cache.py:73-92: nonzero(|x| > 1e-5) -> (batch, pos, feature) int64 triples in
    row-major order + fp32 activations; optional `torch.isin(feature, filter)`."""
    mask = latents.abs() > 1e-5
    loc = torch.nonzero(mask)
    act = latents[mask]
    if selected_features is None:
        return loc, act
    keep = torch.isin(loc[:, 2], selected_features)
    return loc[keep], act[keep]


def cache_add_offset(loc: torch.Tensor, batch_number: int, batch_size: int, shard_size: int) -> torch.Tensor:
    """reference features/cache.py:55: locations[:, 0] += batch_number * batch_size + shard_size."""
    loc = loc.clone()
    loc[:, 0] += batch_number * batch_size + shard_size
    return loc


def generate_split_indices(width: int, n_splits: int) -> List[Tuple[int, int]]:
    """reference features/cache.py:243-247: linspace boundaries, `end = boundary - 1`."""
    b = torch.linspace(0, width, steps=n_splits + 1).long()
    return [(int(s), int(e) - 1) for s, e in zip(b[:-1], b[1:])]


def split_mask(features: torch.Tensor, start: int, end: int) -> torch.Tensor:
    """reference features/cache.py:293-294: `(features >= start) & (features < end)` -- together with
    `end = boundary - 1` this silently drops the last feature id of every split (reference quirk,
    SURVEY.md appendix A.1)."""
    return (features >= start) & (features < end)


def tensorbuffer_select(locations: torch.Tensor, activations: torch.Tensor, feature: int
                        ) -> Tuple[torch.Tensor, torch.Tensor]:
    """reference features/loader.py:74-90: mask = locations[:,2]==feature -> (locations[:, :2], acts)."""
    mask = locations[:, 2] == feature
    return locations[mask][:, :2], activations[mask]


def dataset_bucket_paths(width: int, n_splits: int, selected: torch.Tensor) -> Dict[Tuple[int, int], torch.Tensor]:
    """reference features/loader.py:164-196: route selected features to split files with
    bucketize(right=True); returns {(start, end_inclusive): features}."""
    edges = torch.linspace(0, width, steps=n_splits + 1).long()
    buck = torch.bucketize(selected, edges, right=True)
    out = {}
    for b in torch.unique(buck):
        m = buck == b
        out[(int(edges[b - 1]), int(edges[b]) - 1)] = selected[m]
    return out


# ---------------------------------------------------------------------------
# top-activation scan (reference features/constructors.py)
# ---------------------------------------------------------------------------
def pool_max_activation_windows(locations: torch.Tensor, activations: torch.Tensor, tokens: torch.Tensor,
                                ctx_len: int, max_examples: int):
    """reference features/constructors.py:11-85: densify [rows, seq], keep rows that fired (:19-22),
    max_pool1d(ctx_len) (:31-33), topk(min(max_examples, #nonzero pools)) windows (:55-61).
    Returns (token_windows, activation_windows, window_ids, pooled) where window_ids are
    (row_in_dataset, window_in_row) pairs -- the ids are what the GPU scan reproduces."""
    batch_len, seq_len = tokens.shape
    dense = torch.sparse_coo_tensor(locations.t(), activations, (batch_len, seq_len)).to_dense()
    uniq = torch.unique(locations[:, 0])
    token_batches = tokens[uniq]
    dense = dense[uniq]
    pools = F.max_pool1d(dense, kernel_size=ctx_len, stride=ctx_len)
    act_windows = dense.unfold(1, ctx_len, ctx_len).reshape(-1, ctx_len)
    tok_windows = token_batches.unfold(1, ctx_len, ctx_len).reshape(-1, ctx_len)
    nz = pools != 0
    k = min(max_examples, int(nz.sum()))
    top = torch.topk(pools.flatten(), k)
    n_win = seq_len // ctx_len
    rows = uniq[top.indices // n_win]
    wins = top.indices % n_win
    return tok_windows[top.indices], act_windows[top.indices], torch.stack([rows, wins], 1), top.values


def scan_top_windows(top_acts: torch.Tensor, top_indices: torch.Tensor, num_latents: int, ctx_len: int,
                     n_top: int) -> Tuple[np.ndarray, np.ndarray]:
    """End-to-end restatement of the scan the GPU engine performs on TopK output: tokens are rows of a
    flat [T] stream viewed as windows of `ctx_len`; for every feature the pooled score of a window is the
    max TopK-masked activation (> 1e-5, features/cache.py:80) inside it (constructors.py:31-33) and the
    result is each feature's `n_top` best windows (constructors.py:55-61), ordered by
    (score desc, window id asc).  Returns (scores [N, n_top] f32, window ids [N, n_top] i64; -1 = empty)."""
    T, k = top_acts.shape
    n_win = T // ctx_len
    vals = top_acts[: n_win * ctx_len].to(torch.float32).numpy().reshape(n_win, ctx_len * k)
    idx = top_indices[: n_win * ctx_len].numpy().reshape(n_win, ctx_len * k)
    scores = np.zeros((num_latents, n_top), np.float32)
    wins = np.full((num_latents, n_top), -1, np.int64)
    per_feature: Dict[int, List[Tuple[float, int]]] = {}
    for w in range(n_win):
        v, f = vals[w], idx[w]
        keep = v > 1e-5
        v, f = v[keep], f[keep]
        if v.size == 0:
            continue
        order = np.lexsort((-v, f))
        f_s, v_s = f[order], v[order]
        first = np.ones(f_s.shape, bool)
        first[1:] = f_s[1:] != f_s[:-1]
        for ff, vv in zip(f_s[first].tolist(), v_s[first].tolist()):
            per_feature.setdefault(ff, []).append((vv, w))
    for ff, lst in per_feature.items():
        lst.sort(key=lambda t: (-t[0], t[1]))
        for j, (vv, w) in enumerate(lst[:n_top]):
            scores[ff, j] = vv
            wins[ff, j] = w
    return scores, wins


# ---------------------------------------------------------------------------
# steering hook (reference features/steering.py:105-124)
# ---------------------------------------------------------------------------
def steering_hook(p: SaeParams, hidden: torch.Tensor, feature: int, clamp_value: float) -> torch.Tensor:
    """hidden [1, T, d] (fp16 in the launcher) -> pre_acts -> (T != 1) latents[:, :, feature] = k ->
    select_topk -> decode(top_acts[0], top_idx[0]) -> unsqueeze(0).to(fp16); the layer output is
    *replaced* by this reconstruction (:119)."""
    latents = pre_acts(p, hidden)
    if latents.shape[1] != 1:
        latents[:, :, feature] = clamp_value
    top_acts, top_indices = select_topk(latents, p.k)
    return decode(p, top_acts[0], top_indices[0]).unsqueeze(0).to(torch.float16)


# ---------------------------------------------------------------------------
# probing tool (reference tools/probe_activations.py:109-126)
# ---------------------------------------------------------------------------
def probe_mean_topk(p: SaeParams, hidden: torch.Tensor, interval: Sequence[int], drop_first: bool = False
                    ) -> Tuple[torch.Tensor, torch.Tensor]:
    """The tool's hook body: `latents = sae.pre_acts(hidden)` (:112); image-only llama inputs drop the BOS position
    `latents[:, 1:, :]` (:115-116); `topk_indices = latents.squeeze(0).mean(dim=0).topk(k=interval[1]).indices
    [interval[0]:]` (:119-121); `topk_acts = latents[:, :, topk_indices].squeeze(0).permute(1, 0)` (:122).
    hidden [1, T, d] -> (indices [interval[1] - interval[0]], acts [n_features, T'])."""
    latents = pre_acts(p, hidden)
    if drop_first:
        latents = latents[:, 1:, :]
    topk_indices = latents.squeeze(0).mean(dim=0).topk(k=interval[1]).indices[interval[0]:]
    topk_acts = latents[:, :, topk_indices].squeeze(0).permute(1, 0)
    return topk_indices, topk_acts
