"""Pin the oracle (`oracle/sae_oracle.py`) against fixtures produced by the unmodified reference
(`oracle/gen_golden.py`).  CPU only."""
import os

import numpy as np
import pytest
import torch

import sae_oracle as O

from conftest import GOLDEN


def _params(g):
    return O.SaeParams(torch.from_numpy(g["W_enc"]), torch.from_numpy(g["b_enc"]), torch.from_numpy(g["W_dec"]),
                       torch.from_numpy(g["b_dec"]), int(g["k"]))


@pytest.mark.parametrize("name", ["forward_c1.npz", "forward_c1_bf16.npz", "forward_wide.npz"])
def test_forward_matches_reference(name):
    """reference Sae.forward (sae/sae.py:193-247): TopK index sets bit-identical, values/sae_out/fvu equal."""
    g = np.load(os.path.join(GOLDEN, name))
    p = _params(g)
    x = torch.from_numpy(g["x"])
    out = O.forward(p, x)
    idx, val = O.canonical_topk(out.latent_acts, out.latent_indices)
    assert np.array_equal(idx, g["top_idx"])
    np.testing.assert_allclose(val, g["top_val"], rtol=1e-6, atol=1e-7)
    np.testing.assert_allclose(out.sae_out.numpy(), g["sae_out"], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(float(out.fvu), float(g["fvu"]), rtol=1e-5)
    pa = O.pre_acts(p, x)
    np.testing.assert_allclose(pa.double().sum(-1).numpy(), g["pre_acts_sum"], rtol=1e-6)
    assert np.array_equal((pa > 0).sum(-1).numpy(), g["pre_acts_nnz"])


def test_init_params_matches_reference_constructor_semantics():
    """W_dec is the row-normalised clone of encoder.weight (sae/sae.py:62-64,:249-255)."""
    p = O.init_params(64, 256, 8, seed=3)
    n = torch.norm(p.W_dec, dim=1)
    np.testing.assert_allclose(n.numpy(), 1.0, rtol=1e-5)
    bound = 1 / 64 ** 0.5
    assert float(p.W_enc.abs().max()) <= bound


def test_decode_property_of_reference_test():
    """train/sae/tests/test_decode.py:6-20: sparse decode == scatter + dense matmul (the reference's only test)."""
    g = np.load(os.path.join(GOLDEN, "decode_test.npz"))
    W_dec = torch.from_numpy(g["W_dec"])
    idx, val = torch.from_numpy(g["top_idx"]), torch.from_numpy(g["top_vals"])
    eager = O.eager_decode(idx, val, W_dec.mT)
    np.testing.assert_allclose(eager.numpy(), g["eager"], rtol=1e-6, atol=1e-6)
    sparse = O.sparse_decode(idx, val, W_dec)
    torch.testing.assert_close(sparse, torch.from_numpy(g["eager"]))  # same tolerance family as assert_allclose


def test_decode_backward_matches_reference_autograd():
    """Gradients of the decoder seam (TritonDecoder.backward, sae/kernels.py:411-429) against autograd through the
    reference's eager_decode / Sae.decode."""
    g = np.load(os.path.join(GOLDEN, "decode_backward.npz"))
    idx, val = torch.from_numpy(g["top_idx"]), torch.from_numpy(g["top_vals"])
    W, go = torch.from_numpy(g["W_dec"]), torch.from_numpy(g["grad_out"])
    assert float(val[3, 2]) == 0.0   # a zero activation still receives its gathered-dot-product gradient
    d_vals, dW = O.decode_backward(idx, val, W, go)
    np.testing.assert_allclose(d_vals.numpy(), g["d_vals"], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(dW.numpy(), g["d_W_dec"], rtol=1e-5, atol=1e-6)
    # through the module the same gradients reach W_dec, and b_dec receives the column sums of grad_out
    np.testing.assert_allclose(g["sae_d_vals"], g["d_vals"], rtol=1e-6, atol=1e-7)
    np.testing.assert_allclose(g["sae_d_W_dec"], g["d_W_dec"], rtol=1e-6, atol=1e-7)
    np.testing.assert_allclose(go.sum(0).numpy(), g["sae_d_b_dec"], rtol=1e-5, atol=1e-6)
    assert np.count_nonzero(np.abs(g["d_W_dec"]).sum(1)) <= idx.numel()   # only selected rows receive gradient


def test_cache_chain_matches_reference():
    """FeatureCache.run -> Cache.add/get_nonzeros (features/cache.py:42-92,206-218)."""
    g = np.load(os.path.join(GOLDEN, "cache_chain.npz"))
    hidden = torch.from_numpy(np.load(os.path.join(GOLDEN, "cache_chain_hidden.npy")))
    p = _params(g)
    bs, shard = 2, 100
    for tag, filt in (("nofilter", None), ("filter", torch.tensor([1, 5, 15, 16, 31, 40, 63]))):
        locs, acts = [], []
        for b in range(hidden.shape[0] // bs):
            dense = O.topk_masked_latents(p, hidden[b * bs:(b + 1) * bs])
            loc, act = O.get_nonzeros(dense, filt)
            locs.append(O.cache_add_offset(loc, b, bs, shard))
            acts.append(act)
        loc, act = torch.cat(locs), torch.cat(acts)
        assert np.array_equal(loc.numpy(), g[f"{tag}_locations"])
        np.testing.assert_allclose(act.numpy(), g[f"{tag}_activations"], rtol=1e-6)


def test_split_files_match_reference_including_off_by_one():
    """save_splits / concate_safetensors (features/cache.py:243-309): file names and the silently dropped
    last feature id of every split."""
    g = np.load(os.path.join(GOLDEN, "cache_chain.npz"))
    loc = torch.from_numpy(g["nofilter_locations"])
    act = torch.from_numpy(g["nofilter_activations"])
    splits = O.generate_split_indices(64, 4)
    assert [f"{s}_{e}.safetensors" for s, e in splits] == sorted(g["split_files"].tolist(),
                                                                key=lambda n: int(n.split("_")[0]))
    total = 0
    for s, e in splits:
        m = O.split_mask(loc[:, 2], s, e)
        assert np.array_equal(loc[m].numpy(), g[f"split_{s}_{e}.safetensors_locations"])
        np.testing.assert_array_equal(act[m].numpy(), g[f"split_{s}_{e}.safetensors_activations"])
        total += int(m.sum())
    dropped = int(sum((loc[:, 2] == e).sum() for _, e in splits))
    assert total + dropped == loc.shape[0] and dropped > 0


def test_loader_and_constructor_match_reference():
    """TensorBuffer.__getitem__ (features/loader.py:74-90), FeatureDataset._build_selected (:164-196),
    pool_max_activation_windows (features/constructors.py:11-85)."""
    g = np.load(os.path.join(GOLDEN, "cache_chain.npz"))
    tokens = torch.from_numpy(g["tokens"])
    big = torch.zeros(100 + tokens.shape[0], tokens.shape[1], dtype=torch.long)
    big[100:] = tokens
    sel = torch.tensor([1, 5, 14, 16, 33, 40, 62])
    buckets = O.dataset_bucket_paths(64, 4, sel)
    seen = []
    for (s, e), feats in buckets.items():
        loc = torch.from_numpy(g[f"split_{s}_{e}.safetensors_locations"])
        act = torch.from_numpy(g[f"split_{s}_{e}.safetensors_activations"])
        for f in feats.tolist():
            seen.append(f)
            l2, a = O.tensorbuffer_select(loc, act, f)
            assert np.array_equal(l2.numpy(), g[f"feat{f}_locations"])
            np.testing.assert_array_equal(a.numpy(), g[f"feat{f}_activations"])
            if a.numel() == 0:
                continue
            tw, aw, ids, pooled = O.pool_max_activation_windows(l2, a, big, ctx_len=4, max_examples=5)
            assert np.array_equal(tw.numpy(), g[f"feat{f}_ex_tokens"])
            np.testing.assert_allclose(aw.numpy(), g[f"feat{f}_ex_acts"], rtol=1e-6)
    assert seen == g["loader_features"].tolist()


def test_steering_hook_matches_reference():
    """SteeringController.clamp_features_max hook (features/steering.py:105-124)."""
    g = np.load(os.path.join(GOLDEN, "steering.npz"))
    p = _params(g)
    for tag in ("prefill", "step"):
        h = torch.from_numpy(g[f"{tag}_in"])
        out = O.steering_hook(p, h, int(g["feature"]), float(g["clamp"]))
        assert out.dtype == torch.float16
        np.testing.assert_allclose(out.float().numpy(), g[f"{tag}_out"].astype(np.float32), rtol=2e-3, atol=2e-3)


def test_scan_top_windows_agrees_with_constructor_semantics():
    """The flat-stream scan restatement equals per-feature pool_max_activation_windows on the same data."""
    torch.manual_seed(5)
    p = O.init_params(32, 128, 4, seed=31)
    x = torch.randn(64, 32)
    enc = O.encode(p, x)
    ctx, n_top = 8, 3
    scores, wins = O.scan_top_windows(enc.top_acts, enc.top_indices, 128, ctx, n_top)
    dense = O.topk_masked_latents(p, x[None])[0]  # [T, N]
    tokens = torch.zeros(1, 64, dtype=torch.long)
    for f in range(128):
        col = dense[:, f]
        if not (col.abs() > 1e-5).any():
            assert (wins[f] == -1).all()
            continue
        pos = torch.nonzero(col.abs() > 1e-5)[:, 0]
        loc = torch.stack([torch.zeros_like(pos), pos], 1)
        _, _, ids, pooled = O.pool_max_activation_windows(loc, col[pos], tokens, ctx, n_top)
        n = pooled.numel()
        np.testing.assert_allclose(scores[f, :n], pooled.numpy(), rtol=1e-6)
        # ties between equal pooled values may order differently; compare as sets of windows per value
        assert sorted(wins[f, :n].tolist()) == sorted(ids[:, 1].tolist())


@pytest.mark.parametrize("tag", ["image", "text"])
def test_probe_mean_topk_matches_reference(tag):
    """tools/probe_activations.py:109-126 executed on the reference's own Sae (oracle/gen_golden.py::gen_probe)."""
    g = np.load(os.path.join(GOLDEN, "probe.npz"))
    p = O.SaeParams(torch.from_numpy(g["W_enc"]), torch.from_numpy(g["b_enc"]), torch.from_numpy(g["W_dec"]),
                    torch.from_numpy(g["b_dec"]), int(g["k"]))
    idx, acts = O.probe_mean_topk(p, torch.from_numpy(g[f"{tag}_hidden"]), g[f"{tag}_interval"].tolist(),
                                  bool(g[f"{tag}_drop"]))
    assert np.array_equal(idx.numpy(), g[f"{tag}_indices"])
    np.testing.assert_allclose(acts.numpy(), g[f"{tag}_acts"], rtol=1e-6, atol=1e-7)
