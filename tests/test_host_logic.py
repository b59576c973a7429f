"""CPU: host-side mirror of the reference interface (checkpoint format, split files, loader, constructors) against
the golden fixtures produced by the unmodified reference."""
import json
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN


def test_sae_checkpoint_roundtrip_and_attribute_names(tmp_path):
    """cfg.json + sae.safetensors with keys encoder.weight / encoder.bias / W_dec / b_dec (reference sae/sae.py:126-162)."""
    from safetensors.torch import load_file
    from sae_auto_interp.sae import Sae, SaeConfig

    sae = Sae(16, SaeConfig(num_latents=48, k=4))
    with torch.no_grad():
        sae.encoder.bias.normal_()
        sae.b_dec.normal_()
    sae.save_to_disk(tmp_path / "layers.0")
    keys = sorted(load_file(str(tmp_path / "layers.0" / "sae.safetensors")).keys())
    assert keys == ["W_dec", "b_dec", "encoder.bias", "encoder.weight"]
    cfg = json.load(open(tmp_path / "layers.0" / "cfg.json"))
    assert cfg["d_in"] == 16 and cfg["k"] == 4 and cfg["num_latents"] == 48 and "expansion_factor" in cfg
    back = Sae.load_from_disk(tmp_path / "layers.0")
    for a, b in zip(sae.state_dict().values(), back.state_dict().values()):
        assert torch.equal(a, b)
    many = Sae.load_many(str(tmp_path), local=True)
    assert list(many) == ["layers.0"]
    # decoder rows are unit norm at construction (sae/sae.py:62-64)
    np.testing.assert_allclose(torch.norm(Sae(16, SaeConfig(num_latents=48)).W_dec, dim=1).detach().numpy(), 1.0,
                               rtol=1e-5)


class _FakeSae:
    class cfg:
        num_latents = 64
        expansion_factor = 32
    d_in = 32
    num_latents = 64


class _FakeModel(torch.nn.Module):
    def __init__(self):
        super().__init__()
        self.layers = torch.nn.ModuleList([torch.nn.Identity()])


def _cache_with_golden():
    from sae_auto_interp.features import FeatureCache

    g = np.load(os.path.join(GOLDEN, "cache_chain.npz"))
    fc = FeatureCache(_FakeModel(), None, {"layers.0": _FakeSae()}, batch_size=2, shard_size=100)
    fc.cache.feature_locations["layers.0"] = torch.from_numpy(g["nofilter_locations"])
    fc.cache.feature_activations["layers.0"] = torch.from_numpy(g["nofilter_activations"])
    return fc, g


def test_save_splits_and_concat_match_reference_files(tmp_path):
    """Split files are byte-for-byte the reference's tensors, including the dropped last feature id per split."""
    from safetensors.torch import load_file

    fc, g = _cache_with_golden()
    fc.save_splits(4, str(tmp_path), rank=0)
    fc.concate_safetensors(4, str(tmp_path))
    files = sorted(os.listdir(tmp_path / "layers.0"))
    assert files == sorted(g["split_files"].tolist())
    for f in files:
        data = load_file(str(tmp_path / "layers.0" / f))
        assert np.array_equal(data["locations"].numpy(), g[f"split_{f}_locations"])
        assert np.array_equal(data["activations"].numpy(), g[f"split_{f}_activations"])


def test_fix_split_bounds_keeps_every_feature(tmp_path):
    from safetensors.torch import load_file

    fc, g = _cache_with_golden()
    fc.fix_split_bounds = True
    fc.save_splits(4, str(tmp_path), rank=3)
    n = sum(load_file(str(tmp_path / "layers.0" / f))["activations"].numel()
            for f in os.listdir(tmp_path / "layers.0"))
    assert n == g["nofilter_activations"].shape[0]


def test_loader_and_constructor_match_reference(tmp_path):
    from sae_auto_interp.config import FeatureConfig
    from sae_auto_interp.features import FeatureDataset, pool_max_activation_windows
    from sae_auto_interp.features.features import FeatureRecord

    fc, g = _cache_with_golden()
    fc.save_splits(4, str(tmp_path), rank=0)
    fc.concate_safetensors(4, str(tmp_path))
    cfg = FeatureConfig(width=64, example_ctx_len=4, min_examples=0, max_examples=5, n_splits=4)
    sel = torch.tensor([1, 5, 14, 16, 33, 40, 62])
    ds = FeatureDataset(str(tmp_path), cfg, modules=["layers.0"], features={"layers.0": sel})
    tokens = torch.from_numpy(g["tokens"])
    big = torch.zeros(100 + tokens.shape[0], tokens.shape[1], dtype=torch.long)
    big[100:] = tokens
    seen = []
    for buf in ds.buffers:
        buf._load()
        for i in range(len(buf)):
            bo = buf[i]["buffer"]
            f = bo.feature.feature_index
            seen.append(f)
            assert np.array_equal(bo.locations.numpy(), g[f"feat{f}_locations"])
            assert np.array_equal(bo.activations.numpy(), g[f"feat{f}_activations"])
            if bo.activations.numel() == 0:
                continue
            rec = FeatureRecord(bo.feature)
            pool_max_activation_windows(rec, bo, big, cfg)
            np.testing.assert_allclose(torch.stack([e.activations for e in rec.examples]).numpy(),
                                       g[f"feat{f}_ex_acts"], rtol=1e-6)
            assert np.array_equal(torch.stack([e.tokens for e in rec.examples]).numpy(), g[f"feat{f}_ex_tokens"])
    assert seen == g["loader_features"].tolist()
    # the callback protocol of FeatureDataset.load
    recs = ds.load(collate=True, constructor=lambda record, buffer_output: setattr(record, "examples", []))
    assert [r.feature.feature_index for r in recs] == seen and repr(recs[0].feature) == "layers.0_feature1"


def test_filter_bitmap_bits():
    from saeb200.engine import make_filter_bitmap

    bm = make_filter_bitmap(torch.tensor([0, 31, 32, 63, 95]), 96)
    words = bm.to(torch.int64) & 0xFFFFFFFF
    assert words.tolist() == [(1 << 0) | (1 << 31), (1 << 0) | (1 << 31), 1 << 31]


def test_decoder_seam_autograd_plumbing(monkeypatch):
    """`decoder_impl` is a torch.autograd.Function like the reference's TritonDecoder (sae/kernels.py:403-429).  The
    CUDA kernels behind it are exercised by the -m gpu tests; here the two engine calls are replaced by oracle
    stand-ins so that the plumbing is checked on CPU against the reference's own gradients (fixture generated by
    autograd through the reference's eager_decode / Sae.decode): which inputs get gradients, the layout of the gradient
    of the transposed view `W_dec.mT`, and `+ b_dec` in Sae.decode."""
    import sae_oracle as O
    from saeb200 import engine
    from sae_auto_interp.sae import Sae, SaeConfig
    from sae_auto_interp.sae import utils as seam

    calls = []

    def fake_decode(top_indices, top_acts, W_dec, b_dec, *, out_dtype=torch.float32, **kw):
        assert W_dec.is_contiguous() and b_dec is None
        return O.sparse_decode(top_indices, top_acts, W_dec).to(out_dtype)

    def fake_backward(top_indices, top_acts, W_dec, grad_out, *, need_acts=True, need_weight=True):
        calls.append((need_acts, need_weight))
        d_acts, dW = O.decode_backward(top_indices, top_acts, W_dec, grad_out)
        return (d_acts if need_acts else None), (dW if need_weight else None)

    monkeypatch.setattr(engine, "decode", fake_decode)
    monkeypatch.setattr(engine, "decode_backward", fake_backward)
    g = np.load(os.path.join(GOLDEN, "decode_backward.npz"))
    idx, go = torch.from_numpy(g["top_idx"]), torch.from_numpy(g["grad_out"])
    # the bare seam, gradients for both inputs
    vals = torch.from_numpy(g["top_vals"]).requires_grad_(True)
    W = torch.from_numpy(g["W_dec"]).requires_grad_(True)
    out = seam.decoder_impl(idx, vals, W.mT)
    np.testing.assert_allclose(out.detach().numpy(), g["out"], rtol=1e-5, atol=1e-6)
    out.backward(go)
    np.testing.assert_allclose(vals.grad.numpy(), g["d_vals"], rtol=1e-5, atol=1e-6)
    assert W.grad.shape == W.shape
    np.testing.assert_allclose(W.grad.numpy(), g["d_W_dec"], rtol=1e-5, atol=1e-6)
    # through the module (sae/sae.py:187-191): W_dec and b_dec are parameters, the activations a plain tensor
    sae = Sae(g["W_dec"].shape[1], SaeConfig(num_latents=g["W_dec"].shape[0], k=idx.shape[1]))
    with torch.no_grad():
        sae.W_dec.copy_(torch.from_numpy(g["W_dec"]))
        sae.b_dec.copy_(torch.from_numpy(g["b_dec"]))
    out2 = sae.decode(torch.from_numpy(g["top_vals"]), idx)
    assert out2.requires_grad   # what features/patching/attribution.py:160-161 (`tensor.retain_grad()`) relies on
    out2.backward(go)
    np.testing.assert_allclose(sae.W_dec.grad.numpy(), g["sae_d_W_dec"], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(sae.b_dec.grad.numpy(), g["sae_d_b_dec"], rtol=1e-5, atol=1e-6)
    assert calls == [(True, True), (False, True)]
    # no autograd graph under no_grad (cache / steering paths: features/cache.py:175, features/steering.py:85)
    with torch.no_grad():
        assert not sae.decode(torch.from_numpy(g["top_vals"]), idx).requires_grad


def _oracle_backend(monkeypatch):
    """CPU stand-ins for the engine calls behind the mirror `Sae` (the CUDA kernels have their own -m gpu tests): lets
    the host-side glue above them run on CPU."""
    import sae_oracle as O
    from saeb200 import engine
    from sae_auto_interp.sae import Sae
    from sae_auto_interp.sae.sae import EncoderOutput

    def params(sae):
        return O.SaeParams(sae.encoder.weight.detach(), sae.encoder.bias.detach(), sae.W_dec.detach(),
                           sae.b_dec.detach(), sae.cfg.k)

    def pre_acts(self, x):
        return O.pre_acts(params(self), x.detach().float())

    def encode(self, x, *, clamp_feature=-1, clamp_value=0.0, exact_values=None):
        lat = pre_acts(self, x)
        if clamp_feature >= 0:
            lat[..., clamp_feature] = clamp_value
        return EncoderOutput(*O.select_topk(lat, self.cfg.k))

    monkeypatch.setattr(Sae, "pre_acts", pre_acts)
    monkeypatch.setattr(Sae, "encode", encode)
    monkeypatch.setattr(Sae, "select_topk", lambda self, lat: EncoderOutput(*O.select_topk(lat, self.cfg.k)))
    def decode(i, a, W, b, *, out_dtype=torch.float32, **kw):
        y = O.sparse_decode(i, a.float(), W.float())
        return (y if b is None else y + b).to(out_dtype)

    monkeypatch.setattr(engine, "decode", decode)

    def bwd(i, a, W, g, *, need_acts=True, need_weight=True):
        da, dw = O.decode_backward(i, a, W, g)
        return (da if need_acts else None), (dw if need_weight else None)

    monkeypatch.setattr(engine, "decode_backward", bwd)


class _ToyLogitLM(torch.nn.Module):
    """same host model as oracle/gen_golden.py::ToyLogitLM (the fixture stores its random seed, not its weights)"""

    class Layer(torch.nn.Module):
        def __init__(self, d):
            super().__init__()
            self.lin = torch.nn.Linear(d, d)

        def forward(self, h):
            return (self.lin(h), None)

    def __init__(self, vocab, d, seed):
        super().__init__()
        g = torch.Generator().manual_seed(seed)
        self.emb = torch.nn.Embedding(vocab, d)
        self.layers = torch.nn.ModuleList([_ToyLogitLM.Layer(d)])
        self.head = torch.nn.Linear(d, vocab, bias=False)
        with torch.no_grad():
            self.emb.weight.copy_(torch.randn(vocab, d, generator=g))
            self.layers[0].lin.weight.copy_(torch.randn(d, d, generator=g) / d ** 0.5)
            self.layers[0].lin.bias.zero_()
            self.head.weight.copy_(torch.randn(vocab, d, generator=g) / d ** 0.5)

    def forward(self, input_ids):
        h = self.layers[0](self.emb(input_ids))[0]
        return {"logits": self.head(h.float())}


def test_attribution_patching_matches_reference(monkeypatch):
    """features/patching (utils.py:22-80, attribution.py:131-184): clean / corrupted passes with the SAE
    reconstruction spliced in, one latent switched off through the clamp argument of `encode`, gradient of the
    logit difference retained on the reconstruction."""
    from functools import partial

    from sae_auto_interp.features.patching import (attribution_for_feature, get_logit_diff,
                                                   get_model_forward_cache_with_sae)
    from sae_auto_interp.sae import Sae, SaeConfig

    _oracle_backend(monkeypatch)
    g = np.load(os.path.join(GOLDEN, "attribution.npz"))
    N, d = g["W_enc"].shape
    sae = Sae(d, SaeConfig(num_latents=N, k=int(g["k"])))
    with torch.no_grad():
        sae.encoder.weight.copy_(torch.from_numpy(g["W_enc"]))
        sae.encoder.bias.copy_(torch.from_numpy(g["b_enc"]))
        sae.W_dec.copy_(torch.from_numpy(g["W_dec"]))
        sae.b_dec.copy_(torch.from_numpy(g["b_dec"]))
    model = _ToyLogitLM(40, d, seed=42)
    inputs = {"input_ids": torch.from_numpy(g["input_ids"])}
    metric = partial(get_logit_diff, answer_token_indices=torch.from_numpy(g["answers"]))
    sae_dict, m2n = {"layers.0": sae}, {model.layers[0]: "layers.0"}
    with torch.no_grad():
        logits, clean = get_model_forward_cache_with_sae(model, inputs, sae_dict, m2n)
    assert clean["layers.0"].dtype == torch.float16 and clean["layers.0"].shape == (2, 6, d)
    np.testing.assert_allclose(logits.numpy(), g["clean_logits"], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(float(metric(logits)), float(g["logit_diff_clean"]), rtol=1e-5)
    nonzero = 0
    for f in g["features"].tolist():
        att = attribution_for_feature(model, inputs, sae_dict, m2n, metric, f, clean_cache=clean)["layers.0"]
        np.testing.assert_allclose(att.float().numpy(), g[f"att_{f}"], rtol=2e-3, atol=1e-4)   # fp16 products
        nonzero += int(np.abs(g[f"att_{f}"]).max() > 0)
    assert nonzero >= 3
    # several latents at once take the dense route (pre_acts -> mask -> select_topk) like the reference
    feats = g["features"].tolist()[:2]
    lg2, _ = get_model_forward_cache_with_sae(model, inputs, sae_dict, m2n, off_features=feats)
    assert not np.allclose(lg2.detach().numpy(), g["clean_logits"])


def test_image_constructor_matches_reference():
    """pool_max_activations_windows_image (features/constructors.py:88-148: mean over the 576 base image tokens, best
    max_examples + 50, repeated dataset ids dropped) and prepare_image_examples (features/features.py:49-92: mask
    upsampling + compositing), against the reference's examples on the same synthetic dataset."""
    from sae_auto_interp.config import FeatureConfig
    from sae_auto_interp.features import pool_max_activations_windows_image, random_activations_image
    from sae_auto_interp.features.features import Feature, FeatureRecord, ImageExample
    from sae_auto_interp.features.loader import BufferOutput
    from synth_images import synth_image_cache

    g = np.load(os.path.join(GOLDEN, "image_constructor.npz"))
    ds, loc, act = synth_image_cache()
    assert np.array_equal(loc.numpy(), g["locations"]) and np.array_equal(act.numpy(), g["activations"])
    cfg = FeatureConfig(width=64, max_examples=int(g["max_examples"]))
    bo = BufferOutput(Feature("layers.0", 7), loc, act)
    rec = FeatureRecord(bo.feature)
    pool_max_activations_windows_image(rec, bo, ds, cfg, None)
    assert len(rec.examples) == int(g["n_examples"]) == cfg.max_examples
    for i, ex in enumerate(rec.examples):
        assert isinstance(ex, ImageExample) and ex.tokens.shape == (8000,)
        assert np.array_equal(ex.activations.numpy(), g[f"top{i}_acts"])        # same image, same dense row
        assert np.array_equal(np.asarray(ex.image), g[f"top{i}_image"])
        assert np.array_equal(np.asarray(ex.mask), g[f"top{i}_mask"])
        assert np.array_equal(np.asarray(ex.activation_image), g[f"top{i}_shown"])
    # ranked by the mean over the base tokens, no dataset id twice
    means = [float(ex.activations[:576].mean()) for ex in rec.examples]
    assert means == sorted(means, reverse=True)
    torch.manual_seed(5)
    rec2 = FeatureRecord(bo.feature)
    random_activations_image(rec2, bo, ds, cfg, None)
    for i, ex in enumerate(rec2.examples):
        assert np.array_equal(ex.activations.numpy(), g[f"rand{i}_acts"])
        assert np.array_equal(np.asarray(ex.image), g[f"rand{i}_image"])


def test_mode4_hi_lo_arithmetic_model():
    """CPU model of packed mode 4 with the scale conventions of csrc/pack.cu (hi = fp16(W * 2^s), the largest |W| landing
    in [2^13, 2^14); lo = fp16((W * 2^s - hi) * 2^11); activations scaled per row by a power of two into fp16): the
    tensor-core value x . hi plus the correction x . lo * 2^-s * 2^-11 (csrc/refine.cu, refine_body<.., true>) must
    reproduce the fp64 product to fp32 grade, and bf16 activations must survive the row scaling exactly."""
    rng = np.random.default_rng(0)
    d, N = 512, 96
    W = (rng.uniform(-1, 1, (N, d)) / np.sqrt(d)).astype(np.float32)
    W[5] *= 1e-4   # a tiny-norm feature: its residual still has to stay inside the fp16 range
    x = torch.randn(d, generator=torch.Generator().manual_seed(1)).to(torch.bfloat16).float().numpy()
    _, e = np.frexp(np.abs(W).max())
    scale, unscale = np.float32(2.0 ** (14 - e)), np.float32(2.0 ** (e - 14))
    ws = (W * scale).astype(np.float32)
    hi = ws.astype(np.float16)
    lo = ((ws - hi.astype(np.float32)) * np.float32(2048)).astype(np.float16)
    assert np.isfinite(lo.astype(np.float32)).all() and np.abs(lo.astype(np.float32)).max() <= 8192
    _, ex = np.frexp(np.abs(x).max())
    x16 = (x * np.float32(2.0 ** (14 - ex))).astype(np.float16)
    row_scale = np.float32(2.0 ** (ex - 14))
    assert np.array_equal(x16.astype(np.float32) * row_scale, x)
    exact = W.astype(np.float64) @ x.astype(np.float64)
    a = (hi.astype(np.float64) @ x16.astype(np.float64)) * float(row_scale) * float(unscale)
    corrected = a + (lo.astype(np.float64) @ x.astype(np.float64)) * float(unscale) / 2048.0
    ref = np.abs(exact).max()
    assert np.abs(a - exact).max() / ref > 1e-5          # the single fp16 pass alone is not parity grade
    assert np.abs(corrected - exact).max() / ref < 2e-7  # with the residual plane it is
    assert abs(corrected[5] - exact[5]) <= 2e-7 * np.abs(exact[5]) + 1e-12 * ref


@pytest.mark.parametrize("slots", [1, 2])
def test_push_exchange_protocol_simulation(slots):
    """Host simulation of the peer-memory exchange protocol (csrc/exchange.cu + saeb200/dist.py): R ranks as threads,
    symmetric buffers as shared arrays, one monotonically increasing sequence number per (channel, source rank), a rank
    waits only for "everybody has delivered THIS exchange" -- there is no barrier protecting a region from the next
    write.  The claim it checks: because the two exchanges of a chunk alternate, a region is never overwritten while a
    peer still reads it, with one region per channel (sequential schedule) as well as with two (pipelined).  Random
    delays shake the interleavings; a torn or stale slab fails the content check."""
    import random
    import threading
    import time

    R, chunks, width = 4, 40, 8
    region = [[np.zeros((slots, R, width), np.int64) for _ in range(2)] for _ in range(R)]   # [rank][channel]
    flags = [np.zeros((2, R), np.int64) for _ in range(R)]                                     # [rank][channel, source]
    errors = []

    def push_gather(rank, channel, slot, seq, payload, rng):
        for q in range(R):                       # the kernel's peer walk, starting with the own copy
            p = (rank + q) % R
            region[p][channel][slot, rank, :] = payload
            if rng.random() < 0.3:
                time.sleep(rng.random() * 1e-4)
        for p in range(R):                       # publish after all stores (release), then wait for all peers (acquire)
            flags[p][channel, rank] = seq
        deadline = time.time() + 20
        while (flags[rank][channel] < seq).any():
            if time.time() > deadline:
                errors.append(f"rank {rank}: timeout at seq {seq}")
                return None
            time.sleep(0)
        return region[rank][channel][slot]       # a VIEW: consumed after the call, like the kth kernel does

    def rank_main(rank):
        rng = random.Random(rank)
        seq = [0, 0]
        for c in range(chunks):
            for channel in (0, 1):               # exchange 1 (bounds), exchange 2 (exact values)
                seq[channel] += 1
                got = push_gather(rank, channel, c % slots, seq[channel], 1000 * c + 10 * channel + rank, rng)
                if got is None:
                    return
                if rng.random() < 0.5:
                    time.sleep(rng.random() * 2e-4)   # the consumer kernel runs some time after the exchange
                want = np.array([1000 * c + 10 * channel + r for r in range(R)])[:, None].repeat(width, 1)
                if not np.array_equal(got, want):
                    errors.append(f"rank {rank} chunk {c} channel {channel}: stale or torn slab")
                    return

    threads = [threading.Thread(target=rank_main, args=(r,)) for r in range(R)]
    for t in threads:
        t.start()
    for t in threads:
        t.join(60)
    assert not errors, errors[:3]


@pytest.mark.skipif(not os.path.isdir("/root/reference/sae_auto_interp"), reason="needs the reference checkout")
def test_dropin_overlay_on_the_real_reference():
    """saeb200.dropin.install(): the reference's own launchers / loaders pick up the fused classes (also when they are
    imported after the call), `eager_decode` and everything outside the hot path stay the reference's, `uninstall()`
    restores.  Build-container only: it imports the unmodified reference."""
    import subprocess
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "tests", "_dropin_check.py"), root], capture_output=True,
                       text=True, timeout=600)
    assert r.returncode == 0 and "DROPIN_OK" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]


@pytest.mark.skipif(not os.path.isdir("/root/reference/sae_auto_interp"), reason="needs the reference checkout")
def test_reference_launchers_import_against_the_mirror():
    """The "replace the package" route: with the mirror installed under the reference's package name, the reference's
    cache / steering / attribution launchers (loaded from their own files, unmodified) find every name they import."""
    import subprocess
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = f"""
import sys, importlib.util, types
sys.path[:0] = [{os.path.join(root, 'oracle')!r}, {os.path.join(root, 'multimodal-sae_b200')!r}]
import ref_shims; ref_shims.install_shims()
try:
    import loguru
except Exception:
    lg = types.ModuleType("loguru")
    lg.logger = types.SimpleNamespace(info=print, warning=print, error=print)
    sys.modules["loguru"] = lg
import sae_auto_interp
assert sae_auto_interp.__file__.startswith({os.path.join(root, 'multimodal-sae_b200')!r})
for rel in ("launch/cache/cache.py", "launch/cache/cache_image.py", "launch/features/steering.py",
            "launch/features/attribution_patching.py"):
    spec = importlib.util.spec_from_file_location("ref_" + rel.replace("/", "_")[:-3],
                                                  "/root/reference/sae_auto_interp/" + rel)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    assert hasattr(mod, "main")
print("LAUNCHERS_OK")
"""
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=600, cwd="/tmp")
    assert r.returncode == 0 and "LAUNCHERS_OK" in r.stdout, r.stdout[-1500:] + r.stderr[-1500:]


def _oracle_coo_extract(top_acts, top_indices, seq_len, *, row_offset=0, threshold=1e-5, filter_bitmap=None):
    """CPU stand-in for engine.coo_extract: (row, pos, feature) triples in torch.nonzero order"""
    k = top_acts.shape[-1]
    vals, idx = top_acts.reshape(-1, k).float(), top_indices.reshape(-1, k).long()
    keep = vals.abs() > threshold
    if filter_bitmap is not None:
        words = filter_bitmap.to(torch.int64) & 0xFFFFFFFF
        keep &= ((words[idx >> 5] >> (idx & 31)) & 1).bool()
    tok = torch.arange(vals.shape[0])[:, None].expand_as(idx)
    t, f, a = tok[keep], idx[keep], vals[keep]
    order = torch.argsort(t * (1 << 32) + f)
    t, f, a = t[order], f[order], a[order]
    return torch.stack([row_offset + t // seq_len, t % seq_len, f], 1), a


@pytest.mark.parametrize("tag", ["nofilter", "filter"])
def test_feature_cache_run_host_glue_matches_reference(monkeypatch, tag):
    """FeatureCache.run of the mirror (hooks, batching, row offsets, filter bitmap, host accumulation) against the
    reference's cached triples; the two engine calls are oracle stand-ins here, the kernels have the same test with
    -m gpu (tests/test_gpu_parity.py::test_feature_cache_matches_reference_golden)."""
    from saeb200 import engine
    from sae_auto_interp.features import FeatureCache
    from sae_auto_interp.sae import Sae, SaeConfig

    _oracle_backend(monkeypatch)
    monkeypatch.setattr(engine, "coo_extract", _oracle_coo_extract)
    g = np.load(os.path.join(GOLDEN, "cache_chain.npz"))
    N, d = g["W_enc"].shape
    sae = Sae(d, SaeConfig(num_latents=N, k=int(g["k"])))
    with torch.no_grad():
        sae.encoder.weight.copy_(torch.from_numpy(g["W_enc"]))
        sae.encoder.bias.copy_(torch.from_numpy(g["b_enc"]))
        sae.W_dec.copy_(torch.from_numpy(g["W_dec"]))
        sae.b_dec.copy_(torch.from_numpy(g["b_dec"]))

    class ToyLM(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.emb = torch.nn.Embedding.from_pretrained(torch.from_numpy(g["emb"]).clone())
            self.layers = torch.nn.ModuleList([torch.nn.Linear(d, d)])
            with torch.no_grad():
                self.layers[0].weight.copy_(torch.from_numpy(g["layer_w"]))
                self.layers[0].bias.zero_()

        @property
        def device(self):
            return self.emb.weight.device

        def forward(self, input_ids):
            return self.layers[0](self.emb(input_ids))

    tokens = torch.from_numpy(g["tokens"])
    dataset = [{"input_ids": tokens[i]} for i in range(tokens.shape[0])]
    filters = {"layers.0": torch.tensor([1, 5, 15, 16, 31, 40, 63])} if tag == "filter" else None
    fc = FeatureCache(ToyLM(), None, {"layers.0": sae}, batch_size=2, shard_size=100, filters=filters)
    fc.run(16, dataset)
    loc, act = fc.cache.feature_locations["layers.0"], fc.cache.feature_activations["layers.0"]
    assert loc.dtype == torch.int64 and act.dtype == torch.float32
    assert np.array_equal(loc.numpy(), g[f"{tag}_locations"])
    np.testing.assert_allclose(act.numpy(), g[f"{tag}_activations"], rtol=1e-5)


def test_steering_hook_host_glue_matches_reference(monkeypatch):
    """SteeringController.clamp_features_max (features/steering.py:102-128): the hook replaces the layer output by the
    fp16 reconstruction with one latent clamped at prefill and unclamped at single-token steps.  Oracle stand-ins for
    the engine calls; same fixture as the -m gpu test."""
    from sae_auto_interp.features.steering import SteeringController
    from sae_auto_interp.sae import Sae, SaeConfig

    _oracle_backend(monkeypatch)
    g = np.load(os.path.join(GOLDEN, "steering.npz"))
    N, d = g["W_enc"].shape
    sae = Sae(d, SaeConfig(num_latents=N, k=int(g["k"])))
    with torch.no_grad():
        sae.encoder.weight.copy_(torch.from_numpy(g["W_enc"]))
        sae.encoder.bias.copy_(torch.from_numpy(g["b_enc"]))
        sae.W_dec.copy_(torch.from_numpy(g["W_dec"]))
        sae.b_dec.copy_(torch.from_numpy(g["b_dec"]))

    class Layer(torch.nn.Module):
        def forward(self, h):
            return (h, None)

    layer = Layer()
    handles = SteeringController.clamp_features_max(None, sae, int(g["feature"]), layer, k=float(g["clamp"]))
    try:
        for tag in ("prefill", "step"):
            out = layer(torch.from_numpy(g[f"{tag}_in"]))
            assert out[0].dtype == torch.float16 and out[1] is None and out[0].shape == g[f"{tag}_out"].shape
            ref = g[f"{tag}_out"].astype(np.float32)
            assert np.linalg.norm(out[0].float().numpy() - ref) / np.linalg.norm(ref) < 1e-3
    finally:
        for h in handles:
            h.remove()


def test_save_splits_property_vs_oracle(tmp_path):
    """save_splits for random widths / split counts / feature multisets (empty splits, uneven boundaries, one split)
    against the oracle's restatement of the reference's boolean-mask passes (features/cache.py:243-309), in both
    bound modes: the reference's (last feature id of every split dropped) and the fixed one."""
    from hypothesis import given, settings, strategies as st
    from safetensors.torch import load_file
    from sae_auto_interp.features import FeatureCache
    import sae_oracle as O

    counter = [0]

    @settings(max_examples=25, deadline=None)
    @given(width=st.integers(4, 300), n_splits=st.integers(1, 7), nnz=st.integers(0, 200), seed=st.integers(0, 10 ** 6),
           fix=st.booleans())
    def check(width, n_splits, nnz, seed, fix):
        n_splits = min(n_splits, width)
        g = torch.Generator().manual_seed(seed)
        feats = torch.randint(0, width, (nnz,), generator=g)
        loc = torch.stack([torch.randint(0, 50, (nnz,), generator=g), torch.randint(0, 16, (nnz,), generator=g), feats], 1)
        act = torch.rand(nnz, generator=g)

        class _S:
            class cfg:
                num_latents = width
                expansion_factor = 1
            d_in = width
            num_latents = width

        fc = FeatureCache(_FakeModel(), None, {"layers.0": _S()}, batch_size=2, shard_size=0)
        fc.fix_split_bounds = fix
        fc.cache.feature_locations["layers.0"], fc.cache.feature_activations["layers.0"] = loc, act
        out = tmp_path / f"case{counter[0]}"
        counter[0] += 1
        out.mkdir()
        fc.save_splits(n_splits, str(out), rank=0)
        total = 0
        for start, end in O.generate_split_indices(width, n_splits):
            data = load_file(str(out / "layers.0" / f"Rank0_{start}_{end}.safetensors"))
            mask = O.split_mask(feats, start, end + 1 if fix else end)
            assert torch.equal(data["locations"], loc[mask]) and torch.equal(data["activations"], act[mask])
            total += int(mask.sum())
        if fix:
            assert total == nnz

    check()
