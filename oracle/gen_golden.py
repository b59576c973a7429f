"""TEST INFRASTRUCTURE ONLY -- generate `tests/golden/*.npz` by running the UNMODIFIED reference.

Run in the build container (where /root/reference exists):

    python oracle/gen_golden.py

The fixtures pin `oracle/sae_oracle.py` (tests/test_oracle_golden.py) and are also what the GPU
parity tests compare against at the small sizes.  Every array saved here is an output of reference
code (imported via `oracle/ref_shims.py`), never of the oracle.
"""
from __future__ import annotations

import os
import sys
import tempfile

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, HERE)

import ref_shims  # noqa: E402
import sae_oracle as O  # noqa: E402  (only for the synthetic-parameter generator)
from synth_images import synth_image_cache  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")


def build_ref_sae(ref_sae_mod, p: O.SaeParams):
    Sae, SaeConfig = ref_sae_mod.Sae, ref_sae_mod.SaeConfig
    sae = Sae(p.d_in, SaeConfig(num_latents=p.num_latents, k=p.k), device="cpu")
    with torch.no_grad():
        sae.encoder.weight.copy_(p.W_enc)
        sae.encoder.bias.copy_(p.b_enc)
        sae.W_dec.copy_(p.W_dec)
        sae.b_dec.copy_(p.b_dec)
    return sae


def canon(acts, idx):
    i, v = O.canonical_topk(acts, idx)
    return i, v


def gen_forward(ref, name, d, N, k, T, seed, bf16_x):
    """BASELINE.json configs[0] (C1) and a second small shape: reference Sae.forward."""
    p = O.init_params(d, N, k, seed)
    sae = build_ref_sae(ref.sae, p)
    g = torch.Generator().manual_seed(seed + 1)
    x = torch.randn(T, d, generator=g)
    if bf16_x:
        x = x.to(torch.bfloat16).to(torch.float32)
    with torch.no_grad():
        out = sae(x)
        pa = sae.pre_acts(x)
    idx, val = canon(out.latent_acts, out.latent_indices)
    np.savez_compressed(
        os.path.join(GOLD, name),
        W_enc=p.W_enc.numpy(), b_enc=p.b_enc.numpy(), W_dec=p.W_dec.numpy(), b_dec=p.b_dec.numpy(),
        x=x.numpy(), k=np.int64(k),
        top_idx=idx, top_val=val, sae_out=out.sae_out.numpy(), fvu=out.fvu.numpy(),
        pre_acts_sum=pa.double().sum(-1).numpy(), pre_acts_nnz=(pa > 0).sum(-1).numpy(),
    )
    print(name, "fvu", float(out.fvu))


def gen_decode_test(ref):
    """The reference's only test (train/sae/tests/test_decode.py:6-20) at its own shapes, on CPU:
    eager_decode(top_idx, top_vals, W_dec.mT)."""
    from sae_auto_interp.sae.utils import eager_decode

    g = torch.Generator().manual_seed(7)
    latents = torch.rand(2, 100, generator=g)
    W_dec = torch.randn(100, 50, generator=g)
    top_vals, top_idx = latents.topk(10)
    res = eager_decode(top_idx, top_vals, W_dec.mT)
    np.savez_compressed(os.path.join(GOLD, "decode_test.npz"), latents=latents.numpy(), W_dec=W_dec.numpy(),
                        top_vals=top_vals.numpy(), top_idx=top_idx.numpy(), eager=res.numpy())


def gen_decode_backward(ref):
    """Gradients of the decoder seam (`decoder_impl(top_indices, top_acts, W_dec.mT)`, sae/utils.py:107-129; autograd
    contract of TritonDecoder, sae/kernels.py:403-429) from the reference's own CPU path: autograd through
    `eager_decode`, and through `Sae.decode` (adds b_dec).  One activation is exactly zero: its gradient is still the
    gathered dot product."""
    from sae_auto_interp.sae.utils import eager_decode

    d, N, k, T = 48, 160, 6, 20
    p = O.init_params(d, N, k, seed=31)
    g = torch.Generator().manual_seed(32)
    x = torch.randn(T, d, generator=g)
    sae = build_ref_sae(ref.sae, p)
    with torch.no_grad():
        enc = sae.encode(x)
    top_idx = enc.top_indices.clone()
    top_vals = enc.top_acts.clone()
    top_vals[3, 2] = 0.0
    grad_out = torch.randn(T, d, generator=g)
    vals_a = top_vals.clone().requires_grad_(True)
    W = p.W_dec.clone().requires_grad_(True)
    out = eager_decode(top_idx, vals_a, W.mT)
    out.backward(grad_out)
    # through the module: Sae.decode = decoder_impl(...) + b_dec (sae/sae.py:187-191)
    sae.zero_grad()
    vals_b = top_vals.clone().requires_grad_(True)
    out2 = sae.decode(vals_b, top_idx)
    out2.backward(grad_out)
    np.savez_compressed(os.path.join(GOLD, "decode_backward.npz"), W_dec=p.W_dec.numpy(), b_dec=p.b_dec.numpy(),
                        top_idx=top_idx.numpy(), top_vals=top_vals.numpy(), grad_out=grad_out.numpy(),
                        out=out.detach().numpy(), d_vals=vals_a.grad.numpy(), d_W_dec=W.grad.numpy(),
                        sae_d_vals=vals_b.grad.numpy(), sae_d_W_dec=sae.W_dec.grad.numpy(),
                        sae_d_b_dec=sae.b_dec.grad.numpy())
    print("decode_backward |dW|", float(W.grad.abs().sum()))


def gen_image_constructor(ref):
    """pool_max_activations_windows_image / random_activations_image (features/constructors.py:88-181) +
    prepare_image_examples (features/features.py:49-92) on a synthetic image dataset with repeated ids."""
    from sae_auto_interp.config import FeatureConfig
    from sae_auto_interp.features import pool_max_activations_windows_image, random_activations_image
    from sae_auto_interp.features.features import Feature, FeatureRecord
    from sae_auto_interp.features.loader import BufferOutput

    ds, loc, act = synth_image_cache()
    cfg = FeatureConfig(width=64, max_examples=6)
    bo = BufferOutput(Feature("layers.0", 7), loc, act)
    rec = FeatureRecord(bo.feature)
    pool_max_activations_windows_image(rec, bo, ds, cfg, None)
    out = {"n_examples": np.int64(len(rec.examples))}
    for i, ex in enumerate(rec.examples):
        out[f"top{i}_acts"] = ex.activations.numpy()
        out[f"top{i}_mask"] = np.asarray(ex.mask)
        out[f"top{i}_shown"] = np.asarray(ex.activation_image)
        out[f"top{i}_image"] = np.asarray(ex.image)
    torch.manual_seed(5)
    rec2 = FeatureRecord(bo.feature)
    random_activations_image(rec2, bo, ds, cfg, None)
    for i, ex in enumerate(rec2.examples):
        out[f"rand{i}_acts"] = ex.activations.numpy()
        out[f"rand{i}_image"] = np.asarray(ex.image)
    np.savez_compressed(os.path.join(GOLD, "image_constructor.npz"), locations=loc.numpy(), activations=act.numpy(),
                        max_examples=np.int64(cfg.max_examples), **out)
    print("image_constructor examples", len(rec.examples), len(rec2.examples))


class ToyLogitLM(torch.nn.Module):
    """Host model for the patching fixtures: embedding -> hooked layer (returns a tuple, like a decoder layer: the
    reference hook only handles tuple outputs) -> vocabulary head; `model(**inputs)["logits"]`."""

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
        self.layers = torch.nn.ModuleList([ToyLogitLM.Layer(d)])
        self.head = torch.nn.Linear(d, vocab, bias=False)
        with torch.no_grad():
            self.emb.weight.copy_(torch.randn(vocab, d, generator=g))
            self.layers[0].lin.weight.copy_(torch.randn(d, d, generator=g) / d ** 0.5)
            self.layers[0].lin.bias.zero_()
            self.head.weight.copy_(torch.randn(vocab, d, generator=g) / d ** 0.5)

    def forward(self, input_ids):
        h = self.layers[0](self.emb(input_ids))[0]
        return {"logits": self.head(h.float())}


def gen_attribution(ref):
    """Attribution patching: reference get_model_forward_cache_with_sae / get_model_backward_cache_with_sae
    (features/patching/utils.py:22-80) and the attribution formula of Attribution.get_attribution
    (features/patching/attribution.py:131-184, restated here as glue around the reference functions because the class
    constructor needs a tokenizer, an image processor and image files)."""
    from functools import partial

    from sae_auto_interp.features.patching.utils import (get_logit_diff, get_model_backward_cache_with_sae,
                                                         get_model_forward_cache_with_sae)

    d, N, k, vocab = 32, 64, 4, 40
    p = O.init_params(d, N, k, seed=41)
    sae = build_ref_sae(ref.sae, p)
    model = ToyLogitLM(vocab, d, seed=42)
    g = torch.Generator().manual_seed(43)
    input_ids = torch.randint(0, vocab, (2, 6), generator=g)
    answers = torch.tensor([[3, 7], [11, 2]])
    metric = partial(get_logit_diff, answer_token_indices=answers)
    sae_dict = {"layers.0": sae}
    module_to_name = {model.layers[0]: "layers.0"}
    # latents that are active somewhere in the clean pass, plus one that never fires
    with torch.no_grad():
        h = model.layers[0](model.emb(input_ids))[0]
        enc = sae.encode(h.flatten(0, 1))
    active = torch.unique(enc.top_indices)
    never = [i for i in range(N) if i not in set(active.tolist())][:1]
    # the metric reads the last position only: latents firing there give non-zero attribution, the others exactly 0
    last = torch.unique(enc.top_indices.view(2, 6, k)[:, -1, :]).tolist()
    elsewhere = [i for i in active.tolist() if i not in last][:1]
    feats = last[:5] + elsewhere + never
    out = {}
    for f in feats:
        model.zero_grad(), sae.zero_grad()
        clean_logits, clean = get_model_forward_cache_with_sae(model, {"input_ids": input_ids}, sae_dict,
                                                               module_to_name)
        cor_logits, cor = get_model_forward_cache_with_sae(model, {"input_ids": input_ids}, sae_dict, module_to_name,
                                                           off_features=f)
        for t in cor.values():
            t.retain_grad()
        val = get_model_backward_cache_with_sae(logits=cor_logits, metrics=metric)
        att = ((clean["layers.0"] - cor["layers.0"]) * cor["layers.0"].grad).detach().sum(dim=-1)
        out[f"att_{f}"] = att.float().numpy()
        out[f"metric_{f}"] = np.float32(val.item())
        out[f"logits_{f}"] = cor_logits.detach().numpy()
    np.savez_compressed(os.path.join(GOLD, "attribution.npz"), W_enc=p.W_enc.numpy(), b_enc=p.b_enc.numpy(),
                        W_dec=p.W_dec.numpy(), b_dec=p.b_dec.numpy(), k=np.int64(k), input_ids=input_ids.numpy(),
                        answers=answers.numpy(), features=np.array(feats), clean_logits=clean_logits.detach().numpy(),
                        clean_rec=clean["layers.0"].detach().float().numpy(),
                        logit_diff_clean=np.float32(metric(clean_logits).item()), **out)
    print("attribution features", feats, "max |att|", max(float(np.abs(out[f"att_{f}"]).max()) for f in feats))


class ToyLM(torch.nn.Module):
    """Stand-in host model: embedding -> one 'layer' whose output is hooked (features/cache.py:178-191)."""

    def __init__(self, vocab, d, seed):
        super().__init__()
        g = torch.Generator().manual_seed(seed)
        self.emb = torch.nn.Embedding(vocab, d)
        self.layers = torch.nn.ModuleList([torch.nn.Linear(d, d)])
        with torch.no_grad():
            self.emb.weight.copy_(torch.randn(vocab, d, generator=g))
            self.layers[0].weight.copy_(torch.randn(d, d, generator=g) / d ** 0.5)
            self.layers[0].bias.zero_()

    @property
    def device(self):
        return self.emb.weight.device

    def forward(self, input_ids):
        return self.layers[0](self.emb(input_ids))


def gen_cache_chain(ref):
    """FeatureCache.run -> save_splits -> concate_safetensors -> FeatureDataset/TensorBuffer ->
    pool_max_activation_windows, all reference code, toy sizes."""
    from safetensors.torch import load_file
    from sae_auto_interp.config import FeatureConfig
    from sae_auto_interp.features import FeatureCache, FeatureDataset, pool_max_activation_windows
    from sae_auto_interp.features.features import FeatureRecord

    d, N, k, vocab, seq, bs, n_rows = 32, 64, 4, 50, 16, 2, 8
    p = O.init_params(d, N, k, seed=11)
    sae = build_ref_sae(ref.sae, p)
    model = ToyLM(vocab, d, seed=12)
    g = torch.Generator().manual_seed(13)
    tokens = torch.randint(0, vocab, (n_rows, seq), generator=g)
    dataset = [{"input_ids": tokens[i]} for i in range(n_rows)]
    out = {}
    for tag, filters in (("nofilter", None), ("filter", {"layers.0": torch.tensor([1, 5, 15, 16, 31, 40, 63])})):
        cache = FeatureCache(model, None, {"layers.0": sae}, batch_size=bs, shard_size=100, filters=filters)
        cache.run(seq, dataset)
        loc = cache.cache.feature_locations["layers.0"]
        act = cache.cache.feature_activations["layers.0"]
        out[f"{tag}_locations"] = loc.numpy()
        out[f"{tag}_activations"] = act.numpy()
        if tag == "nofilter":
            with tempfile.TemporaryDirectory() as td:
                n_splits = 4
                cache.save_splits(n_splits, td, rank=0)
                cache.concate_safetensors(n_splits, td)
                files = sorted(os.listdir(os.path.join(td, "layers.0")))
                out["split_files"] = np.array(files)
                for f in files:
                    data = load_file(os.path.join(td, "layers.0", f))
                    out[f"split_{f}_locations"] = data["locations"].numpy()
                    out[f"split_{f}_activations"] = data["activations"].numpy()
                # loader + constructor on a few features (shard_size offset 100 -> rows 100..107)
                cfg = FeatureConfig(width=N, example_ctx_len=4, min_examples=0, max_examples=5, n_splits=n_splits)
                sel = torch.tensor([1, 5, 14, 16, 33, 40, 62])
                ds = FeatureDataset(td, cfg, modules=["layers.0"], features={"layers.0": sel})
                big_tokens = torch.zeros(100 + n_rows, seq, dtype=torch.long)
                big_tokens[100:] = tokens
                feats = []
                for buf in ds.buffers:
                    buf._load()
                    for i in range(len(buf)):
                        bo = buf[i]["buffer"]
                        f = bo.feature.feature_index
                        feats.append(f)
                        out[f"feat{f}_locations"] = bo.locations.numpy()
                        out[f"feat{f}_activations"] = bo.activations.numpy()
                        if bo.activations.numel() == 0:
                            continue
                        rec = FeatureRecord(bo.feature)
                        pool_max_activation_windows(rec, bo, big_tokens, cfg)
                        out[f"feat{f}_ex_tokens"] = torch.stack([e.tokens for e in rec.examples]).numpy()
                        out[f"feat{f}_ex_acts"] = torch.stack([e.activations for e in rec.examples]).numpy()
                out["loader_features"] = np.array(feats)
    np.savez_compressed(os.path.join(GOLD, "cache_chain.npz"), tokens=tokens.numpy(),
                        W_enc=p.W_enc.numpy(), b_enc=p.b_enc.numpy(), W_dec=p.W_dec.numpy(),
                        b_dec=p.b_dec.numpy(), k=np.int64(k), emb=model.emb.weight.detach().numpy(),
                        layer_w=model.layers[0].weight.detach().numpy(), **out)
    # hidden states the hook saw (what the GPU parity test feeds to the engine)
    with torch.no_grad():
        hidden = model(tokens)
    np.save(os.path.join(GOLD, "cache_chain_hidden.npy"), hidden.numpy())
    print("cache_chain nnz", out["nofilter_locations"].shape, out["filter_locations"].shape)


def gen_steering(ref):
    """SteeringController.clamp_features_max hook body (features/steering.py:102-128), prefill (T>1) and
    decode step (T==1)."""
    from sae_auto_interp.features.steering import SteeringController

    d, N, k = 64, 256, 8
    p = O.init_params(d, N, k, seed=21)
    sae = build_ref_sae(ref.sae, p)

    class Layer(torch.nn.Module):
        def forward(self, h):
            return (h, None)

    layer = Layer()
    handles = SteeringController.clamp_features_max(None, sae, 37, layer, k=50.0)
    g = torch.Generator().manual_seed(22)
    res = {}
    with torch.no_grad():
        for tag, T in (("prefill", 24), ("step", 1)):
            h = torch.randn(1, T, d, generator=g).to(torch.float16)
            out = layer(h)
            res[f"{tag}_in"] = h.numpy()
            res[f"{tag}_out"] = out[0].numpy()
    for hnd in handles:
        hnd.remove()
    np.savez_compressed(os.path.join(GOLD, "steering.npz"), W_enc=p.W_enc.numpy(), b_enc=p.b_enc.numpy(),
                        W_dec=p.W_dec.numpy(), b_dec=p.b_dec.numpy(), k=np.int64(k), feature=np.int64(37),
                        clamp=np.float32(50.0), **res)


def gen_probe(ref):
    """tools/probe_activations.py:109-126 -- the hook body is inline in the script's __main__ (no importable
    function), so the fixture executes exactly those statements on the reference's own `Sae` object."""
    d, N, k = 64, 512, 8
    p = O.init_params(d, N, k, seed=61)
    sae = build_ref_sae(ref.sae, p)
    g = torch.Generator().manual_seed(62)
    res = {}
    with torch.no_grad():
        for tag, T, interval, drop in (("image", 41, [0, 12], True), ("text", 23, [3, 10], False)):
            hidden = torch.randn(1, T, d, generator=g).to(torch.float16)
            latents = sae.pre_acts(hidden)                                   # :112
            if drop:
                latents = latents[:, 1:, :]                                  # :115-116
            topk_indices = (latents.squeeze(0).mean(dim=0).topk(k=interval[1]).indices.detach().cpu())[interval[0]:]
            topk_acts = latents[:, :, topk_indices].squeeze(0).permute(1, 0).detach().cpu()   # :122
            res[f"{tag}_hidden"] = hidden.numpy()
            res[f"{tag}_interval"] = np.asarray(interval, np.int64)
            res[f"{tag}_drop"] = np.int64(drop)
            res[f"{tag}_indices"] = topk_indices.numpy()
            res[f"{tag}_acts"] = topk_acts.numpy()
            res[f"{tag}_mean"] = latents.squeeze(0).mean(dim=0).numpy()
    np.savez_compressed(os.path.join(GOLD, "probe.npz"), W_enc=p.W_enc.numpy(), b_enc=p.b_enc.numpy(),
                        W_dec=p.W_dec.numpy(), b_dec=p.b_dec.numpy(), k=np.int64(k), **res)


def main():
    os.makedirs(GOLD, exist_ok=True)
    torch.set_num_threads(8)
    ref = ref_shims.import_reference()
    import sae_auto_interp.sae as ref_sae

    ref.sae = ref_sae
    only = sys.argv[sys.argv.index("--only") + 1] if "--only" in sys.argv else None
    if only is not None:   # regenerate one fixture without touching the others
        {"decode_backward": gen_decode_backward, "attribution": gen_attribution, "image_constructor": gen_image_constructor, "decode_test": gen_decode_test, "cache_chain": gen_cache_chain,
         "steering": gen_steering, "probe": gen_probe}[only](ref)
        return
    gen_decode_backward(ref)
    gen_attribution(ref)
    gen_image_constructor(ref)
    gen_forward(ref, "forward_c1.npz", d=128, N=512, k=16, T=256, seed=1234, bf16_x=False)
    gen_forward(ref, "forward_c1_bf16.npz", d=128, N=512, k=16, T=256, seed=1235, bf16_x=True)
    gen_forward(ref, "forward_wide.npz", d=64, N=2048, k=32, T=96, seed=1236, bf16_x=True)
    gen_decode_test(ref)
    gen_cache_chain(ref)
    gen_steering(ref)
    gen_probe(ref)
    print("golden fixtures written to", GOLD)


if __name__ == "__main__":
    main()
