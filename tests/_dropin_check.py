"""Helper of tests/test_host_logic.py::test_dropin_overlay_on_the_real_reference (runs in its own interpreter because it
imports the reference package under the name the mirror also uses)."""
import os, sys, tempfile
ROOT = sys.argv[1]
sys.path[:0] = [os.path.join(ROOT, "oracle"), os.path.join(ROOT, "multimodal-sae_b200")]
import ref_shims
ref = ref_shims.import_reference()
import sae_auto_interp.utils as ru
import sae_auto_interp.features as rf
import sae_auto_interp.sae as rs
import sae_auto_interp.sae.utils as rsu
import sae_auto_interp.features.patching.utils as rpu
from saeb200 import dropin
ref_Sae, ref_FC, ref_eager = rs.Sae, rf.FeatureCache, rsu.eager_decode
counts = dropin.install()
mir = dropin.load_mirror()
import importlib
m_sae = importlib.import_module("_saeb200_mirror.sae.sae")
m_cache = importlib.import_module("_saeb200_mirror.features.cache")
assert rs.Sae is m_sae.Sae and ru.Sae is m_sae.Sae and rf.FeatureCache is m_cache.FeatureCache
assert rf.FeatureImageCache is m_cache.FeatureImageCache and rf.Attribution.__module__.startswith("_saeb200_mirror")
assert rsu.decoder_impl.__module__.startswith("_saeb200_mirror") and rsu.eager_decode is ref_eager
assert rpu.get_model_forward_cache_with_sae.__module__.startswith("_saeb200_mirror")
# a launcher imported AFTER install sees the fused classes; reference code outside the hot path is untouched
import sae_auto_interp.launch.features.steering as launcher
assert launcher.Sae is m_sae.Sae and launcher.SteeringController.__module__.startswith("_saeb200_mirror")
assert launcher.load_saes is ru.load_saes and ru.load_saes.__module__ == "sae_auto_interp.utils"
# the reference's loader now builds engine-backed SAEs from a reference-format checkpoint
import torch
with tempfile.TemporaryDirectory() as td:
    sae = ref_Sae(16, rs.SaeConfig(num_latents=48, k=4))
    sae.save_to_disk(os.path.join(td, "layers.0"))
    got = ru.load_single_sae(td, "layers.0", device="cpu")
    assert type(got) is m_sae.Sae and torch.equal(got.W_dec, sae.W_dec)
    many = ru.load_saes(td, device="cpu")
    assert type(many["layers.0"]) is m_sae.Sae
assert dropin.install() == {}          # idempotent
dropin.uninstall()
assert rs.Sae is ref_Sae and rf.FeatureCache is ref_FC and launcher.Sae is ref_Sae

# ---- the host-side glue the mirror restates (utils helpers, sae/data.py) against the reference's own functions
import itertools
import json
import numpy as np

mu = importlib.import_module("_saeb200_mirror.utils")
md = importlib.import_module("_saeb200_mirror.sae.data")
import sae_auto_interp.sae.data as rd

pins = [[336, 672], [672, 336], [672, 672], [1008, 336], [336, 1008]]


def call(f, *a):
    try:
        return tuple(f(*a))
    except ZeroDivisionError:   # the reference divides by zero when the grid cell is smaller than the base size
        return "ZeroDivisionError"


for oh, ow in itertools.product([100, 300, 336, 480, 500, 640, 700, 1000, 1333], [100, 300, 336, 480, 640, 777, 1000, 1500]):
    for h, w, ps in [(336, 336, 14), (384, 384, 14), (336, 336, 24), (336, 224, 14), (224, 336, 16)]:
        args = (oh, ow, h, w, pins, ps)
        assert call(ru.get_anyres_unpadded_size, *args) == call(mu.get_anyres_unpadded_size, *args), args
for ids, t in [([1, 2, 9, 5, 6, 7], 9), ([9, 1], 9), ([3, 4, 9], 9)]:
    assert tuple(ru.get_llava_image_pos(ids, t)) == tuple(mu.get_llava_image_pos(ids, t))
with tempfile.TemporaryDirectory() as td:
    json.dump([{"layers.0_feature1": "a", "prompt": "p"}, {"layers.0_feature2": "b", "prompt": "q"}],
              open(os.path.join(td, "x.json"), "w"))
    os.mkdir(os.path.join(td, "sub"))
    assert ru.load_explanation(td) == mu.load_explanation(td) == {"layers.0_feature1": "a", "layers.0_feature2": "b"}
    path = os.path.join(td, "t.bin")
    np.arange(7 * 16, dtype=np.uint16).tofile(path)
    a, b = rd.MemmapDataset(path, 16, max_examples=6), md.MemmapDataset(path, 16, max_examples=6)
    assert len(a) == len(b) == 6 and (a[3]["input_ids"] == b[3]["input_ids"]).all()
    assert (a.select(range(1, 4)).mmap == b.select(range(1, 4)).mmap).all()
    assert all((a.shard(4, s).mmap == b.shard(4, s).mmap).all() for s in range(4))
try:   # GPT-style chunking with a toy word-level tokenizer (needs `tokenizers` + `datasets`)
    from datasets import Dataset
    from tokenizers import Tokenizer, models, pre_tokenizers
    from transformers import PreTrainedTokenizerFast
except Exception:
    Dataset = None
if Dataset is not None:
    words = [f"w{i}" for i in range(50)]
    vocab = {"<eos>": 0, "<unk>": 1, **{w: i + 2 for i, w in enumerate(words)}}
    tk = Tokenizer(models.WordLevel(vocab, unk_token="<unk>"))
    tk.pre_tokenizer = pre_tokenizers.WhitespaceSplit()
    tok = PreTrainedTokenizerFast(tokenizer_object=tk, eos_token="<eos>", unk_token="<unk>", model_max_length=4096)
    rng = np.random.default_rng(0)
    texts = [" ".join(rng.choice(words, size=int(rng.integers(3, 40)))) for _ in range(60)]
    ds = Dataset.from_dict({"text": texts, "meta": list(range(60))})
    for kw in ({}, {"return_final_batch": True}):
        common = dict(format=None, num_proc=1, max_seq_len=32, load_from_cache_file=False, **kw)
        ra, rb = rd.chunk_and_tokenize(ds, tok, **common), md.chunk_and_tokenize(ds, tok, **common)
        A = [list(map(int, r["input_ids"])) for r in ra]
        assert A == [list(map(int, r["input_ids"])) for r in rb] and len(A) > 3
print("DROPIN_OK", sorted(counts.items()))
