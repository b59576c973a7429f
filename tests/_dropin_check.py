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
print("DROPIN_OK", sorted(counts.items()))
