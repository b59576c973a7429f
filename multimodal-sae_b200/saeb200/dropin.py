"""Plug the B200 engine into an INSTALLED reference package, without editing or shadowing it.

    import saeb200.dropin; saeb200.dropin.install()          # then run the reference's launchers as usual
    python -m saeb200.dropin -m sae_auto_interp.launch.cache.cache_image <launcher args>

The reference has no plugin registry; its seams are module-level names (`decoder_impl`, sae/utils.py:119-129) and the
classes its launchers import (`Sae`, `FeatureCache`, `FeatureImageCache`, `SteeringController`, `Attribution`).
`install()` imports the reference, loads this repository's mirror of those modules under a private package name and
rebinds every reference module attribute that IS one of the replaced objects to the mirror's object, so
`from sae_auto_interp.features import FeatureCache` -- executed before or after the call -- yields the fused
implementation while everything outside the hot path (launchers, explainers, clients, loaders, configs) stays the
reference's own code.  The alternative, putting the mirror package ahead of the reference on PYTHONPATH, only works for
callers that need nothing but the mirrored modules.
"""
from __future__ import annotations

import importlib
import importlib.util
import os
import runpy
import sys
from typing import Dict, List, Tuple

PKG_DIR = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
MIRROR_DIR = os.path.join(PKG_DIR, "sae_auto_interp")
MIRROR_NAME = "_saeb200_mirror"

# (reference submodule, attribute names rebound to the mirror's objects of the same submodule)
REPLACED: List[Tuple[str, Tuple[str, ...]]] = [
    ("sae.sae", ("Sae", "EncoderOutput", "ForwardOutput")),
    ("sae.utils", ("decoder_impl",)),
    ("features.cache", ("Cache", "FeatureCache", "FeatureImageCache")),
    ("features.steering", ("SteeringController",)),
    ("features.patching.utils", ("get_logit_diff", "get_model_forward_cache_with_sae",
                                 "get_model_backward_cache_with_sae")),
    ("features.patching.attribution", ("Attribution",)),
]

_undo: List[Tuple[object, str, object]] = []
_installed: Dict[str, Dict[int, object]] = {}   # reference package name -> {id(mirror object): reference object}


def load_mirror():
    """this repository's `sae_auto_interp` mirror, imported under a private name so that it can coexist with the
    reference package of the same name (its modules only use relative imports)"""
    mod = sys.modules.get(MIRROR_NAME)
    if mod is None:
        spec = importlib.util.spec_from_file_location(MIRROR_NAME, os.path.join(MIRROR_DIR, "__init__.py"),
                                                      submodule_search_locations=[MIRROR_DIR])
        mod = importlib.util.module_from_spec(spec)
        sys.modules[MIRROR_NAME] = mod
        spec.loader.exec_module(mod)
    return mod


def install(reference: str = "sae_auto_interp") -> Dict[str, int]:
    """Rebind the reference's hot-path names to the B200 implementations.  Returns {replaced name: number of module
    attributes rebound}.  Idempotent; `uninstall()` restores the reference."""
    ref = importlib.import_module(reference)
    ref_dir = os.path.realpath(os.path.dirname(ref.__file__))
    if ref_dir == os.path.realpath(MIRROR_DIR):
        return {}   # the mirror itself is installed under the reference's name: nothing to overlay
    if _undo:
        return {}
    load_mirror()
    swaps: Dict[int, Tuple[str, object]] = {}
    for sub, names in REPLACED:
        try:
            ref_mod = importlib.import_module(f"{reference}.{sub}")
        except ImportError:
            continue   # an older reference without this module
        mir_mod = importlib.import_module(f"{MIRROR_NAME}.{sub}")
        for name in names:
            if hasattr(ref_mod, name):
                swaps[id(getattr(ref_mod, name))] = (name, getattr(mir_mod, name))
                _installed.setdefault(reference, {})[id(getattr(mir_mod, name))] = getattr(ref_mod, name)
    counts: Dict[str, int] = {}
    for mod_name, mod in list(sys.modules.items()):
        if mod is None or not (mod_name == reference or mod_name.startswith(reference + ".")):
            continue
        for attr, value in list(vars(mod).items()):
            hit = swaps.get(id(value))
            # classes are rebound under any alias; functions only under their seam name (`decoder_impl` IS
            # `eager_decode` or `triton_decode` in the reference -- those two names keep pointing at the originals)
            if hit is not None and (isinstance(value, type) or attr == hit[0]):
                _undo.append((mod, attr, value))
                setattr(mod, attr, hit[1])
                counts[hit[0]] = counts.get(hit[0], 0) + 1
    return counts


def uninstall() -> None:
    """Restore the reference's own objects, also in reference modules that were imported after `install()` (they
    copied the rebound names)."""
    while _undo:
        mod, attr, value = _undo.pop()
        setattr(mod, attr, value)
    for reference, back in _installed.items():
        for mod_name, mod in list(sys.modules.items()):
            if mod is None or not (mod_name == reference or mod_name.startswith(reference + ".")):
                continue
            for attr, value in list(vars(mod).items()):
                if id(value) in back:
                    setattr(mod, attr, back[id(value)])
    _installed.clear()


def main(argv=None) -> None:
    argv = list(sys.argv[1:] if argv is None else argv)
    if len(argv) < 2 or argv[0] != "-m":
        raise SystemExit("usage: python -m saeb200.dropin -m <reference launcher module> [launcher args]")
    install()
    sys.argv = [argv[1]] + argv[2:]
    runpy.run_module(argv[1], run_name="__main__", alter_sys=True)


if __name__ == "__main__":
    main()
