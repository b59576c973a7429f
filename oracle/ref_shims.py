"""TEST INFRASTRUCTURE ONLY -- import shims that let the *unmodified* reference
(`/root/reference/sae_auto_interp`) run on CPU in the build container.

The reference needs a few packages that are not installed here (and there is
no network): simple_parsing, natsort, accelerate.utils, torchtyping, blobfile,
orjson.  This module installs minimal in-memory stand-ins, disables the Triton
decoder (`SAE_DISABLE_TRITON=1`, reference sae/utils.py:119-129) and patches
`LlavaNextProcessor.from_pretrained`, which the reference calls at *class
definition time* as a default argument (reference features/cache.py:321-323).

Only `oracle/gen_golden.py` uses this (to generate `tests/golden/*`).  Nothing
in the product, the GPU tests, `smoke()` or `bench.py` imports it:
`/root/reference` does not exist on the GPU box.
"""
from __future__ import annotations

import dataclasses
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("SAE_REFERENCE_ROOT", "/root/reference")


def _install(name: str, module: types.ModuleType) -> None:
    if name not in sys.modules:
        sys.modules[name] = module


def install_shims() -> None:
    os.environ["SAE_DISABLE_TRITON"] = "1"

    # transformers must be imported BEFORE the accelerate stub exists, otherwise
    # its availability probe trips over `accelerate.__spec__ is None`.
    import transformers  # noqa: F401

    # --- simple_parsing -------------------------------------------------
    sp = types.ModuleType("simple_parsing")

    class Serializable:  # minimal: only to_dict is used on the hot path
        def to_dict(self):
            return dataclasses.asdict(self)

    def field(default=dataclasses.MISSING, default_factory=dataclasses.MISSING, **_kw):
        kw = {}
        if default is not dataclasses.MISSING:
            kw["default"] = default
        if default_factory is not dataclasses.MISSING:
            kw["default_factory"] = default_factory
        return dataclasses.field(**kw)

    def list_field(*values, **_kw):
        return dataclasses.field(default_factory=lambda: list(values))

    sp.Serializable = Serializable
    sp.field = field
    sp.list_field = list_field
    sp.parse = lambda *a, **k: (_ for _ in ()).throw(RuntimeError("simple_parsing stub"))
    sp.ArgumentParser = object
    _install("simple_parsing", sp)

    # --- natsort ----------------------------------------------------------
    ns = types.ModuleType("natsort")

    def natsorted(seq, key=None):
        import re

        def nkey(v):
            s = key(v) if key is not None else v
            return [int(t) if t.isdigit() else t for t in re.split(r"(\d+)", str(s))]

        return sorted(seq, key=nkey)

    ns.natsorted = natsorted
    _install("natsort", ns)

    # --- accelerate.utils ---------------------------------------------------
    acc = types.ModuleType("accelerate")
    acc_utils = types.ModuleType("accelerate.utils")
    acc_utils.send_to_device = lambda x, device: x
    acc.utils = acc_utils
    _install("accelerate", acc)
    _install("accelerate.utils", acc_utils)

    # --- torchtyping ----------------------------------------------------------
    tt = types.ModuleType("torchtyping")

    class _TT:
        def __class_getitem__(cls, item):
            return cls

    tt.TensorType = _TT
    _install("torchtyping", tt)

    # --- blobfile / orjson (only used by FeatureRecord.save) ------------------
    bf = types.ModuleType("blobfile")
    bf.BlobFile = open
    _install("blobfile", bf)
    oj = types.ModuleType("orjson")
    import json as _json

    oj.dumps = lambda o: _json.dumps(o).encode()
    oj.loads = _json.loads
    _install("orjson", oj)

    # --- no network at import time --------------------------------------------
    import transformers as _tf

    _tf.LlavaNextProcessor.from_pretrained = classmethod(lambda cls, *a, **k: None)


def import_reference():
    """Return the reference package (`sae_auto_interp`) imported from REFERENCE_ROOT."""
    install_shims()
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    # make sure we do not pick up this repo's mirror of the same package name
    for k in [k for k in sys.modules if k == "sae_auto_interp" or k.startswith("sae_auto_interp.")]:
        del sys.modules[k]
    import importlib

    pkg = importlib.import_module("sae_auto_interp")
    assert os.path.realpath(pkg.__file__).startswith(os.path.realpath(REFERENCE_ROOT)), pkg.__file__
    return pkg
