"""saeb200: B200-native SAE encode / TopK / sparse-decode / activation-cache engine (ctypes over a C ABI)."""
from . import _capi  # noqa: F401
from ._capi import SaebError  # noqa: F401
