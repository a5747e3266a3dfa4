"""mvin_b200 -- B200-native (sm_100a) implementation of the MVIN hot path behind the reference's `MVIN` class API.

Only the path is here: csrc/ (CUDA kernels + C ABI, built into lib/libmvin_b200.so), _lib.py (ctypes binding),
model.py (the reference's Python model interface).  See DESIGN.md / INTEGRATION.md.
"""
from .model import MVIN  # noqa: F401

__all__ = ["MVIN"]
