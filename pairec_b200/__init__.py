"""pairec_b200 — B200-native recall -> feature -> rank -> sort/DPP hot path behind pairec's plugin interfaces.

The product is pairec_b200/libpairec_gpu.so (hand-written sm_100a CUDA behind the C ABI in include/pairec_gpu.h).
This package only binds that library (binding.py) and mirrors the reference's operator/plugin interface for the
path (plugin.py).  There is no CPU implementation in here: loading fails loudly when the library is missing.
"""
from .binding import Batcher, Engine, Group, DppParams, SsdParams, PrgError, lib_path, load_library  # noqa: F401

__all__ = ["Batcher", "Engine", "Group", "DppParams", "SsdParams", "PrgError", "lib_path", "load_library"]
