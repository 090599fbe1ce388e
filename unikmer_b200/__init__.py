"""unikmer_b200 -- B200 (sm_100a) engine for unikmer's k-mer set-operation hot path.

Layout: csrc/ (hand-written CUDA kernels + the C ABI of include/ukm.h -> libukm.so),
_lib.py (ctypes binding), engine.py (host-side mirror of the reference commands' inner
loops), dist.py (key-range sharding across GPUs, one process per GPU).
The package computes only through libukm.so; importing it does not need a GPU, creating
an Engine does.
"""
from . import _lib
from ._lib import (E_ARG, E_CAPACITY, E_CUDA, E_ILLEGAL_BASE, E_INTERNAL, E_NO_TAXONOMY, E_NOMEM,
                   E_NOT_SORTED_UNIQUE, E_PANIC, FOLD_PLAIN, FOLD_REPEATED_CHUNK, FOLD_REPEATED_FINAL,
                   FOLD_UNIQUE, UkmError)
from .engine import Engine, KmerSet

__all__ = ["Engine", "KmerSet", "UkmError", "_lib", "FOLD_PLAIN", "FOLD_UNIQUE", "FOLD_REPEATED_FINAL",
           "FOLD_REPEATED_CHUNK"]
