"""ctypes binding of libukm.so (include/ukm.h).  Fails loudly when the CUDA library is
missing -- there is no CPU fallback anywhere in this package."""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# UKM_LIB_VARIANT=measure (tools/exp_*.py only): the -DUKM_MEASURE build with the pipelines' null modes
LIB_PATH = os.path.join(_HERE, "libukm_measure.so" if os.environ.get("UKM_LIB_VARIANT") == "measure" else "libukm.so")

OK, E_ARG, E_CUDA, E_NOMEM, E_CAPACITY, E_NOT_SORTED_UNIQUE, E_ILLEGAL_BASE, E_NO_TAXONOMY, E_PANIC, E_INTERNAL = (
    0, -1, -2, -3, -4, -5, -6, -7, -8, -9)
HOST, HOST_PINNED, DEVICE = 0, 1, 2
FOLD_PLAIN, FOLD_UNIQUE, FOLD_REPEATED_FINAL, FOLD_REPEATED_CHUNK = 0, 1, 2, 3
OP_INTER, OP_DIFF, OP_UNION = 0, 1, 2
F_TAXID, F_MIX_TAXID, F_COMPARE_TAXID, F_CANONICAL, F_HASHED, F_CIRCULAR, F_SCALED, F_VALIDATE, F_SHARD = 1, 2, 4, 8, 16, 32, 64, 128, 256

# every symbol include/ukm.h declares (tests check that the library exports all of them)
SYMBOLS = [
    "ukm_create", "ukm_destroy", "ukm_last_error", "ukm_version", "ukm_set_stream", "ukm_get_stream", "ukm_sync",
    "ukm_alloc_pinned", "ukm_free_pinned", "ukm_alloc_device", "ukm_free_device", "ukm_copy",
    "ukm_launch_count", "ukm_stats_enable", "ukm_stats_reset", "ukm_stats_get",
    "ukm_set_taxonomy", "ukm_lca_batch",
    "ukm_sort_u64", "ukm_sort_pairs", "ukm_sort_codetaxid16",
    "ukm_fold_sorted", "ukm_merge_sorted", "ukm_union", "ukm_inter", "ukm_diff", "ukm_common", "ukm_setops_stream",
    "ukm_count_seq", "ukm_kmers_seq", "ukm_count_minimizer",
    "ukm_partition_sorted", "ukm_check_sorted_unique",
    "ukm_synth_random_keys", "ukm_synth_member_file", "ukm_synth_bases",
]


class Span(C.Structure):
    _fields_ = [("keys", C.c_void_p), ("taxids", C.c_void_p), ("global_taxid", C.c_uint32),
                ("n", C.c_size_t), ("cap", C.c_size_t), ("where", C.c_int), ("sorted", C.c_int)]


class KernelStat(C.Structure):
    _fields_ = [("name", C.c_char * 48), ("launches", C.c_uint64), ("ms", C.c_double), ("algo_bytes", C.c_double)]


class UkmError(RuntimeError):
    def __init__(self, status: int, message: str):
        super().__init__(f"libukm status {status}: {message}")
        self.status = status
        self.message = message


_lib = None


def load():
    """Load libukm.so.  Raises if it has not been built (python -c 'import __graft_entry__ as g; g.build()')."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(f"{LIB_PATH} not found: build the CUDA library first (make -C unikmer_b200/csrc); "
                          "unikmer_b200 has no CPU fallback")
    L = C.CDLL(LIB_PATH)
    vp, sz, i, u, u64 = C.c_void_p, C.c_size_t, C.c_int, C.c_uint, C.c_uint64
    SP = C.POINTER(Span)
    sig = {
        "ukm_create": ([i], vp), "ukm_destroy": ([vp], None), "ukm_last_error": ([vp], C.c_char_p),
        "ukm_version": ([], C.c_char_p), "ukm_set_stream": ([vp, vp], i), "ukm_get_stream": ([vp], vp),
        "ukm_sync": ([vp], i),
        "ukm_alloc_pinned": ([sz], vp), "ukm_free_pinned": ([vp], None),
        "ukm_alloc_device": ([vp, sz], vp), "ukm_free_device": ([vp, vp], i),
        "ukm_copy": ([vp, vp, i, vp, i, sz], i),
        "ukm_launch_count": ([vp], u64), "ukm_stats_enable": ([vp, i], i), "ukm_stats_reset": ([vp], i),
        "ukm_stats_get": ([vp, C.POINTER(KernelStat), i, C.POINTER(i)], i),
        "ukm_set_taxonomy": ([vp, vp, sz, vp, vp, sz], i),
        "ukm_lca_batch": ([vp, vp, vp, sz, vp, i], i),
        "ukm_sort_u64": ([vp, vp, sz, i, i], i), "ukm_sort_pairs": ([vp, vp, vp, sz, i, i], i),
        "ukm_sort_codetaxid16": ([vp, vp, sz, i], i),
        "ukm_fold_sorted": ([vp, i, SP, u, SP], i),
        "ukm_merge_sorted": ([vp, i, SP, i, u, SP], i),
        "ukm_union": ([vp, SP, i, u, SP], i), "ukm_inter": ([vp, SP, i, u, SP], i),
        "ukm_diff": ([vp, SP, i, u, SP], i), "ukm_common": ([vp, SP, i, u, C.c_uint16, SP], i),
        "ukm_setops_stream": ([vp, SP, i, C.POINTER(i), i, u, SP], i),
        "ukm_count_seq": ([vp, vp, vp, sz, i, u, u64, i, SP], i),
        "ukm_kmers_seq": ([vp, vp, vp, sz, i, u, u64, i, SP], i),
        "ukm_count_minimizer": ([vp, vp, vp, sz, i, i, u, u64, i, SP], i),
        "ukm_partition_sorted": ([vp, SP, vp, i, vp], i), "ukm_check_sorted_unique": ([vp, SP], i),
        "ukm_synth_random_keys": ([vp, u64, sz, u64, vp], i),
        "ukm_synth_member_file": ([vp, u64, sz, u64, u64, u64, i, vp, C.POINTER(sz)], i),
        "ukm_synth_bases": ([vp, u64, u64, sz, u64, vp], i),
    }
    for name, (args, res) in sig.items():
        f = getattr(L, name)
        f.argtypes = args
        f.restype = res
    _lib = L
    return L
