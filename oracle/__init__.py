"""ctypes front-end of the CPU oracle (oracle.c).  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this package (see oracle.c header).  The product package
(unikmer_b200) never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from typing import Optional, Sequence

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "liboracle.so")

FOLD_PLAIN, FOLD_UNIQUE, FOLD_REPEATED_FINAL, FOLD_REPEATED_CHUNK = 0, 1, 2, 3
E_ILLEGAL_BASE, E_ARG, E_NOMEM, E_PANIC = -1, -2, -3, -4


class OracleError(RuntimeError):
    def __init__(self, code: int):
        super().__init__({-1: "illegal base", -2: "bad argument", -3: "out of memory",
                          -4: "reference would panic here"}.get(code, f"error {code}"))
        self.code = code


def build(force: bool = False) -> str:
    """Compile oracle.c -> liboracle.so (recipe: oracle/Makefile)."""
    src = os.path.join(_HERE, "oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"] + (["-B"] if force else []))
    return _SO


class _File(C.Structure):
    _fields_ = [("keys", C.c_void_p), ("taxids", C.c_void_p), ("n", C.c_size_t), ("sorted", C.c_int)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        L = C.CDLL(_SO)
        u64p, u32p, u8p = C.c_void_p, C.c_void_p, C.c_void_p
        L.orc_encode.argtypes = [u8p, C.c_int, C.POINTER(C.c_uint64)]
        L.orc_encode.restype = C.c_int
        L.orc_revcomp.argtypes = [C.c_uint64, C.c_int]
        L.orc_revcomp.restype = C.c_uint64
        L.orc_canonical.argtypes = [C.c_uint64, C.c_int]
        L.orc_canonical.restype = C.c_uint64
        L.orc_decode.argtypes = [C.c_uint64, C.c_int, C.c_char_p]
        L.orc_decode.restype = None
        for f in (L.orc_kmer_iter, L.orc_nthash_iter):
            f.argtypes = [u8p, C.c_int64, C.c_int, C.c_int, C.c_int, u64p]
            f.restype = C.c_int64
        L.orc_tax_new.argtypes = [u32p, C.c_size_t, u32p, u32p, C.c_size_t]
        L.orc_tax_new.restype = C.c_void_p
        L.orc_tax_free.argtypes = [C.c_void_p]
        L.orc_tax_free.restype = None
        L.orc_lca.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32]
        L.orc_lca.restype = C.c_uint32
        L.orc_sort_u64.argtypes = [u64p, C.c_size_t, C.c_int]
        L.orc_sort_u64.restype = C.c_int
        L.orc_sort_pairs.argtypes = [u64p, u32p, C.c_size_t]
        L.orc_sort_pairs.restype = C.c_int
        L.orc_fold.argtypes = [C.c_int, u64p, u32p, C.c_size_t, C.c_void_p, u64p, u32p]
        L.orc_fold.restype = C.c_int64
        L.orc_merge_chunks.argtypes = [C.POINTER(_File), C.c_int, C.c_int, C.c_int, C.c_void_p, u64p, u32p]
        L.orc_merge_chunks.restype = C.c_int64
        L.orc_union.argtypes = [C.POINTER(_File), C.c_int, C.c_int, C.c_void_p, C.c_int, u64p, u32p]
        L.orc_union.restype = C.c_int64
        L.orc_inter.argtypes = [C.POINTER(_File), C.c_int, C.c_int, C.c_int, C.c_void_p, u64p, u32p]
        L.orc_inter.restype = C.c_int64
        L.orc_diff.argtypes = [C.POINTER(_File), C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, u64p, u32p]
        L.orc_diff.restype = C.c_int64
        L.orc_common.argtypes = [C.POINTER(_File), C.c_int, C.c_int, C.c_uint16, C.c_void_p, C.c_int, u64p, u32p]
        L.orc_common.restype = C.c_int64
        L.orc_count.argtypes = [u8p, u64p, C.c_size_t, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                C.c_uint64, C.c_int, u64p, C.c_size_t]
        L.orc_count.restype = C.c_int64
        L.orc_count_minimizer.argtypes = [u8p, u64p, C.c_size_t, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                          C.c_uint64, C.c_int, u64p, C.c_size_t]
        L.orc_count_minimizer.restype = C.c_int64
        L.orc_sm64.argtypes = [C.c_uint64]
        L.orc_sm64.restype = C.c_uint64
        L.orc_universe.argtypes = [C.c_uint64, C.c_size_t, C.c_uint64, C.c_uint64, u64p]
        L.orc_universe.restype = None
        L.orc_member_file.argtypes = [C.c_uint64, C.c_size_t, C.c_uint64, C.c_uint64, C.c_uint64, C.c_int, u64p]
        L.orc_member_file.restype = C.c_size_t
        L.orc_c3_digest.argtypes = [C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint64, C.c_int, u64p]
        L.orc_c3_digest.restype = None
        L.orc_random_keys.argtypes = [C.c_uint64, C.c_size_t, C.c_uint64, u64p]
        L.orc_random_keys.restype = None
        L.orc_synth_bases.argtypes = [C.c_uint64, C.c_uint64, C.c_size_t, C.c_uint64, u8p]
        L.orc_synth_bases.restype = None
        _lib = L
    return _lib


def _p(a: Optional[np.ndarray]):
    return None if a is None else a.ctypes.data


def _u64(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.uint64)


def _u32(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.uint32)


def _chk(r: int) -> int:
    if r < 0:
        raise OracleError(int(r))
    return int(r)


# ---- k-mer arithmetic ------------------------------------------------------
def encode(kmer: bytes) -> int:
    out = C.c_uint64()
    buf = np.frombuffer(kmer, dtype=np.uint8)
    _chk(lib().orc_encode(buf.ctypes.data, len(kmer), C.byref(out)))
    return out.value


def revcomp(code: int, k: int) -> int:
    return lib().orc_revcomp(code, k)


def canonical(code: int, k: int) -> int:
    return lib().orc_canonical(code, k)


def decode(code: int, k: int) -> bytes:
    b = C.create_string_buffer(k)
    lib().orc_decode(code, k, b)
    return b.raw


def _iter(fn, seq, k, canon, circular):
    s = np.frombuffer(seq, dtype=np.uint8) if isinstance(seq, (bytes, bytearray)) else np.ascontiguousarray(seq, dtype=np.uint8)
    out = np.empty(max(len(s) + k, 1), dtype=np.uint64)
    n = _chk(fn(s.ctypes.data, len(s), k, int(canon), int(circular), out.ctypes.data))
    return out[:n].copy()


def kmer_iter(seq, k: int, canonical: bool = False, circular: bool = False) -> np.ndarray:
    return _iter(lib().orc_kmer_iter, seq, k, canonical, circular)


def nthash_iter(seq, k: int, canonical: bool = False, circular: bool = False) -> np.ndarray:
    return _iter(lib().orc_nthash_iter, seq, k, canonical, circular)


# ---- taxonomy ----------------------------------------------------------------
class Taxonomy:
    """parent[t] == 0 -> unknown taxid; root has parent[t] == t (util.go:119-171)."""

    def __init__(self, parent, merged_from=None, merged_to=None):
        self.parent = _u32(parent)
        self.merged_from = _u32(merged_from if merged_from is not None else [])
        self.merged_to = _u32(merged_to if merged_to is not None else [])
        self._h = lib().orc_tax_new(self.parent.ctypes.data, len(self.parent), _p(self.merged_from),
                                    _p(self.merged_to), len(self.merged_from))

    def lca(self, a: int, b: int) -> int:
        return lib().orc_lca(self._h, a, b)

    def __del__(self):
        try:
            lib().orc_tax_free(self._h)
        except Exception:
            pass


def _tax(t: Optional[Taxonomy]):
    return None if t is None else t._h


# ---- sort / fold ---------------------------------------------------------------
def sort_u64(keys, threads: int = 1) -> np.ndarray:
    a = _u64(keys).copy()
    _chk(lib().orc_sort_u64(a.ctypes.data, len(a), threads))
    return a


def sort_pairs(keys, taxids):
    k, v = _u64(keys).copy(), _u32(taxids).copy()
    _chk(lib().orc_sort_pairs(k.ctypes.data, v.ctypes.data, len(k)))
    return k, v


def fold(mode: int, keys, taxids=None, tax: Optional[Taxonomy] = None):
    k = _u64(keys)
    v = None if taxids is None else _u32(taxids)
    ok = np.empty(len(k) + 2, dtype=np.uint64)
    ov = np.empty(len(k) + 2, dtype=np.uint32)
    n = _chk(lib().orc_fold(mode, k.ctypes.data, _p(v), len(k), _tax(tax), ok.ctypes.data, ov.ctypes.data))
    return (ok[:n].copy(), None if v is None else ov[:n].copy())


# ---- set operations --------------------------------------------------------------
def _files(files: Sequence, sorted_flags=None):
    """files: sequence of keys arrays or (keys, taxids) tuples."""
    keep = []
    arr = (_File * len(files))()
    for i, f in enumerate(files):
        if isinstance(f, tuple):
            k, v = _u64(f[0]), (None if f[1] is None else _u32(f[1]))
        else:
            k, v = _u64(f), None
        keep.append((k, v))
        arr[i].keys = k.ctypes.data
        arr[i].taxids = _p(v)
        arr[i].n = len(k)
        arr[i].sorted = 1 if sorted_flags is None else int(sorted_flags[i])
    total = sum(len(k) for k, _ in keep)
    return arr, keep, total


def _setop_out(total):
    return np.empty(total + 2, dtype=np.uint64), np.empty(total + 2, dtype=np.uint32)


def union(files, has_taxid=False, tax=None, threads=1):
    arr, keep, total = _files(files)
    ok, ov = _setop_out(total)
    n = _chk(lib().orc_union(arr, len(files), int(has_taxid), _tax(tax), threads, ok.ctypes.data, ov.ctypes.data))
    return ok[:n].copy(), (ov[:n].copy() if has_taxid else None)


def inter(files, has_taxid=False, mix_taxid=False, tax=None):
    arr, keep, total = _files(files)
    ok, ov = _setop_out(total)
    n = _chk(lib().orc_inter(arr, len(files), int(has_taxid), int(mix_taxid), _tax(tax), ok.ctypes.data, ov.ctypes.data))
    return ok[:n].copy(), (ov[:n].copy() if (has_taxid or mix_taxid) else None)


def diff(files, has_taxid=False, compare_taxid=False, tax=None, sorted_flags=None, threads=1):
    arr, keep, total = _files(files, sorted_flags)
    ok, ov = _setop_out(total)
    n = _chk(lib().orc_diff(arr, len(files), int(has_taxid), int(compare_taxid), _tax(tax), threads,
                            ok.ctypes.data, ov.ctypes.data))
    return ok[:n].copy(), (ov[:n].copy() if has_taxid else None)


def common(files, threshold: int, has_taxid=False, tax=None, threads=1):
    arr, keep, total = _files(files)
    ok, ov = _setop_out(total)
    n = _chk(lib().orc_common(arr, len(files), int(has_taxid), threshold, _tax(tax), threads,
                              ok.ctypes.data, ov.ctypes.data))
    return ok[:n].copy(), (ov[:n].copy() if has_taxid else None)


def merge_chunks(files, mode: int, has_taxid=False, tax=None):
    arr, keep, total = _files(files)
    ok, ov = _setop_out(total)
    n = _chk(lib().orc_merge_chunks(arr, len(files), int(has_taxid), mode, _tax(tax), ok.ctypes.data, ov.ctypes.data))
    return ok[:n].copy(), (ov[:n].copy() if has_taxid else None)


def count(bases, rec_off, k, canonical=True, hashed=False, circular=False, scaled=False, max_hash=0, threads=1):
    b = np.frombuffer(bases, dtype=np.uint8) if isinstance(bases, (bytes, bytearray)) else np.ascontiguousarray(bases, dtype=np.uint8)
    ro = _u64(rec_off)
    cap = len(b) + 1
    out = np.empty(cap, dtype=np.uint64)
    n = _chk(lib().orc_count(b.ctypes.data, ro.ctypes.data, len(ro) - 1, k, int(canonical), int(hashed),
                             int(circular), int(scaled), max_hash, threads, out.ctypes.data, cap))
    return out[:n].copy()


def count_minimizer(bases, rec_off, k, w, canonical=True, circular=False, scaled=False, max_hash=0, threads=1):
    """`count -H -W w` (count.go:316-317, 358-359): distinct sliding-window minima of the ntHash stream, ascending."""
    b = np.frombuffer(bases, dtype=np.uint8) if isinstance(bases, (bytes, bytearray)) else np.ascontiguousarray(bases, dtype=np.uint8)
    ro = _u64(rec_off)
    cap = len(b) + 1
    out = np.empty(cap, dtype=np.uint64)
    n = _chk(lib().orc_count_minimizer(b.ctypes.data, ro.ctypes.data, len(ro) - 1, k, w, int(canonical), int(circular),
                                       int(scaled), max_hash, threads, out.ctypes.data, cap))
    return out[:n].copy()


# ---- synthetic generators ---------------------------------------------
def sm64(x: int) -> int:
    return lib().orc_sm64(x & 0xFFFFFFFFFFFFFFFF)


def universe(j0: int, count_: int, N: int, S: int) -> np.ndarray:
    out = np.empty(count_, dtype=np.uint64)
    lib().orc_universe(j0, count_, N, S, out.ctypes.data)
    return out


def member_file(j0: int, count_: int, N: int, S: int, T: int, f: int) -> np.ndarray:
    out = np.empty(count_, dtype=np.uint64)
    n = lib().orc_member_file(j0, count_, N, S, T, f, out.ctypes.data)
    return out[:n].copy()


def c3_digest(j0: int, count_: int, N: int, S: int, T: int, nf: int) -> dict:
    """{count, sum, xor} of inter / diff / union over files 0..nf-1 of the C3 generator restricted to the universe
    indices [j0, j0 + count) -- computed from the generator's membership bits, no set operation involved."""
    out = np.zeros(9, dtype=np.uint64)
    lib().orc_c3_digest(j0, count_, N, S, T, nf, out.ctypes.data)
    o = [int(x) for x in out]
    return {"inter": tuple(o[0:3]), "diff": tuple(o[3:6]), "union": tuple(o[6:9])}


def digest3(keys: np.ndarray) -> tuple:
    """(count, sum mod 2^64, xor) of a key array."""
    k = np.asarray(keys, dtype=np.uint64)
    return (int(len(k)), int(np.add.reduce(k, dtype=np.uint64)) if len(k) else 0, int(np.bitwise_xor.reduce(k)) if len(k) else 0)


def random_keys(i0: int, count_: int, S: int) -> np.ndarray:
    out = np.empty(count_, dtype=np.uint64)
    lib().orc_random_keys(i0, count_, S, out.ctypes.data)
    return out


def synth_bases(r: int, i0: int, count_: int, S: int) -> np.ndarray:
    out = np.empty(count_, dtype=np.uint8)
    lib().orc_synth_bases(r, i0, count_, S, out.ctypes.data)
    return out
