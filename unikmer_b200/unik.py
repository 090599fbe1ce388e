"""`.unik` v5 files from Python: thin ctypes layer over the host codec (unikmer_b200/host/unik.hpp ->
libukm_host.so).  (De)serialisation stays on the host (north star); format parity with real unikmer files
is UNPINNED -- the layout constants live in host/unik.hpp."""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass
from typing import Optional, Tuple

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libukm_host.so")
CLI_PATH = os.path.join(_HERE, "bin", "unikmer-b200")

COMPACT, CANONICAL, SORTED, INCLUDE_TAXID, HASHED, SCALED = 1, 2, 4, 8, 16, 32


class _Hdr(C.Structure):
    _fields_ = [("k", C.c_int), ("flag", C.c_uint32), ("number", C.c_uint64), ("global_taxid", C.c_uint32),
                ("taxid_bytes", C.c_uint32), ("scale", C.c_uint32), ("max_hash", C.c_uint64), ("description", C.c_char * 1032)]


@dataclass
class Header:
    k: int
    flag: int = 0
    number: int = 0
    global_taxid: int = 0
    taxid_bytes: int = 4
    scale: int = 1
    max_hash: int = 0
    description: str = ""

    def has(self, f: int) -> bool:
        return bool(self.flag & f)


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(f"{LIB_PATH} not found: make -C unikmer_b200/host")
        L = C.CDLL(LIB_PATH)
        pp64, pp32, psz = C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.POINTER(C.c_size_t)
        L.ukmh_last_error.restype = C.c_char_p
        L.ukmh_free.argtypes = [C.c_void_p]
        L.ukmh_unik_encode.argtypes = [C.POINTER(_Hdr), C.c_void_p, C.c_void_p, C.c_size_t, C.POINTER(C.c_void_p), psz]
        L.ukmh_unik_decode.argtypes = [C.c_void_p, C.c_size_t, C.c_int, C.POINTER(_Hdr), pp64, pp32, psz, psz]
        L.ukmh_unik_read_file.argtypes = [C.c_char_p, C.c_int, C.POINTER(_Hdr), pp64, pp32, psz, psz]
        L.ukmh_unik_write_file.argtypes = [C.c_char_p, C.POINTER(_Hdr), C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_int]
        _lib = L
    return _lib


def _to_c(h: Header) -> _Hdr:
    c = _Hdr()
    c.k, c.flag, c.number, c.global_taxid = h.k, h.flag, h.number, h.global_taxid
    c.taxid_bytes, c.scale, c.max_hash = h.taxid_bytes, h.scale, h.max_hash
    c.description = h.description.encode()
    return c


def _from_c(c: _Hdr) -> Header:
    return Header(c.k, c.flag, c.number, c.global_taxid, c.taxid_bytes, c.scale, c.max_hash, c.description.decode())


def _chk(r: int):
    if r != 0:
        raise ValueError(lib().ukmh_last_error().decode())


def _take(L, codes, taxids, n, nt):
    k = np.ctypeslib.as_array(C.cast(codes, C.POINTER(C.c_uint64)), shape=(max(n.value, 1),))[:n.value].copy()
    t = np.ctypeslib.as_array(C.cast(taxids, C.POINTER(C.c_uint32)), shape=(max(nt.value, 1),))[:nt.value].copy()
    L.ukmh_free(codes)
    L.ukmh_free(taxids)
    return k, (t if nt.value else None)


def encode(h: Header, codes, taxids=None) -> bytes:
    L = lib()
    k = np.ascontiguousarray(codes, dtype=np.uint64)
    t = None if taxids is None else np.ascontiguousarray(taxids, dtype=np.uint32)
    out, n = C.c_void_p(), C.c_size_t()
    hc = _to_c(h)
    _chk(L.ukmh_unik_encode(C.byref(hc), k.ctypes.data, None if t is None else t.ctypes.data, len(k), C.byref(out), C.byref(n)))
    b = C.string_at(out, n.value)
    L.ukmh_free(out)
    return b


def decode(buf: bytes, ignore_taxid: bool = False) -> Tuple[Header, np.ndarray, Optional[np.ndarray]]:
    L = lib()
    hc, codes, taxids, n, nt = _Hdr(), C.c_void_p(), C.c_void_p(), C.c_size_t(), C.c_size_t()
    _chk(L.ukmh_unik_decode(buf, len(buf), int(ignore_taxid), C.byref(hc), C.byref(codes), C.byref(taxids), C.byref(n), C.byref(nt)))
    k, t = _take(L, codes, taxids, n, nt)
    return _from_c(hc), k, t


def read_unik(path: str, ignore_taxid: bool = False) -> Tuple[Header, np.ndarray, Optional[np.ndarray]]:
    """Whole file -> (header, codes, taxids).  gzip is sniffed (util-io.go:99-101)."""
    L = lib()
    hc, codes, taxids, n, nt = _Hdr(), C.c_void_p(), C.c_void_p(), C.c_size_t(), C.c_size_t()
    _chk(L.ukmh_unik_read_file(path.encode(), int(ignore_taxid), C.byref(hc), C.byref(codes), C.byref(taxids), C.byref(n), C.byref(nt)))
    k, t = _take(L, codes, taxids, n, nt)
    return _from_c(hc), k, t


def write_unik(path: str, h: Header, codes, taxids=None, compress: bool = True, level: int = -1):
    L = lib()
    k = np.ascontiguousarray(codes, dtype=np.uint64)
    t = None if taxids is None else np.ascontiguousarray(taxids, dtype=np.uint32)
    hc = _to_c(h)
    _chk(L.ukmh_unik_write_file(path.encode(), C.byref(hc), k.ctypes.data, None if t is None else t.ctypes.data, len(k), int(compress), level))
