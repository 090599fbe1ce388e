"""Host-side mirror of the reference's command inner loops over the libukm C ABI.

The method names and argument meanings follow the unikmer commands they stand in for
(`sort [-u|-d]`, `union`, `inter [--mix-taxid]`, `diff [-t]`, `common -n`, `count -k -K -H
--circular -D`), so a parity test reads like the command it checks.  Inputs are "k-mer
sets": what `unik.Reader.ReadCodeWithTaxid` yields for one .unik file, as arrays.

Host arrays are numpy (uint64 codes, uint32 taxids); device arrays are torch CUDA tensors
(int64/uint64 codes, int32/uint32 taxids) and stay on the GPU (outputs too).  Everything
computes in libukm.so (hand-written sm_100a CUDA); there is no CPU path here.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import List, Optional, Sequence, Tuple, Union

import numpy as np

from . import _lib as L

Array = Union[np.ndarray, "torch.Tensor"]  # noqa: F821


def _is_torch(a) -> bool:
    return hasattr(a, "data_ptr") and hasattr(a, "device")


@dataclass
class KmerSet:
    """One k-mer stream (a .unik file's payload): sorted codes, optional per-code taxids or one global taxid."""
    keys: Array
    taxids: Optional[Array] = None
    global_taxid: int = 0
    sorted: bool = True

    def __len__(self):
        return int(self.keys.shape[0])


def _as_set(s) -> KmerSet:
    if isinstance(s, KmerSet):
        return s
    if isinstance(s, tuple):
        return KmerSet(s[0], s[1])
    return KmerSet(s)


class Engine:
    """One GPU, one stream (ukm_ctx).  One process per GPU."""

    def __init__(self, device: int = 0):
        self.lib = L.load()
        self.device = device
        self.ctx = self.lib.ukm_create(device)
        if not self.ctx:
            raise L.UkmError(L.E_CUDA, self.lib.ukm_last_error(None).decode())
        self.has_taxonomy = False

    def close(self):
        if getattr(self, "ctx", None):
            self.lib.ukm_destroy(self.ctx)
            self.ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- plumbing -------------------------------------------------------------------
    def _chk(self, status: int):
        if status != L.OK:
            raise L.UkmError(status, self.lib.ukm_last_error(self.ctx).decode())

    @property
    def stream_ptr(self) -> int:
        return int(self.lib.ukm_get_stream(self.ctx) or 0)

    def use_stream(self, cuda_stream_ptr: int):
        self._chk(self.lib.ukm_set_stream(self.ctx, C.c_void_p(cuda_stream_ptr)))

    def sync(self):
        self._chk(self.lib.ukm_sync(self.ctx))

    def launch_count(self) -> int:
        return int(self.lib.ukm_launch_count(self.ctx))

    def stats_enable(self, on: bool = True):
        self._chk(self.lib.ukm_stats_enable(self.ctx, int(on)))

    def stats_reset(self):
        self._chk(self.lib.ukm_stats_reset(self.ctx))

    def stats(self) -> dict:
        arr = (L.KernelStat * 64)()
        n = C.c_int(0)
        self._chk(self.lib.ukm_stats_get(self.ctx, arr, 64, C.byref(n)))
        return {arr[i].name.decode(): {"launches": int(arr[i].launches), "ms": float(arr[i].ms),
                                       "algo_bytes": float(arr[i].algo_bytes)} for i in range(min(n.value, 64))}

    @staticmethod
    def _host_u64(a) -> np.ndarray:
        return np.ascontiguousarray(a, dtype=np.uint64)

    @staticmethod
    def _host_u32(a) -> np.ndarray:
        return np.ascontiguousarray(a, dtype=np.uint32)

    def _span_in(self, s: KmerSet, keep: list) -> L.Span:
        sp = L.Span()
        if _is_torch(s.keys):
            k = s.keys.contiguous()
            assert k.element_size() == 8, "codes must be a 64-bit tensor"
            keep.append(k)
            # CUDA tensor -> device span; CPU (pinned) tensor -> host span
            sp.keys, sp.n, sp.where = k.data_ptr(), k.shape[0], (L.DEVICE if k.is_cuda else L.HOST_PINNED)
            if s.taxids is not None:
                t = s.taxids.contiguous()
                assert t.element_size() == 4 and t.shape[0] == k.shape[0]
                keep.append(t)
                sp.taxids = t.data_ptr()
        else:
            k = self._host_u64(s.keys)
            keep.append(k)
            sp.keys, sp.n, sp.where = k.ctypes.data, len(k), L.HOST
            if s.taxids is not None:
                t = self._host_u32(s.taxids)
                assert len(t) == len(k)
                keep.append(t)
                sp.taxids = t.ctypes.data
        sp.global_taxid = int(s.global_taxid)
        sp.cap = sp.n
        sp.sorted = int(bool(s.sorted))
        return sp

    def _spans(self, sets: Sequence) -> Tuple[C.Array, list, bool]:
        sets = [_as_set(s) for s in sets]
        if not sets:
            raise ValueError("need at least one k-mer set")
        dev = [_is_torch(s.keys) and s.keys.is_cuda for s in sets]
        if any(dev) and not all(dev):
            raise ValueError("all inputs must live in the same memory space")
        keep: list = []
        arr = (L.Span * len(sets))()
        for i, s in enumerate(sets):
            arr[i] = self._span_in(s, keep)
        return arr, keep, all(dev)

    def _span_out(self, cap: int, want_taxids: bool, device: bool, out=None):
        """Output span.  `out` = caller-provided (keys[, taxids]) buffers (e.g. pinned host memory)."""
        sp = L.Span()
        cap = max(int(cap), 1)
        if out is not None:
            k, t = (out if isinstance(out, tuple) else (out, None))
            if _is_torch(k):
                sp.keys, sp.where = k.data_ptr(), (L.DEVICE if k.is_cuda else L.HOST_PINNED)
                if t is not None:
                    sp.taxids = t.data_ptr()
            else:
                assert k.dtype == np.uint64 and k.flags.c_contiguous
                sp.keys, sp.where = k.ctypes.data, L.HOST
                if t is not None:
                    sp.taxids = t.ctypes.data
            sp.cap = int(k.shape[0])
            return sp, k, t
        if device:
            import torch
            k = torch.empty(cap, dtype=torch.int64, device=f"cuda:{self.device}")
            t = torch.empty(cap, dtype=torch.int32, device=f"cuda:{self.device}") if want_taxids else None
            sp.keys, sp.where = k.data_ptr(), L.DEVICE
            if t is not None:
                sp.taxids = t.data_ptr()
        else:
            k = np.empty(cap, dtype=np.uint64)
            t = np.empty(cap, dtype=np.uint32) if want_taxids else None
            sp.keys, sp.where = k.ctypes.data, L.HOST
            if t is not None:
                sp.taxids = t.ctypes.data
        sp.cap = cap
        return sp, k, t

    @staticmethod
    def _trim(sp: L.Span, k, t):
        n = int(sp.n)
        return k[:n], (None if t is None else t[:n])

    def upload(self, host):
        """Copy a host array (numpy, or a pinned CPU torch tensor) of 64-bit codes into a new device tensor through
        ukm_copy: the returned tensor is a DEVICE span for any later operation (ops then chain without PCIe traffic)."""
        import torch
        if _is_torch(host):
            assert not host.is_cuda and host.is_contiguous() and host.element_size() == 8
            n, ptr, where = host.shape[0], host.data_ptr(), L.HOST_PINNED
            keep = host
        else:
            keep = self._host_u64(host)
            n, ptr, where = len(keep), keep.ctypes.data, L.HOST
        d = torch.empty(n, dtype=torch.int64, device=f"cuda:{self.device}")
        self._chk(self.lib.ukm_copy(self.ctx, d.data_ptr(), L.DEVICE, ptr, where, n * 8))
        return d

    # ---- taxonomy (util.go:119-171) ---------------------------------------------------
    def set_taxonomy(self, parent, merged_from=None, merged_to=None):
        p = self._host_u32(parent)
        mf = self._host_u32(merged_from if merged_from is not None else [])
        mt = self._host_u32(merged_to if merged_to is not None else [])
        self._chk(self.lib.ukm_set_taxonomy(self.ctx, p.ctypes.data, len(p), mf.ctypes.data if len(mf) else None,
                                            mt.ctypes.data if len(mt) else None, len(mf)))
        self.has_taxonomy = True

    def lca(self, a, b) -> np.ndarray:
        a, b = self._host_u32(a), self._host_u32(b)
        out = np.empty(len(a), dtype=np.uint32)
        self._chk(self.lib.ukm_lca_batch(self.ctx, a.ctypes.data, b.ctypes.data, len(a), out.ctypes.data, L.HOST))
        return out

    # ---- sort (sort.go:452-464) -----------------------------------------------------------
    def sort(self, keys: Array, taxids: Optional[Array] = None, key_bits: int = 64):
        """In-place ascending sort (sortutil.Uint64s / sorts.Quicksort(CodeTaxidSlice)).  Returns (keys, taxids)."""
        if _is_torch(keys):
            assert keys.is_cuda and keys.is_contiguous() and keys.element_size() == 8
            if taxids is None:
                self._chk(self.lib.ukm_sort_u64(self.ctx, keys.data_ptr(), keys.shape[0], key_bits, L.DEVICE))
            else:
                assert taxids.is_cuda and taxids.is_contiguous() and taxids.element_size() == 4
                self._chk(self.lib.ukm_sort_pairs(self.ctx, keys.data_ptr(), taxids.data_ptr(), keys.shape[0], key_bits, L.DEVICE))
            return keys, taxids
        k = self._host_u64(keys).copy() if not (isinstance(keys, np.ndarray) and keys.dtype == np.uint64 and keys.flags.c_contiguous and keys.flags.writeable) else keys
        if taxids is None:
            self._chk(self.lib.ukm_sort_u64(self.ctx, k.ctypes.data, len(k), key_bits, L.HOST))
            return k, None
        t = self._host_u32(taxids).copy() if not (isinstance(taxids, np.ndarray) and taxids.dtype == np.uint32 and taxids.flags.c_contiguous and taxids.flags.writeable) else taxids
        self._chk(self.lib.ukm_sort_pairs(self.ctx, k.ctypes.data, t.ctypes.data, len(k), key_bits, L.HOST))
        return k, t

    def sort_codetaxid16(self, records: np.ndarray, key_bits: int = 64) -> np.ndarray:
        """Sort Go's []CodeTaxid (kmers.go:24-28) in place: structured array of {code u64, taxid u32, pad u32}."""
        assert records.dtype.itemsize == 16 and records.flags.c_contiguous
        self._chk(self.lib.ukm_sort_codetaxid16(self.ctx, records.ctypes.data, len(records), key_bits))
        return records

    # ---- fold (sort.go:482-573; util-sort.go:35-190) -----------------------------------------
    def fold(self, mode: int, keys: Array, taxids: Optional[Array] = None):
        arr, keep, dev = self._spans([KmerSet(keys, taxids)])
        has_tax = taxids is not None
        out, k, t = self._span_out(len(keys) + 2, has_tax, dev)
        self._chk(self.lib.ukm_fold_sorted(self.ctx, mode, arr, L.F_TAXID if has_tax else 0, C.byref(out)))
        return self._trim(out, k, t)

    # ---- set operations ---------------------------------------------------------------------
    def union(self, sets: Sequence, has_taxid: bool = False, out=None, validate: bool = False):
        """`unikmer union -s` (union.go:186-208, 260-305)."""
        arr, keep, dev = self._spans(sets)
        out, k, t = self._span_out(sum(int(a.n) for a in arr), has_taxid, dev, out)
        self._chk(self.lib.ukm_union(self.ctx, arr, len(arr), (L.F_TAXID if has_taxid else 0) | (L.F_VALIDATE if validate else 0), C.byref(out)))
        return self._trim(out, k, t)

    def inter(self, sets: Sequence, has_taxid: bool = False, mix_taxid: bool = False, out=None, validate: bool = False,
              shard: bool = False):
        """`unikmer inter [--mix-taxid]` (inter.go:188-286), iterated in file order.  shard=True: the sets are key-range
        slices of files (UKM_F_SHARD): plain set semantics for empty slices instead of the whole-file quirks."""
        arr, keep, dev = self._spans(sets)
        flags = ((L.F_TAXID if has_taxid else 0) | (L.F_MIX_TAXID if mix_taxid else 0) | (L.F_VALIDATE if validate else 0) |
                 (L.F_SHARD if shard else 0))
        out, k, t = self._span_out(int(arr[0].n), has_taxid or mix_taxid, dev, out)
        self._chk(self.lib.ukm_inter(self.ctx, arr, len(arr), flags, C.byref(out)))
        return self._trim(out, k, t)

    def diff(self, sets: Sequence, has_taxid: bool = False, compare_taxid: bool = False, out=None, validate: bool = False):
        """`unikmer diff -s [-t]` (diff.go:136-146, 341-515, 566-594)."""
        arr, keep, dev = self._spans(sets)
        flags = (L.F_TAXID if has_taxid else 0) | (L.F_COMPARE_TAXID if compare_taxid else 0) | (L.F_VALIDATE if validate else 0)
        out, k, t = self._span_out(int(arr[0].n), has_taxid, dev, out)
        self._chk(self.lib.ukm_diff(self.ctx, arr, len(arr), flags, C.byref(out)))
        return self._trim(out, k, t)

    def setops(self, sets: Sequence, ops: Sequence[str] = ("inter", "diff", "union"), outs=None, validate: bool = False,
               shard: bool = False):
        """Several of `inter` / `diff` / `union` over the same k-mer sets in ONE call (ukm_setops_stream): with host-resident
        sets every input byte crosses PCIe once and uploads, kernels and downloads overlap.  `outs`: one caller-provided
        key buffer per operation (pinned host tensors for full overlap), or None.  Returns one key array per operation."""
        arr, keep, dev = self._spans(sets)
        code = {"inter": L.OP_INTER, "diff": L.OP_DIFF, "union": L.OP_UNION}
        opc = (C.c_int * len(ops))(*[code[o] for o in ops])
        total = sum(int(a.n) for a in arr)
        spans = (L.Span * len(ops))()
        bufs = []
        for k, o in enumerate(ops):
            cap = total if o == "union" else int(arr[0].n)
            sp, kb, _ = self._span_out(cap, False, dev, None if outs is None else outs[k])
            spans[k] = sp
            bufs.append(kb)
        flags = (L.F_VALIDATE if validate else 0) | (L.F_SHARD if shard else 0)
        self._chk(self.lib.ukm_setops_stream(self.ctx, arr, len(arr), opc, len(ops), flags, spans))
        return [b[:int(spans[k].n)] for k, b in enumerate(bufs)]

    def common(self, sets: Sequence, threshold: int, has_taxid: bool = False, validate: bool = False):
        """`unikmer common -n threshold` (common.go:220-283, 329-354)."""
        arr, keep, dev = self._spans(sets)
        out, k, t = self._span_out(sum(int(a.n) for a in arr), has_taxid, dev)
        self._chk(self.lib.ukm_common(self.ctx, arr, len(arr), (L.F_TAXID if has_taxid else 0) | (L.F_VALIDATE if validate else 0), threshold, C.byref(out)))
        return self._trim(out, k, t)

    def merge(self, sets: Sequence, mode: int = L.FOLD_PLAIN, has_taxid: bool = False):
        """mergeChunksFile (util-sort.go:227-606): k-way merge of sorted chunks + fold(mode)."""
        arr, keep, dev = self._spans(sets)
        out, k, t = self._span_out(sum(int(a.n) for a in arr) + 2, has_taxid, dev)
        self._chk(self.lib.ukm_merge_sorted(self.ctx, mode, arr, len(arr), L.F_TAXID if has_taxid else 0, C.byref(out)))
        return self._trim(out, k, t)

    # ---- count (count.go:314-322, 355-437, 531-595) ----------------------------------------------
    @staticmethod
    def _count_flags(canonical, hashed, circular, scaled) -> int:
        return ((L.F_CANONICAL if canonical else 0) | (L.F_HASHED if hashed else 0) |
                (L.F_CIRCULAR if circular else 0) | (L.F_SCALED if scaled else 0))

    def _seq_call(self, fn, bases, rec_off, k, flags, max_hash):
        if _is_torch(bases):
            import torch
            assert bases.is_cuda and bases.element_size() == 1 and _is_torch(rec_off) and rec_off.element_size() == 8
            n_rec = rec_off.shape[0] - 1
            cap = int(bases.shape[0]) + 1
            out, ok, _ = self._span_out(cap, False, True)
            self._chk(fn(self.ctx, bases.data_ptr(), rec_off.data_ptr(), n_rec, k, flags, max_hash, L.DEVICE, C.byref(out)))
            return ok[:int(out.n)]
        b = np.frombuffer(bases, dtype=np.uint8) if isinstance(bases, (bytes, bytearray)) else np.ascontiguousarray(bases, dtype=np.uint8)
        ro = self._host_u64(rec_off)
        out, ok, _ = self._span_out(len(b) + 1, False, False)
        self._chk(fn(self.ctx, b.ctypes.data, ro.ctypes.data, len(ro) - 1, k, flags, max_hash, L.HOST, C.byref(out)))
        return ok[:int(out.n)]

    def count(self, bases, rec_off, k: int, canonical: bool = True, hashed: bool = False, circular: bool = False,
              scaled: bool = False, max_hash: int = 0):
        """`unikmer count -k K [-K] [-H] [--circular] [-D] -s`: distinct codes, ascending."""
        return self._seq_call(self.lib.ukm_count_seq, bases, rec_off, k,
                              self._count_flags(canonical, hashed, circular, scaled), max_hash)

    def count_minimizer(self, bases, rec_off, k: int, w: int, canonical: bool = True, circular: bool = False,
                        scaled: bool = False, max_hash: int = 0):
        """`unikmer count -k K -K -H -W w -s` (count.go:316-317, 358-359): distinct window minima of the ntHash stream."""
        def fn(ctx, b, ro, n_rec, k_, flags, mh, where, out):
            return self.lib.ukm_count_minimizer(ctx, b, ro, n_rec, k_, w, flags, mh, where, out)
        return self._seq_call(fn, bases, rec_off, k, self._count_flags(canonical, True, circular, scaled), max_hash)

    def kmers(self, bases, rec_off, k: int, canonical: bool = True, hashed: bool = False, circular: bool = False,
              scaled: bool = False, max_hash: int = 0):
        """The iterator alone (`count --linear`): every code in record-then-position order."""
        return self._seq_call(self.lib.ukm_kmers_seq, bases, rec_off, k,
                              self._count_flags(canonical, hashed, circular, scaled), max_hash)

    # ---- sharding helpers -----------------------------------------------------------------------------
    def partition_sorted(self, keys: Array, splitters) -> np.ndarray:
        arr, keep, dev = self._spans([KmerSet(keys)])
        sp = self._host_u64(splitters)
        off = np.zeros(len(sp) + 2, dtype=np.uint64)
        self._chk(self.lib.ukm_partition_sorted(self.ctx, arr, sp.ctypes.data if len(sp) else None, len(sp), off.ctypes.data))
        return off

    def check_sorted_unique(self, keys: Array) -> bool:
        arr, keep, dev = self._spans([KmerSet(keys)])
        r = self.lib.ukm_check_sorted_unique(self.ctx, arr)
        if r == L.E_NOT_SORTED_UNIQUE:
            return False
        self._chk(r)
        return True

    # ---- synthetic inputs, device tensors --------------------------------------------------
    def synth_random_keys(self, i0: int, count: int, seed: int):
        import torch
        out = torch.empty(count, dtype=torch.int64, device=f"cuda:{self.device}")
        self._chk(self.lib.ukm_synth_random_keys(self.ctx, i0, count, seed, out.data_ptr()))
        return out

    def synth_member_file(self, j0: int, count: int, N: int, S: int, T: int, f: int):
        import torch
        out = torch.empty(max(count, 1), dtype=torch.int64, device=f"cuda:{self.device}")
        n = C.c_size_t(0)
        self._chk(self.lib.ukm_synth_member_file(self.ctx, j0, count, N, S, T, f, out.data_ptr(), C.byref(n)))
        return out[:n.value]

    def synth_bases(self, r: int, i0: int, count: int, S: int):
        import torch
        out = torch.empty(count, dtype=torch.uint8, device=f"cuda:{self.device}")
        self._chk(self.lib.ukm_synth_bases(self.ctx, r, i0, count, S, out.data_ptr()))
        return out
