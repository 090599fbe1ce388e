"""CPU-side checks of the drop-in boundary: libukm.so loads, exports every symbol
include/ukm.h declares, and fails loudly (no CPU fallback) when no GPU is present."""
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    src = open(os.path.join(ROOT, "include", "ukm.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(ukm_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from unikmer_b200 import _lib
    lib = _lib.load()
    declared = header_symbols()
    assert len(declared) >= 30
    for name in declared:
        assert hasattr(lib, name), f"libukm.so does not export {name}"
    assert sorted(_lib.SYMBOLS) == declared, "unikmer_b200/_lib.py SYMBOLS out of sync with include/ukm.h"
    assert b"sm_100a" in lib.ukm_version()


def test_span_layout_matches_header():
    """ukm_span is {u64* keys; u32* taxids; u32 global_taxid; size_t n; size_t cap; int where; int sorted}."""
    import ctypes as C

    from unikmer_b200 import _lib
    assert C.sizeof(_lib.Span) == 48
    assert _lib.Span.n.offset == 24 and _lib.Span.cap.offset == 32 and _lib.Span.where.offset == 40
    assert C.sizeof(_lib.KernelStat) == 72


def test_no_cpu_fallback_without_gpu():
    import torch

    from unikmer_b200 import Engine, UkmError
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(UkmError):
        Engine(0)


def test_product_never_imports_the_oracle():
    """oracle/ is test infrastructure: nothing under unikmer_b200/ may reference it."""
    for dirpath, _, files in os.walk(os.path.join(ROOT, "unikmer_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                txt = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "import oracle" not in txt and "from oracle" not in txt and "liboracle" not in txt, f
