"""Host-side .unik v5 codec (unikmer_b200/host/unik.hpp): round trips over every payload form.
Format parity with real unikmer files is unpinned (no .unik fixture or unik/v5 source in the reference,
SURVEY.md F7); these tests pin self-consistency and the documented size facts."""
import gzip
import os
import subprocess

import numpy as np
import pytest

from unikmer_b200 import unik

U64 = np.uint64


def rng(s):
    return np.random.default_rng(s)


@pytest.mark.parametrize("n", [0, 1, 2, 3, 1000, 10_001])
@pytest.mark.parametrize("form", ["plain", "compact", "sorted", "sorted_tax", "plain_tax", "hashed_sorted"])
def test_round_trip(tmp_path, n, form):
    r = rng(n)
    k = 31
    hashed = form.startswith("hashed")
    codes = r.integers(0, 2**64 if hashed else 4**k, n, dtype=U64)
    tax = None
    flag = unik.CANONICAL
    if "sorted" in form:
        codes = np.sort(codes)
        flag |= unik.SORTED
    if form == "compact":
        flag |= unik.COMPACT
    if hashed:
        flag |= unik.HASHED
    if form.endswith("_tax"):
        flag |= unik.INCLUDE_TAXID
        tax = r.integers(0, 2_000_000, n).astype(np.uint32)
    h = unik.Header(k=k, flag=flag, number=n, taxid_bytes=3 if tax is not None else 4, description="round trip")
    buf = unik.encode(h, codes, tax)
    h2, c2, t2 = unik.decode(buf)
    assert (h2.k, h2.flag, h2.number, h2.taxid_bytes, h2.description) == (k, flag, n, h.taxid_bytes, "round trip")
    assert np.array_equal(c2, codes)
    if tax is not None:
        assert np.array_equal(t2 if t2 is not None else np.zeros(0, dtype=np.uint32), tax)
    else:
        assert t2 is None
    for compress in (False, True):
        p = str(tmp_path / f"x{int(compress)}.unik")
        unik.write_unik(p, h, codes, tax, compress=compress)
        raw = open(p, "rb").read()
        assert (raw[:2] == b"\x1f\x8b") == compress  # gzip sniffing relies on the magic (util-io.go:99-101)
        if compress:
            assert gzip.decompress(raw) == buf
        h3, c3, t3 = unik.read_unik(p)
        assert np.array_equal(c3, codes) and h3.flag == flag


def test_payload_sizes():
    """Plain payload = 8 B per code (testdata/table.tsv: 3 355 443 200 B for 100 Mi k-mers = 2^22*100*8);
    compact = ceil(k/4) B; sorted = 1 control byte per pair + the delta bytes."""
    n, k = 4096, 31
    codes = np.sort(rng(1).integers(0, 4**k, n, dtype=U64))
    base = len(unik.encode(unik.Header(k=k), np.zeros(0, dtype=U64)))
    assert len(unik.encode(unik.Header(k=k), codes)) - base == 8 * n
    assert len(unik.encode(unik.Header(k=k, flag=unik.COMPACT), codes)) - base == ((k + 3) // 4) * n
    d = np.diff(np.concatenate([[0], codes]).astype(object))
    nbytes = lambda v: max(1, (int(v).bit_length() + 7) // 8)  # noqa: E731
    exp = n // 2 + sum(nbytes(x) for x in d)
    assert len(unik.encode(unik.Header(k=k, flag=unik.SORTED), codes)) - base == exp


def test_global_taxid_is_broadcast():
    """count -t files: no per-k-mer taxids stored; ReadCodeWithTaxid returns the global one (README.md:169-171)."""
    codes = np.arange(10, dtype=U64)
    buf = unik.encode(unik.Header(k=5, flag=unik.SORTED, global_taxid=562), codes)
    h, c, t = unik.decode(buf)
    assert h.global_taxid == 562 and np.array_equal(t, np.full(10, 562, dtype=np.uint32))
    _, _, t2 = unik.decode(buf, ignore_taxid=True)
    assert t2 is None


def test_taxids_dropped_without_include_flag():
    """WriteCodeWithTaxid silently drops the taxid when IncludeTaxID is off (diff.go:593,597)."""
    codes = np.arange(6, dtype=U64)
    a = unik.encode(unik.Header(k=5), codes, np.arange(6, dtype=np.uint32))
    assert a == unik.encode(unik.Header(k=5), codes)


def test_corrupt_inputs_are_errors():
    with pytest.raises(ValueError):
        unik.decode(b"not a unik file at all........................................................................")
    good = unik.encode(unik.Header(k=21, flag=unik.SORTED), np.arange(0, 1000, 7, dtype=U64))
    with pytest.raises(ValueError):
        unik.decode(good[:-3])
    bad = bytearray(good)
    bad[8] = 4  # main version
    with pytest.raises(ValueError):
        unik.decode(bytes(bad))


def test_cli_view_and_info_need_no_gpu(tmp_path):
    p = str(tmp_path / "a.unik")
    codes = np.array([0, 1, 4**5 - 1], dtype=U64)
    unik.write_unik(p, unik.Header(k=5, flag=unik.SORTED | unik.INCLUDE_TAXID, number=3, taxid_bytes=2), codes, np.array([7, 8, 9], dtype=np.uint32))
    out = subprocess.run([unik.CLI_PATH, "view", "-t", p], capture_output=True, text=True, check=True).stdout.split("\n")
    assert out[:3] == ["AAAAA\t7", "AAAAC\t8", "TTTTT\t9"]
    info = subprocess.run([unik.CLI_PATH, "info", p], capture_output=True, text=True, check=True).stdout.strip().split("\n")
    assert info[1].split("\t")[1:] == ["5", "0", "0", "0", "1", "0", "1", "0", "3", "3"]


def test_fastx_reader(tmp_path):
    """The CLI's FASTA / FASTQ reader (unikmer_b200/host/fastx.hpp): wrapped records, CRLF, blanks inside sequence lines,
    FASTQ quality lines that start with '@' or '>', truncated records -- tests/host/fastx_test.cpp."""
    import shutil
    import subprocess
    if shutil.which("g++") is None:
        pytest.skip("needs g++")
    exe = tmp_path / "fastx_test"
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    subprocess.check_call(["g++", "-O2", "-std=c++17", os.path.join(root, "tests", "host", "fastx_test.cpp"), "-o", str(exe)])
    out = subprocess.run([str(exe)], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0 and "fastx ok" in out.stdout, out.stderr[-2000:]
