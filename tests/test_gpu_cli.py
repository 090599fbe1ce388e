"""C1 (BASELINE.json configs[0]) plumbing: real `.unik` files in, the unikmer-shaped CLI
(unikmer_b200/bin/unikmer-b200 -> libukm.so) in the middle, real `.unik` files out, compared with the
oracle and with the header contracts of SURVEY.md Appendix D."""
import gzip
import hashlib
import os
import subprocess

import numpy as np
import pytest

import oracle
from unikmer_b200 import unik

pytestmark = pytest.mark.gpu
U64 = np.uint64
CLI = unik.CLI_PATH


def run(*args, cwd=None):
    p = subprocess.run([CLI, *args], capture_output=True, text=True, cwd=cwd)
    assert p.returncode == 0, p.stderr
    return p


@pytest.fixture(scope="module")
def c1(tmp_path_factory):
    """pool P = U(j; 1.5e6, S=1); A = P[0:1e6], B = P[5e5:1.5e6]; k=31 canonical sorted."""
    d = tmp_path_factory.mktemp("c1")
    P = oracle.universe(0, 1_500_000, 1_500_000, 1)
    A, B = P[:1_000_000], P[500_000:]
    h = lambda n: unik.Header(k=31, flag=unik.CANONICAL | unik.SORTED, number=n)  # noqa: E731
    unik.write_unik(str(d / "A.unik"), h(len(A)), A, compress=True)
    unik.write_unik(str(d / "B.unik"), h(len(B)), B, compress=False)
    return d, A, B


def test_c1_union_inter_diff_common(c1):
    d, A, B = c1
    a, b = str(d / "A.unik"), str(d / "B.unik")
    exp = {"union": oracle.union([A, B])[0], "inter": oracle.inter([A, B])[0], "diff": oracle.diff([A, B])[0],
           "common": oracle.common([A, B], 2)[0]}
    assert len(exp["union"]) == 1_500_000 and len(exp["inter"]) == 500_000 and len(exp["diff"]) == 500_000
    for cmd, extra in (("union", ["-s"]), ("inter", []), ("diff", ["-s"]), ("common", ["-n", "2"])):
        for comp in ([], ["-C"]):
            out = str(d / f"{cmd}{len(comp)}")
            run(cmd, *extra, *comp, "-o", out, a, b)
            raw = open(out + ".unik", "rb").read()  # `.unik` is appended (sort.go:110-113 etc.)
            assert (raw[:2] == b"\x1f\x8b") == (not comp)
            h, codes, tax = unik.read_unik(out + ".unik")
            assert np.array_equal(codes, exp[cmd]), cmd
            assert tax is None
            assert h.k == 31 and h.flag == (unik.SORTED | unik.CANONICAL) and h.number == len(codes)  # Appendix D
    # the uncompressed stream is identical whether or not the file is gzip'd ("bit-identical .unik" on the -C stream, F6)
    assert gzip.decompress(open(str(d / "union0.unik"), "rb").read()) == open(str(d / "union1.unik"), "rb").read()


def test_k10_union_s_equals_sort_u(c1):
    """K10 (README.md:222-229): `union -s` and `sort -u` give identical `view` output."""
    d, A, B = c1
    a, b = str(d / "A.unik"), str(d / "B.unik")
    run("union", "-s", "-o", str(d / "u"), a, b)
    run("sort", "-u", "-o", str(d / "s"), a, b)
    v1 = run("view", str(d / "u.unik")).stdout
    v2 = run("view", str(d / "s.unik")).stdout
    assert hashlib.md5(v1.encode()).hexdigest() == hashlib.md5(v2.encode()).hexdigest()
    assert v1.split("\n")[0] == oracle.decode(int(min(A[0], B[0])), 31).decode()
    hs, cs, _ = unik.read_unik(str(d / "s.unik"))
    assert hs.number == 0 and hs.has(unik.SORTED)  # sort -u leaves Number unset (Appendix D)
    run("sort", "-d", "-o", str(d / "dup"), a, b)
    assert np.array_equal(unik.read_unik(str(d / "dup.unik"))[1], oracle.inter([A, B])[0])  # README.md:234,271
    run("sort", "-o", str(d / "plain"), b, a)
    hp, cp, _ = unik.read_unik(str(d / "plain.unik"))
    assert hp.number == len(A) + len(B) and np.array_equal(cp, np.sort(np.concatenate([A, B])))


def test_split_then_merge_is_the_external_sort(tmp_path):
    """`split -m` (stage 1 of sort -m: sorted chunk files, util-sort.go:35-190) + `merge` (mergeChunksFile,
    util-sort.go:227-606) give what the in-memory sort gives: plain, -u and -d (K10: README.md:222-229)."""
    r = np.random.default_rng(9)
    keys = r.integers(0, 40_000, 100_000).astype(U64)  # many duplicates, also across chunks
    src = str(tmp_path / "in.unik")
    unik.write_unik(src, unik.Header(k=31, flag=unik.CANONICAL, number=len(keys)), keys, compress=False)
    uniq, cnt = np.unique(keys, return_counts=True)
    for flag, exp in ((None, np.sort(keys)), ("-u", uniq), ("-d", uniq[cnt >= 2])):
        d = str(tmp_path / f"chunks{flag or ''}")
        extra = [flag] if flag else []
        run("split", "-m", "16K", "-O", d, *extra, src)  # 16K = 16384 k-mers per chunk (util.go:291)
        chunks = sorted(os.listdir(d))
        assert chunks == [f"chunk_{i:03d}.unik" for i in range(7)]
        seen = []
        for c in chunks:
            h, codes, _ = unik.read_unik(os.path.join(d, c))
            assert h.has(unik.SORTED) and (np.diff(codes.astype(np.int64)) >= 0).all()
            seen.append(codes)
        if flag is None:
            assert np.array_equal(np.sort(np.concatenate(seen)), np.sort(keys))
        out = str(tmp_path / f"merged{flag or ''}")
        run("merge", "-D", "-C", "-o", out, *extra, d)
        h, codes, _ = unik.read_unik(out + ".unik")
        assert np.array_equal(codes, exp), flag
        assert h.has(unik.SORTED) and h.has(unik.CANONICAL) and h.number == 0
        run("sort", "-C", "-o", out + "_mem", *extra, src)
        assert np.array_equal(unik.read_unik(out + "_mem.unik")[1], exp)
    # merge of explicit files, and the sorted-input check (merge.go:168-170)
    run("merge", "-u", "-o", str(tmp_path / "m2"), os.path.join(str(tmp_path / "chunks"), "chunk_000.unik"),
        os.path.join(str(tmp_path / "chunks"), "chunk_001.unik"))
    p = subprocess.run([CLI, "merge", "-o", str(tmp_path / "bad"), src], capture_output=True, text=True)
    assert p.returncode != 0 and "sorted" in p.stderr


def test_single_file_is_a_byte_copy(c1):
    d, A, B = c1
    run("union", "-C", "-o", str(d / "copy"), str(d / "A.unik"))
    assert open(str(d / "copy.unik"), "rb").read() == gzip.decompress(open(str(d / "A.unik"), "rb").read())


def test_input_checks(c1, tmp_path):
    d, A, B = c1
    unsorted = str(tmp_path / "x.unik")
    unik.write_unik(unsorted, unik.Header(k=31, flag=unik.CANONICAL), B[::-1].copy())
    p = subprocess.run([CLI, "inter", "-o", str(tmp_path / "o"), str(d / "A.unik"), unsorted], capture_output=True, text=True)
    assert p.returncode != 0 and "sorted" in p.stderr  # inter.go:139
    k21 = str(tmp_path / "k21.unik")
    unik.write_unik(k21, unik.Header(k=21, flag=unik.CANONICAL | unik.SORTED), np.arange(5, dtype=U64))
    p = subprocess.run([CLI, "union", "-o", str(tmp_path / "o"), str(d / "A.unik"), k21], capture_output=True, text=True)
    assert p.returncode != 0 and "K" in p.stderr  # checkCompatibility (util-binary-file.go:31-44)
    # diff accepts an unsorted subject (diff.go:341-367)
    run("diff", "-s", "-C", "-o", str(tmp_path / "dd"), str(d / "A.unik"), unsorted)
    assert np.array_equal(unik.read_unik(str(tmp_path / "dd.unik"))[1], oracle.diff([A, B])[0])


def test_count_from_fasta(tmp_path):
    """`count -k 31 -K -H -s` (C4's mode) and `-K` 2-bit codes from a multi-record, multi-line, gzip'd FASTA."""
    recs = [oracle.synth_bases(r, 0, n, 5).tobytes() for r, n in enumerate((70_001, 30, 5_000, 31))]
    recs[2] = recs[2][:100] + b"acgtnNRY" + recs[2][108:]
    fa = str(tmp_path / "x.fa.gz")
    with gzip.open(fa, "wb") as fh:
        for i, s in enumerate(recs):
            fh.write(b">seq%d some description\n" % i)
            for j in range(0, len(s), 80):
                fh.write(s[j:j + 80] + b"\n")
    bases = b"".join(recs)
    off = np.concatenate([[0], np.cumsum([len(s) for s in recs])]).astype(U64)
    run("count", "-k", "31", "-K", "-H", "-s", "-C", "-o", str(tmp_path / "h"), fa)
    h, codes, _ = unik.read_unik(str(tmp_path / "h.unik"))
    assert np.array_equal(codes, oracle.count(bases, off, 31, canonical=True, hashed=True))
    assert h.flag == (unik.SORTED | unik.CANONICAL | unik.HASHED) and h.number == len(codes) and h.k == 31
    run("count", "-k", "23", "-K", "-s", "-t", "562", "-o", str(tmp_path / "c"), fa)
    h, codes, tax = unik.read_unik(str(tmp_path / "c.unik"))
    assert np.array_equal(codes, oracle.count(bases, off, 23, canonical=True))
    assert h.global_taxid == 562 and (tax == 562).all() and not h.has(unik.INCLUDE_TAXID)
    run("count", "-k", "31", "-K", "-D", "15", "-s", "-o", str(tmp_path / "sc"), fa)  # -D switches hashing on (count.go:96-99)
    h, codes, _ = unik.read_unik(str(tmp_path / "sc.unik"))
    mh = int(float(2**64 - 1) / 15.0)
    assert np.array_equal(codes, oracle.count(bases, off, 31, canonical=True, hashed=True, scaled=True, max_hash=mh))
    assert h.has(unik.SCALED) and h.has(unik.HASHED) and h.scale == 15 and h.max_hash == mh
    run("count", "-k", "31", "-K", "-W", "15", "-s", "-C", "-o", str(tmp_path / "mz"), fa)  # -W switches hashing on (count.go:105-109)
    h, codes, _ = unik.read_unik(str(tmp_path / "mz.unik"))
    assert np.array_equal(codes, oracle.count_minimizer(bases, off, 31, 15, canonical=True))
    assert h.flag == (unik.SORTED | unik.CANONICAL | unik.HASHED) and h.number == len(codes)


def test_taxonomy_and_lca_through_files(tmp_path):
    """README workflow: per-genome files with a global taxid -> `common` / `union` / `inter` LCA-fold them
    (needs nodes.dmp under --data-dir, util.go:119-171)."""
    n_nodes = 2000
    parent = np.zeros(n_nodes + 1, dtype=np.uint32)
    parent[1] = 1
    for t in range(2, n_nodes + 1):
        parent[t] = 1 + oracle.sm64(8 + t) % (t - 1)
    dd = tmp_path / "taxdump"
    dd.mkdir()
    with open(dd / "nodes.dmp", "w") as fh:
        for t in range(1, n_nodes + 1):
            fh.write(f"{t}\t|\t{parent[t]}\t|\tno rank\t|\n")
    with open(dd / "merged.dmp", "w") as fh:
        fh.write("5000\t|\t17\t|\n")
    otax = oracle.Taxonomy(parent, [5000], [17])
    keys = [oracle.member_file(0, 200_000, 200_000, 6, 7, f) for f in range(3)]
    leaves = [1999, 1500, 5000]
    paths = []
    for i, (k, leaf) in enumerate(zip(keys, leaves)):
        p = str(tmp_path / f"g{i}.unik")
        unik.write_unik(p, unik.Header(k=31, flag=unik.CANONICAL | unik.SORTED, number=len(k), global_taxid=leaf), k)
        paths.append(p)
    ofiles = [(k, np.full(len(k), t, dtype=np.uint32)) for k, t in zip(keys, leaves)]
    for cmd, extra, exp in (("common", ["-n", "2"], oracle.common(ofiles, 2, has_taxid=True, tax=otax)),
                            ("union", ["-s"], oracle.union(ofiles, has_taxid=True, tax=otax)),
                            ("inter", [], oracle.inter(ofiles, has_taxid=True, tax=otax))):
        out = str(tmp_path / cmd)
        run(cmd, *extra, "--data-dir", str(dd), "-o", out, *paths)
        h, codes, tax = unik.read_unik(out + ".unik")
        assert np.array_equal(codes, exp[0]), cmd
        assert np.array_equal(tax, exp[1]), cmd
        assert h.has(unik.INCLUDE_TAXID) and h.global_taxid == 0  # global taxids come out as per-k-mer taxids (B-15)
        assert h.taxid_bytes == 2  # SetMaxTaxid(taxonomy max = 2000) -> 2 bytes (util.go:169)
