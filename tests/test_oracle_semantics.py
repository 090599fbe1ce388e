"""The oracle's restatement of the reference's edge semantics, checked against
independent pure-Python/numpy renderings of the Go loops on small cases."""
import numpy as np
import pytest

import oracle

U64 = np.uint64


def rng(s):
    return np.random.default_rng(s)


def py_inter(files):
    """inter.go:205-267 two-pointer with one-to-one matching (multiset min, quirk B-4), B-3 on empty later files."""
    mc = list(files[0])
    for f in files[1:]:
        if not mc:
            raise IndexError
        if len(f) == 0:
            break
        out, ii, j = [], 0, 0
        while ii < len(mc) and j < len(f):
            if mc[ii] < f[j]:
                ii += 1
            elif mc[ii] == f[j]:
                out.append(mc[ii]); ii += 1; j += 1
            else:
                j += 1
        mc = out
        if not mc:
            break
    return np.array(mc, dtype=U64)


def test_inter_multiset_semantics_and_quirks():
    r = rng(1)
    files = [np.sort(r.integers(0, 50, n).astype(U64)) for n in (40, 35, 60)]  # duplicates inside files
    assert np.array_equal(oracle.inter(files)[0], py_inter(files))
    a = np.arange(10, dtype=U64)
    e = np.zeros(0, dtype=U64)
    assert np.array_equal(oracle.inter([a, e, a[:3]])[0], a)  # B-3: empty later file keeps the current set
    with pytest.raises(oracle.OracleError) as ei:
        oracle.inter([e, a])
    assert ei.value.code == oracle.E_PANIC


def test_common_counts_occurrences_and_wraps_uint16():
    """B-7: file 0 contributes 1 per distinct code, later files one per OCCURRENCE; uint16 wraps."""
    f0 = np.array([1, 1, 2], dtype=U64)
    f1 = np.array([1, 1, 1, 3], dtype=U64)
    assert list(oracle.common([f0, f1], 4)[0]) == [1]          # 1 + 3 occurrences
    assert list(oracle.common([f0, f1], 1)[0]) == [1, 2, 3]
    big = np.zeros(65535, dtype=U64)                           # 1 + 65535 = 65536 -> wraps to 0
    assert list(oracle.common([np.array([0], dtype=U64), big], 1)[0]) == []
    assert list(oracle.common([f0, f1], 0)[0]) == [1, 2, 3]    # threshold 0 (small -p) keeps everything


def test_diff_collapses_duplicates_and_keeps_file0_taxid():
    f0 = (np.array([1, 1, 2, 5], dtype=U64), np.array([7, 8, 9, 10], dtype=np.uint32))
    f1 = (np.array([2], dtype=U64), np.array([9], dtype=np.uint32))
    k, t = oracle.diff([f0, f1], has_taxid=True)
    assert list(k) == [1, 5] and list(t) == [8, 10]            # map collapse: last taxid of a duplicate wins (diff.go:450-452)
    # -t: a shared k-mer stays when taxids are equal (diff.go:361-364)
    k, _ = oracle.diff([f0, f1], has_taxid=True, compare_taxid=True)
    assert list(k) == [1, 2, 5]
    # one input file / no stored map -> empty output (B-5)
    assert len(oracle.diff([f0[0]])[0]) == 0


def test_fold_variants_against_python():
    r = rng(3)
    keys = np.sort(r.integers(0, 30, 200).astype(U64))
    vals, cnt = np.unique(keys, return_counts=True)
    assert np.array_equal(oracle.fold(oracle.FOLD_UNIQUE, keys)[0], vals)
    assert np.array_equal(oracle.fold(oracle.FOLD_REPEATED_FINAL, keys)[0], vals[cnt >= 2])
    chunk = np.concatenate([[v] * (2 if c >= 2 else 1) for v, c in zip(vals, cnt)]).astype(U64)
    assert np.array_equal(oracle.fold(oracle.FOLD_REPEATED_CHUNK, keys)[0], chunk)
    # re-folding the chunk output with the final rule gives the repeated set (the two-round merge of sort -m -d)
    assert np.array_equal(oracle.fold(oracle.FOLD_REPEATED_FINAL, chunk)[0], vals[cnt >= 2])


def test_lca_edge_cases():
    #        1
    #      2   3
    #     4 5   6
    parent = np.array([0, 1, 1, 1, 2, 2, 3, 0], dtype=np.uint32)  # 7 unknown
    t = oracle.Taxonomy(parent, [9], [4])                        # merged 9 -> 4
    assert t.lca(4, 5) == 2 and t.lca(4, 6) == 1 and t.lca(4, 2) == 2 and t.lca(6, 6) == 6
    assert t.lca(0, 4) == 0 and t.lca(4, 0) == 0                  # 0 is absorbing
    assert t.lca(7, 4) == 0 and t.lca(7, 7) == 7                  # unknown -> 0, but a == b is returned unchanged
    assert t.lca(9, 5) == 2 and t.lca(9, 4) == 4                  # merged ids are remapped
    assert t.lca(100, 4) == 0


def test_iterators_match_naive_definitions():
    r = rng(5)
    seq = r.choice(np.frombuffer(b"ACGT", dtype=np.uint8), 300).tobytes()
    code = {c: i for i, c in enumerate(b"ACGT")}
    for k in (1, 5, 31, 32):
        naive = []
        for i in range(len(seq) - k + 1):
            v = 0
            for c in seq[i:i + k]:
                v = v * 4 + code[c]
            rc = 0
            for c in reversed(seq[i:i + k]):
                rc = rc * 4 + (3 - code[c])
            naive.append(min(v, rc))
        assert [int(x) for x in oracle.kmer_iter(seq, k, canonical=True)] == naive
    # circular: len(seq) k-mers over seq + seq[:k-1]
    assert np.array_equal(oracle.kmer_iter(seq, 7, True, True), oracle.kmer_iter(seq + seq[:6], 7, True, False))
    assert np.array_equal(oracle.nthash_iter(seq, 9, True, True), oracle.nthash_iter(seq + seq[:8], 9, True, False))
    # rolling ntHash == from-scratch ntHash of every window
    scratch = np.array([int(oracle.nthash_iter(seq[i:i + 21], 21, True)[0]) for i in range(len(seq) - 20)], dtype=U64)
    assert np.array_equal(oracle.nthash_iter(seq, 21, True), scratch)
    assert len(oracle.kmer_iter(b"ACG", 5)) == 0                   # ErrShortSeq -> record skipped


def test_parallel_sort_matches_numpy():
    keys = rng(9).integers(0, 2**64, 500_000, dtype=U64)
    assert np.array_equal(oracle.sort_u64(keys, threads=4), np.sort(keys))
    assert np.array_equal(oracle.sort_u64(keys >> U64(40), threads=3), np.sort(keys >> U64(40)))
