"""Pins the CPU oracle (oracle/oracle.c) to every known answer the reference holds for
the hot path: K1..K9 of SURVEY.md section 4 (README.md:156-278,
analysis/distance/README.md:9).  The reference ships no *_test.go, so these
documented answers + its three genomes are the whole pinning set."""
import json
import os

import numpy as np
import pytest

import oracle
from tests.golden.make_golden import GENOMES, REF, digest, read_fasta_records, unpack2


@pytest.fixture(scope="module")
def kat(golden_dir):
    with open(os.path.join(golden_dir, "kat.json")) as fh:
        return json.load(fh)


@pytest.fixture(scope="module")
def genomes(golden_dir):
    """MG1655 + IAI39 from the committed 2-bit fixture (made by make_golden.py)."""
    z = np.load(os.path.join(golden_dir, "genomes.npz"))
    return {n: unpack2(z[n], int(z[n + "_len"])) for n in ("mg1655", "iai39")}


def _count(seq, k, **kw):
    return oracle.count(seq, np.array([0, len(seq)], dtype=np.uint64), k, **kw)


@pytest.fixture(scope="module")
def sets23(genomes):
    return {n: _count(s, 23, canonical=True) for n, s in genomes.items()}


def test_k1_k2_unique_canonical_23mers(kat, sets23):
    assert len(sets23["mg1655"]) == kat["K1_mg1655_k23_canonical_unique"] == 4546632
    assert len(sets23["iai39"]) == kat["K2_iai39_k23_canonical_unique"] == 4902266


def test_k4_k5_k6_set_cardinalities(kat, sets23):
    a, b = sets23["iai39"], sets23["mg1655"]
    u, _ = oracle.union([a, b])
    i, _ = oracle.inter([a, b])
    d, _ = oracle.diff([a, b])
    assert (len(u), len(i), len(d)) == (6872728, 2576170, 2326096)
    assert digest(u) == kat["digests"]["union"]
    assert digest(i) == kat["digests"]["inter"]
    assert digest(d) == kat["digests"]["diff"]
    # sort -d over both files gives the same "dup" count as inter (README.md:234,271)
    k = oracle.sort_u64(np.concatenate([a, b]))
    rep, _ = oracle.fold(oracle.FOLD_REPEATED_FINAL, k)
    assert np.array_equal(rep, i)
    # union -s == sort -u (K10, README.md:222-229)
    uq, _ = oracle.fold(oracle.FOLD_UNIQUE, k)
    assert np.array_equal(uq, u)


def test_k7_first_three_sorted(kat, sets23):
    got = [oracle.decode(int(c), 23).decode() for c in sets23["mg1655"][:3]]
    assert got == kat["K7_first3_sorted_mg1655"]


def test_k8_nthash_known_values(kat):
    for kmer, h in kat["K8_nthash_k23_canonical"].items():
        assert int(oracle.nthash_iter(kmer.encode(), 23, canonical=True)[0]) == h


def test_k9_scaled_minhash_whole_genome(kat, genomes):
    mh = kat["max_hash_scale15"]
    assert mh == int(float(2**64 - 1) / 15.0)
    sc = _count(genomes["mg1655"], 31, canonical=True, hashed=True, scaled=True, max_hash=mh)
    assert len(sc) == kat["K9_mg1655_k31_nthash_scaled15"] == 586734


def test_k12_minimizer_whole_genome(kat, genomes):
    """analysis/distance/README.md:8,15,39: `count -k 31 -K -H -W 15` on MG1655 holds 549 963 k-mers."""
    mg = genomes["mg1655"]
    off = np.array([0, len(mg)], dtype=np.uint64)
    mz = oracle.count_minimizer(mg, off, 31, 15, canonical=True)
    assert len(mz) == kat["K12_mg1655_k31_minimizer_w15"] == 549963
    assert digest(mz) == kat["digests"]["mg1655_k31_minimizer_w15"]


def test_minimizer_semantics_small():
    """w = 1 is the plain hashed count; windows never span records; a record with fewer than w k-mers gives nothing."""
    r = np.random.default_rng(5)
    recs = [r.choice(np.frombuffer(b"ACGT", dtype=np.uint8), L).astype(np.uint8) for L in (0, 40, 31, 45, 2000, 33)]
    bases = np.concatenate(recs)
    off = np.concatenate([[0], np.cumsum([len(x) for x in recs])]).astype(np.uint64)
    k = 31
    assert np.array_equal(oracle.count_minimizer(bases, off, k, 1), oracle.count(bases, off, k, canonical=True, hashed=True))
    for w in (2, 5, 15, 100):
        exp = set()
        for s in recs:
            h = oracle.nthash_iter(s, k, canonical=True) if len(s) >= k else np.zeros(0, dtype=np.uint64)
            for i in range(0, len(h) - w + 1):
                exp.add(int(h[i:i + w].min()))
        got = oracle.count_minimizer(bases, off, k, w)
        assert np.array_equal(got, np.array(sorted(exp), dtype=np.uint64)), w
    mh = int(float(2**64 - 1) / 3.0)
    full = oracle.count_minimizer(bases, off, k, 5)
    assert np.array_equal(oracle.count_minimizer(bases, off, k, 5, scaled=True, max_hash=mh), full[full <= np.uint64(mh)])


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference testdata not present on this box")
def test_k3_and_fixture_matches_reference_files(kat, genomes):
    """Against the reference's own files: K3 (third genome) and that the committed 2-bit
    fixture is exactly the reference genomes."""
    for name, fn in GENOMES.items():
        recs = read_fasta_records(os.path.join(REF, fn))
        assert len(recs) == 1
        if name in genomes:
            assert recs[0] == genomes[name]
        else:
            assert len(_count(recs[0], 23, canonical=True)) == kat["K3_amuc_k23_canonical_unique"] == 2630905
