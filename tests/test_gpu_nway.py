"""Parity of the single-pass N-way union (unikmer_b200/csrc/nway.cu; replaces union.go:186-208,260-305)
against the CPU oracle: every tile shape, 2..20 files (fan-in 8 per pass), misaligned device pointers,
key distributions that stress the partition, and the two-way fallback for inputs that cannot be tiled."""
import numpy as np
import pytest

import oracle
from tests.test_gpu_parity import U64, member_files, rng, same

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def eng():
    from unikmer_b200 import Engine
    e = Engine(0)
    yield e
    e.close()


def exp_union(files):
    return np.unique(np.concatenate([np.asarray(f, dtype=U64) for f in files]))


@pytest.mark.parametrize("cfg", ["0", "1", "2", "3", "4"])
@pytest.mark.parametrize("nf", [2, 3, 4, 5, 7, 8])
def test_nway_union_shapes(eng, cfg, nf, monkeypatch):
    monkeypatch.setenv("UKM_NWAY_CFG", cfg)
    for N in (3_000, 250_000, 2_500_000):
        files = member_files(N, nf)
        same(eng.union(files)[0], oracle.union(files)[0], f"union cfg {cfg} nf {nf} N {N}")


@pytest.mark.parametrize("nf", [9, 16, 20])
def test_nway_union_more_than_eight_files(eng, nf):
    r = rng(nf)
    uni = np.unique(r.integers(0, 2**62, 400_000, dtype=U64))
    files = [uni[r.random(len(uni)) < 0.3] for _ in range(nf)]
    same(eng.union(files)[0], exp_union(files), f"union of {nf} files")


def test_nway_union_matches_two_way_tree(eng, monkeypatch):
    files = member_files(1_000_000, 8)
    a = eng.union(files)[0]
    monkeypatch.setenv("UKM_NWAY", "0")
    b = eng.union(files)[0]
    same(a, b, "nway vs two-way tree")
    same(a, oracle.union(files)[0], "nway vs oracle")


def test_nway_union_distributions(eng):
    r = rng(11)
    cases = {}
    # full 64-bit range with both extremes present in several files
    base = np.unique(r.integers(0, 2**64, 300_000, dtype=U64))
    fs = [base[r.random(len(base)) < 0.6] for _ in range(8)]
    fs[0] = np.unique(np.concatenate([fs[0], np.array([0, 2**64 - 1], dtype=U64)]))
    fs[5] = np.unique(np.concatenate([fs[5], np.array([0, 2**64 - 1], dtype=U64)]))
    cases["extremes"] = fs
    # identical files
    cases["identical"] = [base.copy() for _ in range(8)]
    # disjoint key ranges, very different sizes, empty files in between
    cases["disjoint"] = [np.unique(r.integers(f << 50, (f << 50) + 2**30, (0 if f % 3 == 2 else 1000 << f), dtype=U64))
                         for f in range(8)]
    # dense runs of consecutive integers over a sparse background
    cl = []
    for f in range(8):
        c = U64(1_000_000_007 * (f % 3 + 1))
        cl.append(np.unique(np.concatenate([r.integers(0, 2**63, 30_000, dtype=U64),
                                            c + np.arange(200_000, dtype=U64) * U64(f % 2 + 1)])))
    cases["clustered"] = cl
    # density varying over 40 binades
    cases["geometric"] = [np.unique(np.exp2(r.random(150_000) * 40 + 20).astype(U64)) for _ in range(8)]
    # one big file, seven tiny ones
    cases["skewed"] = [np.unique(r.integers(0, 2**62, 2_000_001, dtype=U64))] + \
                      [np.unique(r.integers(0, 2**62, 97 + 13 * f, dtype=U64)) for f in range(7)]
    cases["tiny"] = [np.array([5], dtype=U64), np.array([5, 7], dtype=U64), np.array([1, 5, 9], dtype=U64)]
    for name, files in cases.items():
        same(eng.union(files)[0], exp_union(files), f"union {name}")


def test_nway_union_misaligned_device_pointers(eng):
    import torch
    files = member_files(700_000, 8)
    d = [torch.from_numpy(f.view(np.int64)).cuda()[(i % 2):] for i, f in enumerate(files)]  # 8 mod 16 pointers on odd files
    exp = exp_union([x.cpu().numpy().view(U64) for x in d])
    got = eng.union(d)[0]
    same(got.cpu().numpy().view(U64), exp, "misaligned union")


def test_nway_union_falls_back_on_inputs_that_cannot_be_tiled(eng, monkeypatch):
    """One key repeated far beyond a tile (not duplicate-free: outside the contract) must never be mis-merged: the
    partition refuses the input and the two-way tree takes over, which behaves as it always did -- it either reports
    UKM_E_NOT_SORTED_UNIQUE (a tile seam fell inside the run of duplicates) or returns the merged keys."""
    import unikmer_b200 as ub
    r = rng(3)
    files = []
    for f in range(8):
        a = np.unique(r.integers(0, 2**40, 20_000, dtype=U64))
        files.append(np.concatenate([a, np.full(5_000, 2**41, dtype=U64)]))

    def run():
        try:
            return eng.union(files)[0]
        except ub.UkmError as e:
            assert e.status == ub.E_NOT_SORTED_UNIQUE
            return None
    got = run()
    monkeypatch.setenv("UKM_NWAY", "0")
    old = run()
    assert (got is None) == (old is None)
    if got is not None:
        same(got, old, "fallback result = two-way tree result")
        assert np.array_equal(np.unique(got), exp_union(files))
