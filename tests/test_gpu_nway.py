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


@pytest.mark.parametrize("cfg", ["rows", "0", "1", "2", "3", "4"])
@pytest.mark.parametrize("nf", [2, 3, 4, 5, 7, 8])
def test_nway_union_shapes(eng, cfg, nf, monkeypatch):
    """cfg "rows": the row-based kernel (nunion.cu, opt-in); "0".."4": the tile shapes of the sequential-walk kernel."""
    if cfg == "rows":
        monkeypatch.setenv("UKM_NUNION", "1")
    else:
        monkeypatch.setenv("UKM_NWAY_CFG", cfg)
    monkeypatch.setenv("UKM_NWAY_FORCE", "1")  # two sets go to the two-way pipeline by default
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


@pytest.mark.parametrize("kernel", ["rows", "walk"])
def test_nway_union_distributions(eng, kernel, monkeypatch):
    if kernel == "rows":
        monkeypatch.setenv("UKM_NUNION", "1")
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


@pytest.mark.parametrize("kernel", ["rows", "walk"])
def test_nway_union_misaligned_device_pointers(eng, kernel, monkeypatch):
    import torch
    if kernel == "rows":
        monkeypatch.setenv("UKM_NUNION", "1")
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


# ---------------------------------------------------------------------------------------
# N-way inter / diff (hash filter per tile; replaces inter.go:205-267 / diff.go:380-435 file by file)
# ---------------------------------------------------------------------------------------
def exp_inter(files):
    r = np.asarray(files[0], dtype=U64)
    for f in files[1:]:
        r = r[np.isin(r, f)]
    return r


def exp_diff(files):
    r = np.asarray(files[0], dtype=U64)
    for f in files[1:]:
        r = r[~np.isin(r, f)]
    return r


@pytest.mark.parametrize("cfg", ["0", "1", "2", "3", "4"])
@pytest.mark.parametrize("nf", [3, 4, 5, 8])
def test_nway_filter_shapes(eng, cfg, nf, monkeypatch):
    monkeypatch.setenv("UKM_NWAY_CFG", cfg)
    monkeypatch.setenv("UKM_NWAY_FORCE", "1")
    for N in (3_000, 250_000, 2_500_000):
        files = member_files(N, nf)
        same(eng.inter(files)[0], oracle.inter(files)[0], f"inter cfg {cfg} nf {nf} N {N}")
        same(eng.diff(files)[0], oracle.diff(files)[0], f"diff cfg {cfg} nf {nf} N {N}")


@pytest.mark.parametrize("nf", [9, 15, 16, 20])
def test_nway_filter_more_than_eight_files(eng, nf, monkeypatch):
    monkeypatch.setenv("UKM_NWAY_FORCE", "1")
    r = rng(nf)
    uni = np.unique(r.integers(0, 2**62, 300_000, dtype=U64))
    files = [uni[r.random(len(uni)) < 0.9] for _ in range(nf)]
    same(eng.inter(files)[0], exp_inter(files), f"inter of {nf} files")
    files = [uni[r.random(len(uni)) < 0.5]] + [uni[r.random(len(uni)) < 0.1] for _ in range(nf - 1)]
    same(eng.diff(files)[0], exp_diff(files), f"diff of {nf} files")


def test_nway_filter_matches_file_by_file(eng, monkeypatch):
    monkeypatch.setenv("UKM_NWAY_FILTER", "1")
    files = member_files(1_000_000, 8)
    a_i, a_d = eng.inter(files)[0], eng.diff(files)[0]
    monkeypatch.setenv("UKM_NWAY", "0")
    same(a_i, eng.inter(files)[0], "inter: nway vs file by file")
    same(a_d, eng.diff(files)[0], "diff: nway vs file by file")


def test_nway_filter_distributions(eng, monkeypatch):
    monkeypatch.setenv("UKM_NWAY_FORCE", "1")
    r = rng(21)
    cases = {}
    base = np.unique(r.integers(0, 2**64, 300_000, dtype=U64))
    fs = [base[r.random(len(base)) < 0.7] for _ in range(8)]
    for q in (0, 3, 7):
        fs[q] = np.unique(np.concatenate([fs[q], np.array([0, 2**64 - 1], dtype=U64)]))
    cases["extremes"] = fs
    cases["identical"] = [base.copy() for _ in range(5)]
    # file 0 dominates some tiles (dense runs of consecutive integers): the binary-search path of a tile
    big0 = np.unique(np.concatenate([r.integers(0, 2**63, 20_000, dtype=U64), U64(10**12) + np.arange(300_000, dtype=U64)]))
    cases["file0_dense"] = [big0] + [np.unique(np.concatenate([r.integers(0, 2**63, 20_000, dtype=U64),
                                                               U64(10**12) + np.arange(0, 300_000, 3 + q, dtype=U64)])) for q in range(4)]
    # one tile holding key 0 and key 2^64-1 of file 0: no value left for "empty"
    cases["tiny_extremes"] = [np.array([0, 5, 9, 2**64 - 1], dtype=U64), np.array([0, 9, 2**64 - 1], dtype=U64),
                              np.array([5, 9, 2**64 - 1], dtype=U64)]
    cases["disjoint"] = [np.unique(r.integers(q << 50, (q << 50) + 2**30, 5000 << (q % 5), dtype=U64)) for q in range(6)]
    cases["geometric"] = [np.unique(np.exp2(r.random(150_000) * 40 + 20).astype(U64)) for _ in range(8)]
    cases["small_first"] = [base[::50].copy()] + [base[r.random(len(base)) < 0.8] for _ in range(4)]
    cases["with_empty_subject"] = [base[::2].copy(), base[::3].copy(), np.zeros(0, dtype=U64), base[::5].copy(), base[::7].copy()]
    for name, files in cases.items():
        same(eng.inter(files)[0], oracle.inter(files)[0], f"inter {name}")
        # an EMPTY subject: the reference's worker leaves its loop there (quirk B-5, diff.go:387-392, depends on -j and
        # on scheduling); the library skips the empty file and goes on, with and without the N-way pass
        exp = exp_diff(files) if name == "with_empty_subject" else oracle.diff(files)[0]
        same(eng.diff(files)[0], exp, f"diff {name}")


def test_nway_filter_misaligned_device_pointers(eng, monkeypatch):
    import torch
    monkeypatch.setenv("UKM_NWAY_FORCE", "1")
    files = member_files(700_000, 6)
    d = [torch.from_numpy(f.view(np.int64)).cuda()[(i % 2):] for i, f in enumerate(files)]
    h = [x.cpu().numpy().view(U64) for x in d]
    same(eng.inter(d)[0].cpu().numpy().view(U64), exp_inter(h), "misaligned inter")
    same(eng.diff(d)[0].cpu().numpy().view(U64), exp_diff(h), "misaligned diff")


# ---------------------------------------------------------------------------------------
# single-pass inter / diff over file-0 chunks (unikmer_b200/csrc/nfilter.cu, the default keys-only path)
# ---------------------------------------------------------------------------------------
@pytest.mark.parametrize("cfg", ["0", "1", "2"])
@pytest.mark.parametrize("sub", ["0", "64", "32", "16", "8"])
@pytest.mark.parametrize("nf", [2, 3, 5, 8])
def test_nfilter_shapes(eng, cfg, sub, nf, monkeypatch):
    """Every tile shape (file-0 keys per warp x slot size); a forced shape that is too large for the size ratio
    sends segments down the global-memory look-up path of a tile."""
    monkeypatch.setenv("UKM_NFILTER_CFG", cfg)
    monkeypatch.setenv("UKM_NFILTER_SUB", sub)
    monkeypatch.setenv("UKM_NFILTER_FORCE", "1")
    for N in (3_000, 250_000, 2_500_000):
        files = member_files(N, nf)
        same(eng.inter(files)[0], oracle.inter(files)[0], f"inter cfg {cfg} sub {sub} nf {nf} N {N}")
        same(eng.diff(files)[0], oracle.diff(files)[0], f"diff cfg {cfg} sub {sub} nf {nf} N {N}")


def test_nfilter_is_the_default_path(eng):
    files = member_files(600_000, 8)
    eng.stats_reset()
    eng.stats_enable(True)
    gi, gd = eng.inter(files)[0], eng.diff(files)[0]
    eng.stats_enable(False)
    st = eng.stats()
    assert st.get("setop_inter_nway", {}).get("launches") == 1 and st.get("setop_diff_nway", {}).get("launches") == 1, st
    same(gi, oracle.inter(files)[0], "inter default")
    same(gd, oracle.diff(files)[0], "diff default")


@pytest.mark.parametrize("nf", [9, 15, 16, 20])
def test_nfilter_more_than_eight_files(eng, nf, monkeypatch):
    monkeypatch.setenv("UKM_NFILTER_FORCE", "1")
    r = rng(nf)
    uni = np.unique(r.integers(0, 2**62, 300_000, dtype=U64))
    files = [uni[r.random(len(uni)) < 0.9] for _ in range(nf)]
    same(eng.inter(files)[0], exp_inter(files), f"inter of {nf} files")
    files = [uni[r.random(len(uni)) < 0.5]] + [uni[r.random(len(uni)) < 0.1] for _ in range(nf - 1)]
    same(eng.diff(files)[0], exp_diff(files), f"diff of {nf} files")


@pytest.mark.parametrize("sub", ["0", "64", "8"])
def test_nfilter_distributions(eng, sub, monkeypatch):
    monkeypatch.setenv("UKM_NFILTER_FORCE", "1")
    monkeypatch.setenv("UKM_NFILTER_SUB", sub)
    r = rng(21)
    cases = {}
    base = np.unique(r.integers(0, 2**64, 300_000, dtype=U64))
    fs = [base[r.random(len(base)) < 0.7] for _ in range(8)]
    for q in (0, 3, 7):
        fs[q] = np.unique(np.concatenate([fs[q], np.array([0, 2**64 - 1], dtype=U64)]))
    cases["extremes"] = fs
    cases["identical"] = [base.copy() for _ in range(5)]
    big0 = np.unique(np.concatenate([r.integers(0, 2**63, 20_000, dtype=U64), U64(10**12) + np.arange(300_000, dtype=U64)]))
    cases["file0_dense"] = [big0] + [np.unique(np.concatenate([r.integers(0, 2**63, 20_000, dtype=U64),
                                                               U64(10**12) + np.arange(0, 300_000, 3 + q, dtype=U64)])) for q in range(4)]
    # subjects locally far denser than file 0 (runs of consecutive integers where file 0 has a handful of keys):
    # those segments do not fit a slot and are probed in global memory
    sparse0 = np.unique(np.concatenate([r.integers(0, 2**63, 30_000, dtype=U64), U64(10**12) + np.arange(0, 400_000, 997, dtype=U64)]))
    cases["subjects_dense"] = [sparse0] + [np.unique(np.concatenate([r.integers(0, 2**63, 30_000, dtype=U64),
                                                                     U64(10**12) + np.arange(q, 400_000, 1 + q, dtype=U64)])) for q in range(5)]
    cases["tiny_extremes"] = [np.array([0, 5, 9, 2**64 - 1], dtype=U64), np.array([0, 9, 2**64 - 1], dtype=U64),
                              np.array([5, 9, 2**64 - 1], dtype=U64)]
    cases["one_key"] = [np.array([7], dtype=U64), np.array([7], dtype=U64), np.array([1, 7, 9], dtype=U64)]
    cases["disjoint"] = [np.unique(r.integers(q << 50, (q << 50) + 2**30, 5000 << (q % 5), dtype=U64)) for q in range(6)]
    cases["geometric"] = [np.unique(np.exp2(r.random(150_000) * 40 + 20).astype(U64)) for _ in range(8)]
    cases["small_first"] = [base[::50].copy()] + [base[r.random(len(base)) < 0.8] for _ in range(4)]
    cases["big_first"] = [base.copy()] + [base[::40 + q].copy() for q in range(4)]
    cases["subject_below_and_above"] = [base[1000:2000].copy(), base.copy(), base[::2].copy()]
    cases["with_empty_subject"] = [base[::2].copy(), base[::3].copy(), np.zeros(0, dtype=U64), base[::5].copy(), base[::7].copy()]
    for name, files in cases.items():
        same(eng.inter(files)[0], oracle.inter(files)[0], f"inter {name} sub {sub}")
        exp = exp_diff(files) if name == "with_empty_subject" else oracle.diff(files)[0]
        same(eng.diff(files)[0], exp, f"diff {name} sub {sub}")


def test_nfilter_tile_and_mask_block_seams(eng, monkeypatch):
    """File-0 lengths around multiples of the warp group (64), the tile (512) and the gather block (2048 masks)."""
    monkeypatch.setenv("UKM_NFILTER_FORCE", "1")
    r = rng(5)
    base = np.unique(r.integers(0, 2**62, 600_000, dtype=U64))
    for n0 in (1, 63, 64, 65, 511, 512, 513, 4095, 4096, 4097, 64 * 2048 - 1, 64 * 2048, 64 * 2048 + 1, 300_001):
        files = [base[:n0].copy()] + [base[r.random(len(base)) < 0.6] for _ in range(3)]
        same(eng.inter(files)[0], exp_inter(files), f"inter n0 {n0}")
        same(eng.diff(files)[0], exp_diff(files), f"diff n0 {n0}")


def test_nfilter_misaligned_device_pointers(eng, monkeypatch):
    import torch
    monkeypatch.setenv("UKM_NFILTER_FORCE", "1")
    files = member_files(700_000, 6)
    d = [torch.from_numpy(f.view(np.int64)).cuda()[(i % 2):] for i, f in enumerate(files)]
    h = [x.cpu().numpy().view(U64) for x in d]
    same(eng.inter(d)[0].cpu().numpy().view(U64), exp_inter(h), "misaligned inter")
    same(eng.diff(d)[0].cpu().numpy().view(U64), exp_diff(h), "misaligned diff")
