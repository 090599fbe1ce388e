"""Parity of the CUDA path (through the C ABI) against the CPU oracle.  Bit-exact: every
quantity on this path is an integer (uint64 codes, uint32 taxids, uint16 counts)."""
import json
import os

import numpy as np
import pytest

import oracle
from tests.golden.make_golden import digest, unpack2

pytestmark = pytest.mark.gpu

U64 = np.uint64


@pytest.fixture(scope="module")
def eng():
    from unikmer_b200 import Engine
    e = Engine(0)
    yield e
    e.close()


def same(got, exp, what=""):
    got = np.asarray(got)
    exp = np.asarray(exp)
    if got.shape != exp.shape:
        n = min(len(got), len(exp))
        first = int(np.argmax(got[:n] != exp[:n])) if n and (got[:n] != exp[:n]).any() else n
        raise AssertionError(f"{what}: length {len(got)} != {len(exp)}; first difference at {first}: "
                             f"got {got[first:first + 4]} exp {exp[first:first + 4]}")
    if not np.array_equal(got, exp):
        bad = np.nonzero(got != exp)[0]
        i = int(bad[0])
        raise AssertionError(f"{what}: {len(bad)} of {len(exp)} differ; first at {i}: got {got[max(0, i - 1):i + 3]} "
                             f"exp {exp[max(0, i - 1):i + 3]}")


def rng(seed):
    return np.random.default_rng(seed)


def synth_tax(n_nodes=10_000, Q=8):
    """parent[t] = 1 + sm64(Q+t) % (t-1) for t >= 2, root 1."""
    parent = np.zeros(n_nodes + 1, dtype=np.uint32)
    parent[1] = 1
    for t in range(2, n_nodes + 1):
        parent[t] = 1 + oracle.sm64(Q + t) % (t - 1)
    return parent


@pytest.fixture(scope="module")
def tax(eng):
    parent = synth_tax()
    # a few merged ids (old -> new) and holes (unknown ids) for the edge cases
    parent[777] = 0
    parent[4242] = 0
    # children of removed nodes must not dangle: re-parent them to the root
    for t in range(2, len(parent)):
        if parent[t] in (777, 4242):
            parent[t] = 1
    mf = np.array([777, 20_001], dtype=np.uint32)
    mt = np.array([778, 5], dtype=np.uint32)
    eng.set_taxonomy(parent, mf, mt)
    return oracle.Taxonomy(parent, mf, mt), len(parent)


# ---------------------------------------------------------------------------------------
# a12: LCA
# ---------------------------------------------------------------------------------------
def test_lca_matches_oracle(eng, tax):
    otax, n = tax
    r = rng(1)
    a = r.integers(0, n + 50, 20_000).astype(np.uint32)
    b = r.integers(0, n + 50, 20_000).astype(np.uint32)
    a[:10] = [0, 5, 777, 4242, 20_001, 1, 9999, 10_000, 123, 123]
    b[:10] = [5, 0, 778, 7, 5, 9, 9999, 1, 123, 0]
    exp = np.array([otax.lca(int(x), int(y)) for x, y in zip(a, b)], dtype=np.uint32)
    same(eng.lca(a, b), exp, "lca")


# ---------------------------------------------------------------------------------------
# a4/a5: sort
# ---------------------------------------------------------------------------------------
@pytest.mark.parametrize("n", [0, 1, 2, 3, 255, 256, 257, 4095, 4096, 4097, 100_003, 1_000_000, 3_333_333])
def test_sort_u64_sizes(eng, n):
    keys = oracle.random_keys(0, n, 2)
    got, _ = eng.sort(keys.copy(), key_bits=62)
    same(got, np.sort(keys), f"sort n={n}")


@pytest.mark.parametrize("match", ["any", "ballot"])
@pytest.mark.parametrize("cfg", ["0", "1", "2", "3"])
def test_sort_tile_configs(eng, cfg, match, monkeypatch):
    monkeypatch.setenv("UKM_SORT_CFG", cfg)
    monkeypatch.setenv("UKM_SORT_MATCH", match)
    keys = rng(int(cfg)).integers(0, 2**64, 700_001, dtype=U64)
    got, _ = eng.sort(keys.copy(), key_bits=64)
    same(got, np.sort(keys), f"sort cfg={cfg}")
    t = rng(7).integers(0, 2**32, len(keys), dtype=np.uint32)
    gk, gt = eng.sort(keys.copy(), t.copy(), key_bits=64)
    order = np.argsort(keys, kind="stable")
    same(gk, keys[order], f"sort_pairs keys cfg={cfg}")
    same(gt, t[order], f"sort_pairs taxids cfg={cfg}")


def test_sort_adversarial(eng):
    n = 500_000
    for name, keys in {
        "all_equal": np.full(n, 12345, dtype=U64),
        "sorted": np.arange(n, dtype=U64) * U64(977),
        "reversed": (np.arange(n, dtype=U64) * U64(977))[::-1].copy(),
        "two_values": rng(3).integers(0, 2, n).astype(U64) * U64(2**63),
        "max_keys": np.concatenate([np.full(1000, 2**64 - 1, dtype=U64), rng(4).integers(0, 2**64, n, dtype=U64)]),
        "low_byte_only": rng(5).integers(0, 256, n).astype(U64),
    }.items():
        got, _ = eng.sort(keys.copy(), key_bits=64)
        same(got, np.sort(keys), f"sort {name}")


def test_sort_pairs_is_stable_and_matches_oracle(eng):
    r = rng(11)
    keys = r.integers(0, 5000, 300_000).astype(U64)  # many ties
    tx = r.integers(1, 1000, len(keys)).astype(np.uint32)
    gk, gt = eng.sort(keys.copy(), tx.copy(), key_bits=16)
    ok, ot = oracle.sort_pairs(keys, tx)
    same(gk, ok, "pairs keys")
    same(gt, ot, "pairs taxids (stable)")


def test_sort_codetaxid16_aos(eng):
    r = rng(12)
    rec = np.zeros(200_000, dtype=[("code", "<u8"), ("taxid", "<u4"), ("pad", "<u4")])
    rec["code"] = r.integers(0, 2**62, len(rec), dtype=U64)
    rec["taxid"] = r.integers(0, 2**32, len(rec), dtype=np.uint32)
    exp = np.sort(rec, order="code", kind="stable")
    got = eng.sort_codetaxid16(rec.copy(), key_bits=62)
    same(got["code"], exp["code"], "aos code")
    same(got["taxid"], exp["taxid"], "aos taxid")


def test_sort_device_resident(eng):
    import torch
    keys = oracle.random_keys(0, 2_000_001, 2)
    d = torch.from_numpy(keys.view(np.int64)).cuda()
    eng.sort(d, key_bits=62)
    same(d.cpu().numpy().view(U64), np.sort(keys), "device sort")
    # unaligned (8 mod 16) device pointer: sort a view that starts one element in
    d2 = torch.from_numpy(keys.view(np.int64)).cuda()
    v = d2[1:]
    assert v.data_ptr() % 16 == 8 and v.is_contiguous()
    eng.sort(v, key_bits=62)
    same(v.cpu().numpy().view(U64), np.sort(keys[1:]), "unaligned device sort")
    assert int(d2[0].item()) == int(keys[0].view(np.int64))


# ---------------------------------------------------------------------------------------
# a6: folds
# ---------------------------------------------------------------------------------------
@pytest.mark.parametrize("mode", [oracle.FOLD_PLAIN, oracle.FOLD_UNIQUE, oracle.FOLD_REPEATED_FINAL, oracle.FOLD_REPEATED_CHUNK])
@pytest.mark.parametrize("with_tax", [False, True])
def test_fold_modes(eng, tax, mode, with_tax):
    otax, n_tax = tax
    r = rng(20 + mode)
    for n, hi in [(0, 10), (1, 10), (2, 2), (5000, 1500), (200_000, 60_000), (100_000, 3)]:
        keys = np.sort(r.integers(0, hi, n).astype(U64) * U64(1_000_003))
        tx = r.integers(1, n_tax, n).astype(np.uint32) if with_tax else None
        ek, et = oracle.fold(mode, keys, tx, otax)
        gk, gt = eng.fold(mode, keys, tx)
        same(gk, ek, f"fold mode={mode} tax={with_tax} n={n} keys")
        if with_tax:
            same(gt, et, f"fold mode={mode} n={n} taxids")


def test_fold_sentinel_quirks(eng, tax):
    """B-1 / B-2: the `last = ^uint64(0)` sentinel (sort.go:505-507, 542-549)."""
    otax, _ = tax
    top = np.full(3, 2**64 - 1, dtype=U64)
    same(eng.fold(oracle.FOLD_UNIQUE, top)[0], oracle.fold(oracle.FOLD_UNIQUE, top)[0], "B-2")
    assert len(eng.fold(oracle.FOLD_UNIQUE, top)[0]) == 0
    empty = np.zeros(0, dtype=U64)
    ek, et = oracle.fold(oracle.FOLD_UNIQUE, empty, np.zeros(0, dtype=np.uint32), otax)
    gk, gt = eng.fold(oracle.FOLD_UNIQUE, empty, np.zeros(0, dtype=np.uint32))
    same(gk, ek, "B-1 keys")
    same(gt, et, "B-1 taxids")


# ---------------------------------------------------------------------------------------
# a8-a11: set operations
# ---------------------------------------------------------------------------------------
def member_files(N, nfiles, S=3, T=4):
    return [oracle.member_file(0, N, N, S, T, f) for f in range(nfiles)]


@pytest.mark.parametrize("kernel", ["pipe:0", "pipe:1", "pipe:2", "pipe:3", "pipe:4", "pipe:5", "vt:11", "vt:15", "vt:19", "vt:23"])
@pytest.mark.parametrize("skew", ["0", "3"])
def test_setop_kernel_variants(eng, kernel, skew, monkeypatch):
    """Every shape of the keys-only kernels (persistent pipeline / one tile per CTA), with and without
    the search path for skewed pairs."""
    kind, cfg = kernel.split(":")
    monkeypatch.setenv("UKM_SETOP_PIPE", cfg if kind == "pipe" else "off")
    if kind == "vt":
        monkeypatch.setenv("UKM_SETOP_VT", cfg)
    monkeypatch.setenv("UKM_SETOP_SKEW", skew)
    for N, nf in ((400_000, 8), (3_000_000, 3), (5_000, 2)):
        files = member_files(N, nf)
        same(eng.inter(files)[0], oracle.inter(files)[0], f"inter {N}")
        same(eng.diff(files)[0], oracle.diff(files)[0], f"diff {N}")
        same(eng.union(files)[0], oracle.union(files)[0], f"union {N}")
    chunks = [np.sort(rng(3).integers(0, 90_000, n).astype(U64)) for n in (120_000, 7, 55_555, 2_000_000)]
    same(eng.merge(chunks, oracle.FOLD_PLAIN)[0], oracle.merge_chunks(chunks, oracle.FOLD_PLAIN)[0], "merge")
    import torch
    d = [torch.from_numpy(f.view(np.int64)).cuda()[1:] for f in member_files(1_000_000, 2)]  # 8 mod 16 pointers
    exp = oracle.union([x.cpu().numpy().view(U64) for x in d])[0]
    same(eng.union(d)[0].cpu().numpy().view(U64), exp, "unaligned union")


@pytest.mark.parametrize("mode", ["0", "1", "2"])
def test_search_path_skewed_pairs(eng, tax, mode, monkeypatch):
    """|B| >> |A|: the look-up kernels (inter/diff after a few files), keys-only and with taxids, for every
    look-up mode (bisection, interpolated start + gallop, anchored)."""
    monkeypatch.setenv("UKM_SEARCH_MODE", mode)
    otax, n_tax = tax
    r = rng(77)
    big = np.unique(r.integers(0, 2**62, 3_000_000, dtype=U64))
    for n_small in (1, 5, 1000, 100_000):
        hit = r.choice(big, n_small // 2 + 1, replace=False)
        miss = r.integers(0, 2**62, n_small, dtype=U64)
        small = np.unique(np.concatenate([hit, miss, [big[0], big[-1]]]))
        for files in ([small, big], [small, big, big[::3].copy()]):
            same(eng.inter(files)[0], oracle.inter(files)[0], f"search inter {n_small}")
            same(eng.diff(files)[0], oracle.diff(files)[0], f"search diff {n_small}")
        ts, tb = r.integers(0, n_tax, len(small)).astype(np.uint32), r.integers(0, n_tax, len(big)).astype(np.uint32)
        tf = [(small, ts), (big, tb)]
        for kw in ({"has_taxid": True}, {"mix_taxid": True}):
            ek, et = oracle.inter(tf, tax=otax, **kw)
            gk, gt = eng.inter(tf, **kw)
            same(gk, ek, f"search inter tax keys {kw}")
            same(gt, et, f"search inter tax taxids {kw}")
        ek, et = oracle.diff(tf, has_taxid=True, compare_taxid=True, tax=otax)
        gk, gt = eng.diff(tf, has_taxid=True, compare_taxid=True)
        same(gk, ek, "search diff -t keys")
        same(gt, et, "search diff -t taxids")


@pytest.mark.parametrize("mode", ["0", "1", "2"])
def test_search_path_key_distributions(eng, mode, monkeypatch):
    """The interpolated look-up must not depend on uniform keys: clustered, stepped and extreme-valued subjects,
    queries below / above / between all of them, windows that span everything or nothing."""
    monkeypatch.setenv("UKM_SEARCH_MODE", mode)
    r = rng(78)
    dense = np.arange(5_000_000, 7_000_000, dtype=U64)                       # one long run of consecutive keys
    clustered = np.unique(np.concatenate([r.integers(c, c + 4000, 60_000, dtype=U64)
                                          for c in r.integers(0, 2**62, 40, dtype=U64)]))
    stepped = np.unique(np.concatenate([np.arange(0, 300_000, dtype=U64), (U64(1) << U64(61)) + np.arange(0, 300_000, dtype=U64) * U64(977),
                                        np.array([2**64 - 1, 2**64 - 2, 2**63], dtype=U64)]))
    expo = np.unique((U64(1) << r.integers(0, 62, 400_000).astype(U64)) + r.integers(0, 2**20, 400_000, dtype=U64))
    for big in (dense, clustered, stepped, expo):
        picks = [big[r.integers(0, len(big), 3000)], big[:50], big[-50:],
                 r.integers(0, 2**62, 2000, dtype=U64), big[r.integers(0, len(big), 3000)] + U64(1),
                 np.array([0, 1, 2**64 - 1, 2**63], dtype=U64)]
        small = np.unique(np.concatenate(picks))
        for q in (small, small[:1], small[-1:], small[small < big[0]], small[small > big[-1]], small[::97].copy()):
            if len(q) == 0 or len(q) * 6 > len(big):
                continue
            files = [q, big]
            same(eng.inter(files)[0], np.intersect1d(q, big), f"inter mode {mode}")
            same(eng.diff(files)[0], np.setdiff1d(q, big), f"diff mode {mode}")


@pytest.mark.parametrize("N,nfiles", [(1000, 2), (20_000, 3), (300_000, 8), (2_000_000, 2), (1_500_000, 5)])
def test_setops_no_taxid(eng, N, nfiles):
    files = member_files(N, nfiles)
    same(eng.inter(files)[0], oracle.inter(files)[0], "inter")
    same(eng.diff(files)[0], oracle.diff(files)[0], "diff")
    same(eng.union(files)[0], oracle.union(files)[0], "union")
    for thr in (1, 2, nfiles):
        same(eng.common(files, thr)[0], oracle.common(files, thr)[0], f"common -n {thr}")


def test_survey_generator_self_check(eng):
    """N=2e6, S=3, T=4: inter of all 8 = 8002, diff = 7834, sizes as listed."""
    files = member_files(2_000_000, 8)
    assert [len(f) for f in files] == [1000177, 1000659, 999925, 1000928, 998895, 999660, 1001359, 1000217]
    assert len(eng.inter(files)[0]) == 8002
    assert len(eng.diff(files)[0]) == 7834


def test_setops_edge_shapes(eng):
    a = np.arange(0, 100_000, dtype=U64) * U64(3)
    cases = {
        "identical": [a, a.copy()],
        "disjoint_interleaved": [a, a + U64(1)],
        "a_below_b": [a, a + U64(10**9)],
        "b_below_a": [a + U64(10**9), a],
        "single_vs_many": [np.array([30_000], dtype=U64), a],
        "many_vs_single": [a, np.array([30_000], dtype=U64)],
        "top_keys": [np.array([5, 2**62 - 1, 2**64 - 1], dtype=U64), np.array([0, 5, 2**64 - 1], dtype=U64)],
        "tiny": [np.array([1], dtype=U64), np.array([1], dtype=U64)],
        "tile_seam": [np.arange(0, 3840 * 3, dtype=U64), np.arange(0, 3840 * 3, dtype=U64)],
    }
    for name, files in cases.items():
        same(eng.inter(files)[0], oracle.inter(files)[0], f"inter {name}")
        same(eng.diff(files)[0], oracle.diff(files)[0], f"diff {name}")
        same(eng.union(files)[0], oracle.union(files)[0], f"union {name}")
        same(eng.common(files, 2)[0], oracle.common(files, 2)[0], f"common {name}")


def test_setops_empty_inputs(eng):
    import unikmer_b200 as ub
    a = np.arange(10, dtype=U64)
    e = np.zeros(0, dtype=U64)
    # union / common treat an empty file as nothing
    same(eng.union([a, e])[0], oracle.union([a, e])[0], "union with empty")
    same(eng.union([e, e])[0], oracle.union([e, e])[0], "union of empties")
    same(eng.common([a, e, a], 2)[0], oracle.common([a, e, a], 2)[0], "common with empty")
    # inter: a later EMPTY file stops the loop and keeps the current set (quirk B-3, inter.go:211-215)
    same(eng.inter([a, e, a])[0], oracle.inter([a, e, a])[0], "inter B-3")
    assert len(eng.inter([a, e])[0]) == len(a)
    # inter: an empty FIRST file panics in the reference (mc[0], inter.go:208)
    with pytest.raises(oracle.OracleError):
        oracle.inter([e, a])
    with pytest.raises(ub.UkmError) as ei:
        eng.inter([e, a])
    assert ei.value.status == ub.E_PANIC
    # diff: empty file 0 -> header-only output (diff.go:155-201)
    assert len(eng.diff([e, a])[0]) == 0 == len(oracle.diff([e, a])[0])
    # single input: the command byte-copies the file; the ABI hands the set back unchanged
    same(eng.inter([a])[0], a, "inter single")
    same(eng.union([a])[0], a, "union single")


def test_setops_reject_unsorted_or_duplicate_input(eng):
    import unikmer_b200 as ub
    r = rng(5)
    a = np.sort(r.integers(0, 2**40, 50_000).astype(U64))
    bad = a.copy()
    bad[20_000] = bad[20_001]  # duplicate
    for files in ([bad, a], [a, bad]):
        for op in (eng.inter, eng.diff, eng.union):
            with pytest.raises(ub.UkmError) as ei:
                op(files, validate=True)  # UKM_F_VALIDATE; without it the header flag is trusted (inter.go:139)
            assert ei.value.status == ub.E_NOT_SORTED_UNIQUE
    with pytest.raises(ub.UkmError):
        eng.common([a, bad], 1, validate=True)
    assert eng.check_sorted_unique(a[np.concatenate([[True], a[1:] != a[:-1]])])
    assert not eng.check_sorted_unique(bad)
    # the context stays usable after an error
    same(eng.inter([a[::2].copy(), a[::2].copy()])[0], a[::2], "after error")


@pytest.mark.parametrize("N,nfiles", [(50_000, 2), (400_000, 4), (1_000_000, 8)])
def test_setops_with_taxids(eng, tax, N, nfiles):
    otax, n_tax = tax
    keys = member_files(N, nfiles, S=6, T=7)
    r = rng(N)
    files = [(k, r.integers(0, n_tax, len(k)).astype(np.uint32)) for k in keys]  # includes 0 and unknown ids
    ek, et = oracle.inter(files, has_taxid=True, tax=otax)
    gk, gt = eng.inter(files, has_taxid=True)
    same(gk, ek, "inter+tax keys")
    same(gt, et, "inter+tax taxids")
    ek, et = oracle.inter(files, mix_taxid=True, tax=otax)
    gk, gt = eng.inter(files, mix_taxid=True)
    same(gk, ek, "inter --mix-taxid keys")
    same(gt, et, "inter --mix-taxid taxids")
    ek, et = oracle.union(files, has_taxid=True, tax=otax)
    gk, gt = eng.union(files, has_taxid=True)
    same(gk, ek, "union+tax keys")
    same(gt, et, "union+tax taxids")
    for thr in (1, max(1, nfiles // 2), nfiles):
        ek, et = oracle.common(files, thr, has_taxid=True, tax=otax)
        gk, gt = eng.common(files, thr, has_taxid=True)
        same(gk, ek, f"common -n {thr} keys")
        same(gt, et, f"common -n {thr} taxids")
    ek, et = oracle.diff(files, has_taxid=True, tax=otax)
    gk, gt = eng.diff(files, has_taxid=True)
    same(gk, ek, "diff+tax keys")
    same(gt, et, "diff+tax taxids")
    ek, et = oracle.diff(files, has_taxid=True, compare_taxid=True, tax=otax)
    gk, gt = eng.diff(files, has_taxid=True, compare_taxid=True)
    same(gk, ek, "diff -t keys")
    same(gt, et, "diff -t taxids")


def test_global_taxid_spans(eng, tax):
    """count -t files carry one global taxid (README.md:169-171); ReadCodeWithTaxid broadcasts it."""
    from unikmer_b200 import KmerSet
    otax, _ = tax
    keys = member_files(200_000, 4, S=6, T=7)
    leaf = [9000, 9500, 9990, 1234]
    ofiles = [(k, np.full(len(k), t, dtype=np.uint32)) for k, t in zip(keys, leaf)]
    gsets = [KmerSet(k, None, global_taxid=t) for k, t in zip(keys, leaf)]
    ek, et = oracle.common(ofiles, 2, has_taxid=True, tax=otax)
    gk, gt = eng.common(gsets, 2, has_taxid=True)
    same(gk, ek, "global taxid keys")
    same(gt, et, "global taxid lca")


def test_diff_unsorted_subject(eng):
    files = member_files(300_000, 3)
    r = rng(9)
    shuffled = files[1].copy()
    r.shuffle(shuffled)
    from unikmer_b200 import KmerSet
    exp = oracle.diff([files[0], shuffled, files[2]], sorted_flags=[1, 0, 1])[0]
    got = eng.diff([KmerSet(files[0]), KmerSet(shuffled, sorted=False), KmerSet(files[2])])[0]
    # NOTE: with a sorted subject AFTER an unsorted one the reference walks a stale copy
    # (diff.go:380-435 after 341-367) and resurrects k-mers; the engine subtracts every subject.
    exp_clean = oracle.diff(files)[0]
    same(got, exp_clean, "diff with unsorted subject")
    assert len(exp) >= len(exp_clean)
    got2 = eng.diff([KmerSet(files[0]), KmerSet(files[2]), KmerSet(shuffled, sorted=False)])[0]
    same(got2, oracle.diff([files[0], files[2], shuffled], sorted_flags=[1, 1, 0])[0], "unsorted subject last")


@pytest.mark.parametrize("mode", [oracle.FOLD_PLAIN, oracle.FOLD_UNIQUE, oracle.FOLD_REPEATED_FINAL, oracle.FOLD_REPEATED_CHUNK])
def test_merge_chunks(eng, tax, mode):
    """mergeChunksFile (util-sort.go:227-606): sorted chunks WITH duplicates inside and across chunks."""
    otax, n_tax = tax
    r = rng(30 + mode)
    chunks = [np.sort(r.integers(0, 40_000, n).astype(U64)) for n in (50_000, 1, 0, 77_777, 30_000)]
    same(eng.merge(chunks, mode)[0], oracle.merge_chunks(chunks, mode)[0], f"merge mode={mode}")
    if mode != oracle.FOLD_PLAIN:  # tie order of the plain variant is undefined in the reference (B-10)
        tchunks = [(c, r.integers(1, n_tax, len(c)).astype(np.uint32)) for c in chunks]
        ek, et = oracle.merge_chunks(tchunks, mode, has_taxid=True, tax=otax)
        gk, gt = eng.merge(tchunks, mode, has_taxid=True)
        same(gk, ek, f"merge+tax mode={mode} keys")
        same(gt, et, f"merge+tax mode={mode} taxids")


def test_k10_union_equals_sort_unique(eng):
    """K10 (README.md:222-229): `union -s` and `sort -u` agree."""
    files = member_files(500_000, 6)
    u = eng.union(files)[0]
    s, _ = eng.sort(np.concatenate(files), key_bits=62)
    same(eng.fold(oracle.FOLD_UNIQUE, s)[0], u, "K10")
    # common -n 1 == union, common -n nfiles == inter on duplicate-free inputs
    same(eng.common(files, 1)[0], u, "common -n 1")
    same(eng.common(files, len(files))[0], eng.inter(files)[0], "common -n all")


def test_setops_device_resident(eng):
    import torch
    files = member_files(1_000_000, 4)
    dfiles = [torch.from_numpy(f.view(np.int64)).cuda() for f in files]
    for name in ("inter", "diff", "union"):
        got = getattr(eng, name)(dfiles)[0]
        assert got.is_cuda
        same(got.cpu().numpy().view(U64), getattr(oracle, name)(files)[0], f"device {name}")
    # device generator == oracle generator
    for f in range(3):
        d = eng.synth_member_file(0, 1_000_000, 1_000_000, 3, 4, f)
        same(d.cpu().numpy().view(U64), files[f], f"synth file {f}")
    # unaligned device slices (8 mod 16)
    sl = [d[1:] for d in dfiles]
    exp = oracle.inter([f[1:] for f in files])[0]
    same(eng.inter(sl)[0].cpu().numpy().view(U64), exp, "unaligned device inter")


def test_partition_sorted(eng):
    import torch
    f = member_files(500_000, 1)[0]
    splitters = np.array([(i << 62) // 8 for i in range(1, 8)], dtype=U64)
    exp = np.concatenate([[0], np.searchsorted(f, splitters, side="left"), [len(f)]]).astype(U64)
    same(eng.partition_sorted(f, splitters), exp, "host partition")
    same(eng.partition_sorted(torch.from_numpy(f.view(np.int64)).cuda(), splitters), exp, "device partition")


# ---------------------------------------------------------------------------------------
# a1-a3: iterators and count
# ---------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def kat(golden_dir):
    with open(os.path.join(golden_dir, "kat.json")) as fh:
        return json.load(fh)


@pytest.fixture(scope="module")
def genomes(golden_dir):
    z = np.load(os.path.join(golden_dir, "genomes.npz"))
    return {n: unpack2(z[n], int(z[n + "_len"])) for n in ("mg1655", "iai39")}


def one_record(seq):
    return np.array([0, len(seq)], dtype=U64)


def test_kat_k1_k2_k4_k7_on_gpu(eng, kat, genomes):
    """K1, K2, K4-K7 (README.md:156-278) computed by the CUDA path."""
    sets = {n: eng.count(s, one_record(s), 23, canonical=True) for n, s in genomes.items()}
    assert len(sets["mg1655"]) == 4546632 and len(sets["iai39"]) == 4902266
    assert digest(sets["mg1655"]) == kat["digests"]["mg1655_k23"]
    assert digest(sets["iai39"]) == kat["digests"]["iai39_k23"]
    a, b = sets["iai39"], sets["mg1655"]
    u, i, d = eng.union([a, b])[0], eng.inter([a, b])[0], eng.diff([a, b])[0]
    assert (len(u), len(i), len(d)) == (6872728, 2576170, 2326096)
    assert digest(u) == kat["digests"]["union"] and digest(i) == kat["digests"]["inter"] and digest(d) == kat["digests"]["diff"]
    assert [oracle.decode(int(c), 23).decode() for c in sets["mg1655"][:3]] == kat["K7_first3_sorted_mg1655"]


def test_kat_k8_k9_nthash_on_gpu(eng, kat, genomes):
    for kmer, h in kat["K8_nthash_k23_canonical"].items():
        got = eng.kmers(kmer.encode(), one_record(kmer), 23, canonical=True, hashed=True)
        assert int(got[0]) == h
    mg = genomes["mg1655"]
    hs = eng.count(mg, one_record(mg), 31, canonical=True, hashed=True)
    assert digest(hs) == kat["digests"]["mg1655_k31_nthash"]
    sc = eng.count(mg, one_record(mg), 31, canonical=True, hashed=True, scaled=True, max_hash=kat["max_hash_scale15"])
    assert len(sc) == 586734 and digest(sc) == kat["digests"]["mg1655_k31_nthash_scaled15"]
    assert digest(eng.count(mg, one_record(mg), 31, canonical=False)) == kat["digests"]["mg1655_k31_kmer_noncanonical"]
    assert digest(eng.count(mg, one_record(mg), 21, canonical=True, circular=True)) == kat["digests"]["mg1655_k21_circular"]


def test_kat_k12_minimizer_on_gpu(eng, kat, genomes):
    """analysis/distance/README.md:8: `count -k 31 -K -H -W 15` on MG1655 = 549 963 k-mers."""
    mg = genomes["mg1655"]
    mz = eng.count_minimizer(mg, one_record(mg), 31, 15, canonical=True)
    assert len(mz) == 549963 and digest(mz) == kat["digests"]["mg1655_k31_minimizer_w15"]


@pytest.mark.parametrize("canonical", [False, True])
@pytest.mark.parametrize("circular", [False, True])
def test_minimizer_multi_record(eng, canonical, circular):
    """count -W: ragged records (empty, shorter than k, fewer than w k-mers), several windows, the scaled filter;
    windows never span records."""
    r = rng(43)
    lens = [0, 5, 30, 31, 32, 45, 46, 100, 17_408, 17_409 + 14, 40_000, 1, 69, 0, 250_000]
    recs = [r.choice(np.frombuffer(b"ACGTacgtN", dtype=np.uint8), L).astype(np.uint8) for L in lens]
    bases = np.concatenate(recs)
    off = np.concatenate([[0], np.cumsum(lens)]).astype(U64)
    mh = int(float(2**64 - 1) / 4.0)
    for k, w in ((31, 15), (31, 1), (21, 2), (31, 16), (33, 50), (31, 300_000)):
        exp = oracle.count_minimizer(bases, off, k, w, canonical=canonical, circular=circular)
        same(eng.count_minimizer(bases, off, k, w, canonical=canonical, circular=circular), exp, f"minimizer k={k} w={w}")
    exp = oracle.count_minimizer(bases, off, 31, 15, canonical=canonical, circular=circular, scaled=True, max_hash=mh)
    same(eng.count_minimizer(bases, off, 31, 15, canonical=canonical, circular=circular, scaled=True, max_hash=mh), exp, "minimizer scaled")
    same(eng.count_minimizer(bases, off, 31, 1, canonical=canonical, circular=circular),
         eng.count(bases, off, 31, canonical=canonical, hashed=True, circular=circular), "w=1 is the plain hashed count")


@pytest.mark.parametrize("hashed", [False, True])
@pytest.mark.parametrize("canonical", [False, True])
@pytest.mark.parametrize("circular", [False, True])
def test_iterator_multi_record(eng, hashed, canonical, circular):
    """Records of ragged lengths incl. empty and shorter-than-k (skipped, count.go:324-328), lower case,
    N and IUPAC codes; every k-mer in record-then-position order."""
    r = rng(41)
    lens = [0, 5, 30, 31, 32, 33, 100, 17_408, 17_409, 40_000, 1, 69, 68 * 256 + 7, 0, 250_000]
    recs = []
    for L in lens:
        s = r.choice(np.frombuffer(b"ACGT", dtype=np.uint8), L)
        if L > 50:
            idx = r.integers(0, L, max(1, L // 40))
            s[idx] = r.choice(np.frombuffer(b"acgtNnRYKMSWBDHVu", dtype=np.uint8), len(idx))
        recs.append(s.astype(np.uint8))
    bases = np.concatenate(recs)
    off = np.concatenate([[0], np.cumsum(lens)]).astype(U64)
    for k in ([5, 31, 32] if not hashed else [5, 31, 33, 64]):
        exp = np.concatenate([(oracle.nthash_iter if hashed else oracle.kmer_iter)(s, k, canonical, circular) for s in recs] + [np.zeros(0, dtype=U64)])
        got = eng.kmers(bases, off, k, canonical=canonical, hashed=hashed, circular=circular)
        same(got, exp, f"iterator k={k} hashed={hashed} canonical={canonical} circular={circular}")
    k = 31
    same(eng.count(bases, off, k, canonical=canonical, hashed=hashed, circular=circular),
         oracle.count(bases, off, k, canonical=canonical, hashed=hashed, circular=circular), "count multi-record")


@pytest.mark.parametrize("limit", ["1000000000", "200000", "30000"])
def test_count_key_range_passes(eng, limit, monkeypatch):
    """count cuts the code space into key-range passes when the k-mers would not fit (C4); forced here on a
    small input.  Includes a skewed 2-bit case (poly-A) that overflows a pass buffer and is retried."""
    monkeypatch.setenv("UKM_COUNT_PASS", limit)
    recs = [oracle.synth_bases(r_, 0, n, 5) for r_, n in enumerate((300_000, 17, 120_000))]
    recs.append(np.frombuffer(b"A" * 50_000 + b"ACGT" * 10_000 + b"T" * 30_000, dtype=np.uint8))
    bases = np.concatenate(recs)
    off = np.concatenate([[0], np.cumsum([len(r_) for r_ in recs])]).astype(U64)
    mh = int(float(2**64 - 1) / 7.0)
    for kw in ({"hashed": True}, {"hashed": False}, {"hashed": True, "scaled": True, "max_hash": mh}, {"hashed": False, "canonical": False}):
        k = 31 if kw.get("hashed") else 27
        kw = {"canonical": True, **kw}
        same(eng.count(bases, off, k, **kw), oracle.count(bases, off, k, **kw), f"count passes={limit} {kw}")


def test_illegal_base_is_an_error(eng):
    import unikmer_b200 as ub
    seq = b"ACGTACGTACGTACGTAC*TACGTACGTACGTACGTACGTACGT"
    with pytest.raises(oracle.OracleError):
        oracle.kmer_iter(seq, 11)
    with pytest.raises(ub.UkmError) as ei:
        eng.kmers(seq, one_record(seq), 11, canonical=True)
    assert ei.value.status == ub.E_ILLEGAL_BASE
    # ntHash does not validate: the byte contributes a zero seed
    same(eng.kmers(seq, one_record(seq), 11, canonical=True, hashed=True), oracle.nthash_iter(seq, 11, True), "nthash with junk byte")


def test_c4_generator_and_count_device(eng):
    """C4 synthetic FASTA generator on the device == oracle; count -k 31 -K -H on it."""
    import torch
    L = 300_000
    recs = [oracle.synth_bases(r_, 0, L, 5) for r_ in range(3)]
    d = torch.cat([eng.synth_bases(r_, 0, L, 5) for r_ in range(3)])
    same(d.cpu().numpy(), np.concatenate(recs), "synth bases")
    off = np.array([0, L, 2 * L, 3 * L], dtype=U64)
    doff = torch.from_numpy(off.view(np.int64)).cuda()
    got = eng.count(d, doff, 31, canonical=True, hashed=True)
    same(got.cpu().numpy().view(U64), oracle.count(np.concatenate(recs), off, 31, canonical=True, hashed=True), "C4 count")


# ---------------------------------------------------------------------------------------
# size-independent properties at large sizes
# ---------------------------------------------------------------------------------------
def test_large_sort_properties(eng):
    """C2-like: 2e8 random 62-bit keys on the device: non-decreasing + multiset checksums preserved."""
    import torch
    n = 200_000_000
    d = eng.synth_random_keys(0, n, 2)
    s0 = int(d.sum().item())
    eng.sort(d, key_bits=62)
    assert bool((d[1:] >= d[:-1]).all().item()), "not sorted"  # keys < 2^62: signed compare is fine
    assert int(d.sum().item()) == s0
    # idempotence + agreement with torch.sort on a slice
    head = d[:1_000_000].clone()
    eng.sort(head, key_bits=62)
    assert torch.equal(head, d[:1_000_000])
    del d


def test_c2_full_size_sort(eng):
    """C2 at BASELINE.json's full size: 1e9 random 62-bit keys (8 GB) sorted on the device, checked through
    size-independent properties: non-decreasing, key sum preserved (mod 2^64), and every generated key that is below
    the 2e6-th smallest (taken from the first 5e7 generated keys, via the oracle's generator) sits in the sorted head."""
    n = 1_000_000_000
    d = eng.synth_random_keys(0, n, 2)
    s0 = int(d.sum().item())
    eng.sort(d, key_bits=62)
    assert bool((d[1:] >= d[:-1]).all().item()), "not sorted"  # keys < 2^62: signed compare is fine
    assert int(d.sum().item()) == s0
    bound = int(d[2_000_000].item())
    gen = oracle.random_keys(0, 50_000_000, 2)
    small = gen[gen < np.uint64(bound)]
    head = d[:2_000_001].cpu().numpy().view(np.uint64)
    assert len(small) > 50_000 and np.isin(small, head).all()
    del d


def test_large_setop_properties(eng):
    """C3-like at 8 x 2.5e7: inter is a subset of every input, diff is disjoint from the subjects,
    |A u B| = |A| + |B| - |A n B|, expected cardinalities ~ N/256."""
    import torch
    N = 50_000_000
    files = [eng.synth_member_file(0, N, N, 3, 4, f).clone() for f in range(8)]
    inter = eng.inter(files)[0]
    diff = eng.diff(files)[0]
    union2 = eng.union(files[:2])[0]
    inter2 = eng.inter(files[:2])[0]
    assert len(union2) == len(files[0]) + len(files[1]) - len(inter2)
    assert abs(len(inter) - N / 256) < 6 * (N / 256) ** 0.5
    assert abs(len(diff) - N / 256) < 6 * (N / 256) ** 0.5
    for f in files:
        assert len(eng.inter([inter, f])[0]) == len(inter)
    for f in files[1:]:
        assert len(eng.inter([diff, f])[0]) == 0
    assert len(eng.inter([diff, files[0]])[0]) == len(diff)
    u8 = eng.union(files)[0]
    assert bool((u8[1:] > u8[:-1]).all().item())
    assert len(eng.common(files, 1)[0]) == len(u8)
    assert len(eng.common(files, 8)[0]) == len(inter)


@pytest.mark.parametrize("variant", ["global", "per_kmer"])
def test_common_c5_shape(eng, tax, variant):
    """BASELINE.json configs[4] at 1/100 of its size: `common -n 32` over 64 files x ~1e6 k-mers with TaxId LCA
    (common.go:93-105 threshold, 220-283 counting + LCA fold, 329-354 select + sort).  SURVEY.md 8(d) generators: universe
    U(j; 2e6, S=6), file f holds U_j iff bit f of sm64(7 + j); either every file carries one global taxid (the README
    workflow, README.md:169-171) or every k-mer its own, 1 + sm64(9 + 64 j + f) % 1e4.  (LCA semantics on unknown / merged
    ids are those of the oracle's restatement: the reference's are unpinned)"""
    from unikmer_b200 import KmerSet
    otax, _ = tax
    N, NF, S, T, R = 2_000_000, 64, 6, 7, 9
    W = (1 << 62) // N
    keys = [oracle.member_file(0, N, N, S, T, f) for f in range(NF)]
    if variant == "global":
        leaf = [10_000 - f for f in range(NF)]
        ofiles = [(k, np.full(len(k), t, dtype=np.uint32)) for k, t in zip(keys, leaf)]
        gsets = [KmerSet(k, None, global_taxid=t) for k, t in zip(keys, leaf)]
    else:
        ofiles = []
        for f, k in enumerate(keys):
            j = (k // U64(W)).astype(np.uint64)
            x = (U64(R) + j * U64(64) + U64(f))
            # sm64 over an array: the same mixer, vectorised
            z = x + U64(0x9E3779B97F4A7C15)
            z = (z ^ (z >> U64(30))) * U64(0xBF58476D1CE4E5B9)
            z = (z ^ (z >> U64(27))) * U64(0x94D049BB133111EB)
            z = z ^ (z >> U64(31))
            assert int(z[0]) == oracle.sm64(int(x[0]))
            ofiles.append((k, (U64(1) + z % U64(10_000)).astype(np.uint32)))
        gsets = [KmerSet(k, t) for k, t in ofiles]
    ek, et = oracle.common(ofiles, 32, has_taxid=True, tax=otax)
    gk, gt = eng.common(gsets, 32, has_taxid=True)
    assert 0.4 * N < len(ek) < 0.7 * N  # P[Bin(64, 1/2) >= 32] ~ 0.55 of the universe
    same(gk, ek, f"C5 {variant}: keys")
    same(gt, et, f"C5 {variant}: taxids")
