"""Streamed set operations on host-resident inputs (ukm_setops_stream; also the path ukm_inter / ukm_diff / ukm_union take
on their own when every input is in host memory): key ranges uploaded, computed and downloaded as a pipeline, every input
byte crossing PCIe once.  Results must equal the whole-file results of the oracle (inter.go:188-286, diff.go:380-435,
union.go:186-208), including the whole-file rules for empty inputs."""
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


@pytest.fixture(autouse=True)
def small_chunks(monkeypatch):
    monkeypatch.setenv("UKM_STREAM_MIN_MB", "0")   # stream whatever the size
    monkeypatch.setenv("UKM_STREAM_CHUNK_MB", "1")  # many key ranges on small inputs


def check(eng, files, what, ops=("inter", "diff", "union")):
    got = eng.setops(files, ops)
    exp = {"inter": oracle.inter, "diff": oracle.diff, "union": oracle.union}
    for o, g in zip(ops, got):
        same(g, exp[o](files)[0], f"{what}: {o}")


@pytest.mark.parametrize("nf", [2, 3, 8])
def test_stream_matches_oracle(eng, nf):
    for N in (5_000, 400_000, 2_000_000):
        check(eng, member_files(N, nf), f"nf {nf} N {N}")


def test_stream_single_ops_take_the_streamed_path(eng):
    files = member_files(1_500_000, 8)
    same(eng.inter(files)[0], oracle.inter(files)[0], "inter")
    same(eng.diff(files)[0], oracle.diff(files)[0], "diff")
    same(eng.union(files)[0], oracle.union(files)[0], "union")


def test_stream_any_subset_and_order_of_operations(eng):
    files = member_files(600_000, 5)
    check(eng, files, "union only", ("union",))
    check(eng, files, "diff, inter", ("diff", "inter"))
    check(eng, files, "twice", ("inter", "inter", "union"))


def test_stream_pinned_host_buffers(eng):
    import torch
    files = member_files(1_000_000, 8)
    pin = [torch.from_numpy(f.view(np.int64)).pin_memory() for f in files]
    outs = [torch.empty(len(files[0]), dtype=torch.int64).pin_memory(), torch.empty(len(files[0]), dtype=torch.int64).pin_memory(),
            torch.empty(sum(len(f) for f in files), dtype=torch.int64).pin_memory()]
    gi, gd, gu = eng.setops(pin, ("inter", "diff", "union"), outs=outs)
    same(gi.numpy().view(U64), oracle.inter(files)[0], "pinned inter")
    same(gd.numpy().view(U64), oracle.diff(files)[0], "pinned diff")
    same(gu.numpy().view(U64), oracle.union(files)[0], "pinned union")


def test_stream_distributions_and_empty_files(eng):
    import unikmer_b200 as ub
    r = rng(7)
    base = np.unique(r.integers(0, 2**64, 400_000, dtype=U64))
    cases = {
        "extremes": [np.unique(np.concatenate([base[r.random(len(base)) < 0.6], np.array([0, 2**64 - 1], dtype=U64)])) for _ in range(4)],
        "disjoint ranges": [np.unique(r.integers(q << 50, (q << 50) + 2**30, 30_000 << (q % 3), dtype=U64)) for q in range(5)],
        "largest is not first": [base[::7].copy(), base.copy(), base[::3].copy()],
        "clustered": [np.unique(np.concatenate([r.integers(0, 2**63, 20_000, dtype=U64), U64(10**12) + np.arange(q, 300_000, 1 + q, dtype=U64)]))
                      for q in range(4)],
        # inter.go:211-215: an empty later file ends the loop and keeps the current set (B-3), on the WHOLE files
        "empty later file": [base[::2].copy(), base[::3].copy(), np.zeros(0, dtype=U64), base[::5].copy()],
        "tiny": [np.array([1, 5, 9], dtype=U64), np.array([5], dtype=U64), np.array([5, 9], dtype=U64)],
    }
    for name, files in cases.items():
        got = eng.setops(files, ("inter", "diff", "union"))
        same(got[0], oracle.inter(files)[0], f"{name}: inter")
        exp_d = files[0]
        for f in files[1:]:
            exp_d = exp_d[~np.isin(exp_d, f)]  # an empty subject is skipped (documented deviation B-5)
        same(got[1], exp_d, f"{name}: diff")
        same(got[2], oracle.union(files)[0], f"{name}: union")
    with pytest.raises(ub.UkmError) as e:  # inter.go:208 panics on an empty first file
        eng.setops([np.zeros(0, dtype=U64), base], ("inter",))
    assert e.value.status == ub.E_PANIC


def test_stream_capacity_error(eng):
    import unikmer_b200 as ub
    files = member_files(300_000, 3)
    small = np.empty(10, dtype=U64)
    with pytest.raises(ub.UkmError) as e:
        eng.setops(files, ("union",), outs=[small])
    assert e.value.status == ub.E_CAPACITY


def test_stale_device_error_is_not_reported_by_the_next_call(eng):
    """A call that fails after a kernel set the device error word (unsorted input under UKM_F_VALIDATE) must not leave it
    behind for the next, unrelated call."""
    import unikmer_b200 as ub
    bad = np.array([5, 3, 9, 9], dtype=U64)
    good = member_files(50_000, 2)
    with pytest.raises(ub.UkmError):
        eng.union([bad, good[0]], validate=True)
    same(eng.union(good)[0], oracle.union(good)[0], "call after a failed one")


def test_inter_shard_flag_turns_whole_file_quirks_off(eng):
    base = member_files(100_000, 3)
    empty = np.zeros(0, dtype=U64)
    assert len(eng.inter([empty, base[0]], shard=True)[0]) == 0              # no panic: an empty slice, not an empty file
    assert len(eng.inter([base[0], empty, base[1]], shard=True)[0]) == 0     # plain set semantics, not quirk B-3
    same(eng.inter([base[0], empty, base[1]])[0], base[0], "B-3 on whole files keeps the current set")


def test_fused_inter_diff_on_device_spans(eng):
    """inter + diff of the same device-resident files in one ukm_setops_stream call share ONE pass (nfilter.cu, NFOP_BOTH)."""
    import torch
    for nf in (2, 3, 5, 8):
        files = member_files(1_200_000, nf)
        dev = [torch.from_numpy(f.view(np.int64)).cuda() for f in files]
        eng.stats_reset()
        eng.stats_enable(True)
        gi, gd, gu = eng.setops(dev, ("inter", "diff", "union"))
        eng.stats_enable(False)
        st = eng.stats()
        if nf >= 3:  # all three results from ONE pass: the union kernel, inter / diff riding along (nway.cu)
            assert st.get("setop_inter_diff_union_nway", {}).get("launches") == 1, st
            assert "setop_inter_nway" not in st and "setop_diff_nway" not in st and "setop_union_nway" not in st, st
        same(gi.cpu().numpy().view(U64), oracle.inter(files)[0], f"fused inter nf {nf}")
        same(gd.cpu().numpy().view(U64), oracle.diff(files)[0], f"fused diff nf {nf}")
        same(gu.cpu().numpy().view(U64), oracle.union(files)[0], f"union nf {nf}")


@pytest.mark.parametrize("sub", ["0", "64", "8"])
def test_fused_inter_diff_distributions(eng, sub, monkeypatch):
    import torch
    monkeypatch.setenv("UKM_NFILTER_SUB", sub)
    r = rng(31)
    base = np.unique(r.integers(0, 2**64, 300_000, dtype=U64))
    cases = {
        "extremes": [np.unique(np.concatenate([base[r.random(len(base)) < 0.7], np.array([0, 2**64 - 1], dtype=U64)])) for _ in range(6)],
        "identical": [base.copy() for _ in range(4)],
        "disjoint": [np.unique(r.integers(q << 50, (q << 50) + 2**30, 5000 << (q % 5), dtype=U64)) for q in range(6)],
        "subjects_dense": [np.unique(np.concatenate([r.integers(0, 2**63, 30_000, dtype=U64), U64(10**12) + np.arange(0, 400_000, 997, dtype=U64)]))] +
                          [np.unique(np.concatenate([r.integers(0, 2**63, 30_000, dtype=U64), U64(10**12) + np.arange(q, 400_000, 1 + q, dtype=U64)])) for q in range(5)],
        "big_first": [base.copy()] + [base[::40 + q].copy() for q in range(4)],
        "tiny": [np.array([1, 5, 9], dtype=U64), np.array([5], dtype=U64), np.array([5, 9], dtype=U64)],
    }
    for name, files in cases.items():
        dev = [torch.from_numpy(f.view(np.int64)).cuda() for f in files]
        gi, gd = eng.setops(dev, ("inter", "diff"))
        same(gi.cpu().numpy().view(U64), oracle.inter(files)[0], f"fused inter {name} sub {sub}")
        same(gd.cpu().numpy().view(U64), oracle.diff(files)[0], f"fused diff {name} sub {sub}")


def test_fused_inter_diff_whole_file_rules_and_shards(eng):
    import torch
    base = member_files(200_000, 4)
    empty = np.zeros(0, dtype=U64)
    files = [base[0], base[1], empty, base[2]]
    dev = [torch.from_numpy(f.view(np.int64)).cuda() for f in files]
    gi, gd = eng.setops(dev, ("inter", "diff"))  # whole files: B-3 keeps file0 AND file1, diff skips the empty subject
    same(gi.cpu().numpy().view(U64), oracle.inter(files)[0], "inter with an empty later file (B-3)")
    exp_d = base[0][~np.isin(base[0], base[1])]
    exp_d = exp_d[~np.isin(exp_d, base[2])]
    same(gd.cpu().numpy().view(U64), exp_d, "diff skips the empty subject")
    gi, gd = eng.setops(dev, ("inter", "diff"), shard=True)  # slices: the empty slice empties the intersection
    assert gi.shape[0] == 0
    same(gd.cpu().numpy().view(U64), exp_d, "diff of slices")
    many = member_files(150_000, 8) + member_files(150_000, 3, S=11, T=12)  # more than eight files: not fused, still right
    devm = [torch.from_numpy(f.view(np.int64)).cuda() for f in many]
    gi, gd = eng.setops(devm, ("inter", "diff"))
    same(gi.cpu().numpy().view(U64), oracle.inter(many)[0], "inter of 11 files")
    same(gd.cpu().numpy().view(U64), oracle.diff(many)[0], "diff of 11 files")


@pytest.mark.parametrize("nf", [3, 4, 5, 7, 8])
def test_union_with_inter_diff_riding_along(eng, nf):
    """ukm_setops_stream with all three operations on device spans: ONE pass (nway.cu -- a run of equal keys in the last
    merge level is as long as the number of files that hold the key: run = nf is the intersection, run = 1 with the key in
    file 0 the difference).  Sizes that put runs across thread, tile and mask-word seams; distributions with long runs."""
    import torch
    r = rng(nf)
    cases = {f"members {N}": member_files(N, nf) for N in (2_000, 100_000, 3_000_000)}
    base = np.unique(r.integers(0, 2**64, 500_000, dtype=U64))
    cases["identical"] = [base.copy() for _ in range(nf)]                       # every run is nf long
    cases["disjoint"] = [base[q::nf].copy() for q in range(nf)]                 # every run is 1 long
    cases["file0 subset"] = [base[::3].copy()] + [base.copy() for _ in range(nf - 1)]
    cases["file0 superset"] = [base.copy()] + [base[q::5].copy() for q in range(nf - 1)]
    cases["extremes"] = [np.unique(np.concatenate([base[r.random(len(base)) < 0.5], np.array([0, 2**64 - 1], dtype=U64)])) for _ in range(nf)]
    for name, files in cases.items():
        dev = [torch.from_numpy(f.view(np.int64)).cuda() for f in files]
        gi, gd, gu = eng.setops(dev, ("inter", "diff", "union"))
        same(gu.cpu().numpy().view(U64), oracle.union(files)[0], f"{name}: union")
        same(gi.cpu().numpy().view(U64), oracle.inter(files)[0], f"{name}: inter")
        same(gd.cpu().numpy().view(U64), oracle.diff(files)[0], f"{name}: diff")


def test_fusion_switch_and_order_of_outputs(eng, monkeypatch):
    import torch
    files = member_files(800_000, 8)
    dev = [torch.from_numpy(f.view(np.int64)).cuda() for f in files]
    exp = {"inter": oracle.inter(files)[0], "diff": oracle.diff(files)[0], "union": oracle.union(files)[0]}
    for ops in (("union", "diff", "inter"), ("diff", "union", "inter", "union")):
        for g, o in zip(eng.setops(dev, ops), ops):
            same(g.cpu().numpy().view(U64), exp[o], f"{ops}: {o}")
    monkeypatch.setenv("UKM_FUSE3", "0")  # inter + diff fused (chunk filter), union on its own
    eng.stats_reset()
    eng.stats_enable(True)
    got = eng.setops(dev, ("inter", "diff", "union"))
    eng.stats_enable(False)
    st = eng.stats()
    assert st.get("setop_inter_diff_nway", {}).get("launches") == 1 and st.get("setop_union_nway", {}).get("launches") == 1, st
    for g, o in zip(got, ("inter", "diff", "union")):
        same(g.cpu().numpy().view(U64), exp[o], f"UKM_FUSE3=0: {o}")
