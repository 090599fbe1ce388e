"""CPU-side checks of the measurement contract and of the constants shared between include/ukm.h and the Python host
layer: the reference arm of bench.py prints the agreed JSON line (on a tiny sample), the flag / operation constants of
the header and of unikmer_b200/_lib.py agree, and the full-size digests the bench checks its results against
(oracle.c3_digest) equal the digests of the oracle's own set operations on a small universe."""
import json
import os
import re
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_the_contract_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--ref-universe", "3e5", "--steps", "2",
                          "--warmup", "1", "--gpus", "1"], capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1, "exactly ONE JSON line"
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "sorted_uint64_kmers_per_sec_union_inter_diff" and d["unit"] == "k-mers/s"
    assert d["higher_is_better"] is True and d["steps"] == 2 and d["warmup"] == 1 and d["n_gpus"] == 1
    assert d["value"] > 0 and d["ms_per_step"] > 0 and d["vs_baseline"] is None and d["gpu_launches"] == 0
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and "universe" in cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and "model" not in d["config"]


def test_reference_arm_other_ranks_do_nothing():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--ref-universe", "3e5", "--steps", "1",
                          "--warmup", "0", "--gpus", "2"], capture_output=True, text=True, timeout=300, cwd=ROOT, env=env)
    assert out.returncode == 0 and out.stdout.strip() == ""


def test_header_constants_match_the_python_layer():
    from unikmer_b200 import _lib as L
    h = open(os.path.join(ROOT, "include", "ukm.h")).read()
    flags = {m.group(1): int(m.group(2)) for m in re.finditer(r"#define\s+UKM_F_([A-Z_]+)\s+(\d+)u", h)}
    assert flags == {"TAXID": L.F_TAXID, "MIX_TAXID": L.F_MIX_TAXID, "COMPARE_TAXID": L.F_COMPARE_TAXID, "CANONICAL": L.F_CANONICAL,
                     "HASHED": L.F_HASHED, "CIRCULAR": L.F_CIRCULAR, "SCALED": L.F_SCALED, "VALIDATE": L.F_VALIDATE, "SHARD": L.F_SHARD}
    m = re.search(r"typedef enum ukm_setop \{([^}]*)\}", h)
    ops = {k: int(v) for k, v in re.findall(r"UKM_OP_([A-Z]+)\s*=\s*(\d+)", m.group(1))}
    assert ops == {"INTER": L.OP_INTER, "DIFF": L.OP_DIFF, "UNION": L.OP_UNION}
    m = re.search(r"typedef enum ukm_where \{([^}]*)\}", h)
    where = {k: int(v) for k, v in re.findall(r"UKM_([A-Z_]+)\s*=\s*(\d+)", m.group(1))}
    assert where == {"HOST": L.HOST, "HOST_PINNED": L.HOST_PINNED, "DEVICE": L.DEVICE}


def test_c3_digest_equals_the_digest_of_the_set_operations():
    """bench.py checks the FULL C3 results on the device against oracle.c3_digest (computed from the generator's membership
    bits, no set operation involved); here the same digests are taken from the oracle's inter / diff / union."""
    import oracle
    N, S, T, NF = 300_000, 3, 4, 8
    files = [oracle.member_file(0, N, N, S, T, f) for f in range(NF)]
    want = oracle.c3_digest(0, N, N, S, T, NF)
    for name, res in (("inter", oracle.inter(files)[0]), ("diff", oracle.diff(files)[0]), ("union", oracle.union(files)[0])):
        got = oracle.digest3(res)
        assert tuple(int(x) for x in got) == tuple(int(x) for x in want[name]), name
    # and a change of one key changes the digest
    u = oracle.union(files)[0].copy()
    u[len(u) // 2] ^= np.uint64(1)
    assert tuple(int(x) for x in oracle.digest3(u)) != tuple(int(x) for x in want["union"])
