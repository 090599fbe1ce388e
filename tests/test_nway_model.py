"""CPU check of the N-way union arithmetic (unikmer_b200/csrc/nway_core.cuh): the host model
(tests/host/nway_model.cpp) compiles the header the CUDA kernel uses with g++ and replays the
kernel's per-thread schedule against std::set_union semantics -- partition by multi-sequence
selection, merge tables, merge-path splits, plain and de-duplicating walks, every tile shape."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(shutil.which("g++") is None, reason="needs g++")
def test_nway_host_model(tmp_path):
    exe = tmp_path / "nway_model"
    src = os.path.join(ROOT, "tests", "host", "nway_model.cpp")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-Wno-unknown-pragmas", src, "-o", str(exe)])
    out = subprocess.run([str(exe)], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    assert "nway model ok" in out.stdout


@pytest.mark.skipif(shutil.which("g++") is None, reason="needs g++")
def test_rows_host_model(tmp_path):
    """The row-based union (unikmer_b200/csrc/rows_core.cuh, nunion.cu): pair tables with reversed / congruent runs,
    merge-path splits in padded rows, rotated gather + bitonic merge network, 8 warps x 31 rows per level -- replayed on
    the CPU against std::set_union, with the bank-conflict-free claim of the gather checked round by round."""
    exe = tmp_path / "rows_model"
    src = os.path.join(ROOT, "tests", "host", "rows_model.cpp")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-Wno-unknown-pragmas", src, "-o", str(exe)])
    out = subprocess.run([str(exe)], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    assert "rows model ok" in out.stdout and " 0 with a bank conflict" in out.stdout


@pytest.mark.skipif(shutil.which("g++") is None, reason="needs g++")
def test_ride_along_host_model(tmp_path):
    """`inter` / `diff` riding along the union's last level: the run-length bit arithmetic the kernel uses
    (nw_lead_heads / nw_run_candidates, nway_core.cuh) laid over the kernel's thread ranges -- every tile shape, 2..8 files,
    identical / disjoint / empty files, runs placed at every offset against the thread boundaries -- against set algebra."""
    exe = tmp_path / "ride_model"
    src = os.path.join(ROOT, "tests", "host", "ride_model.cpp")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-Wno-unknown-pragmas", src, "-o", str(exe)])
    out = subprocess.run([str(exe)], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    assert "ride-along model ok" in out.stdout


@pytest.mark.skipif(shutil.which("g++") is None, reason="needs g++")
def test_nfilter_search_host_model(tmp_path):
    """The per-lane searches of the single-pass inter / diff filter (unikmer_b200/csrc/nfilter_core.cuh: clamped straight-line
    probes with a warp-uniform depth): every segment length 0..300 and around the powers of two up to 2200, every admissible
    depth, keys on and between the elements -- against std::binary_search / std::lower_bound, with every probe address
    checked to stay inside its segment."""
    exe = tmp_path / "nfilter_model"
    src = os.path.join(ROOT, "tests", "host", "nfilter_model.cpp")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-Wno-unknown-pragmas", src, "-o", str(exe)])
    out = subprocess.run([str(exe)], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    assert "nfilter model ok" in out.stdout and "none outside its segment" in out.stdout
