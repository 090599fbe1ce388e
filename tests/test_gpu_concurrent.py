"""Two contexts on ONE GPU running at the same time (include/ukm.h: a ukm_ctx is single-owner, different contexts may run
concurrently).  The persistent kernels (two-way pipeline, N-way union) hand output offsets from CTA to CTA and therefore
need their whole grid resident: they are launched cooperatively, so a second context sharing the GPU can delay them but
never turns a wait into a watchdog UKM_E_INTERNAL; the one-tile-per-CTA kernels order their look-back by ticket."""
import threading

import numpy as np
import pytest

import oracle
from tests.test_gpu_parity import member_files, same

pytestmark = pytest.mark.gpu


def test_two_contexts_share_one_gpu():
    from unikmer_b200 import Engine
    files = member_files(3_000_000, 8)
    pair = files[:2]
    exp = {"union8": oracle.union(files)[0], "inter8": oracle.inter(files)[0], "diff8": oracle.diff(files)[0],
           "union2": oracle.union(pair)[0], "inter2": oracle.inter(pair)[0]}
    errors, results = [], [{}, {}]

    def worker(w):
        try:
            eng = Engine(0)
            import torch
            dev = [torch.from_numpy(f.view(np.int64)).cuda() for f in files]
            for _ in range(6):
                results[w]["union8"] = eng.union(dev)[0].cpu().numpy().view(np.uint64)   # N-way union (persistent, cooperative)
                results[w]["inter8"] = eng.inter(dev)[0].cpu().numpy().view(np.uint64)   # single-pass filter (no CTA dependency)
                results[w]["diff8"] = eng.diff(dev)[0].cpu().numpy().view(np.uint64)
                results[w]["union2"] = eng.union(dev[:2])[0].cpu().numpy().view(np.uint64)  # two-way pipeline (persistent)
                results[w]["inter2"] = eng.inter(dev[:2])[0].cpu().numpy().view(np.uint64)
            eng.close()
        except Exception as e:  # noqa: BLE001 -- reported by the main thread
            errors.append((w, repr(e)))

    ts = [threading.Thread(target=worker, args=(w,)) for w in range(2)]
    for t in ts:
        t.start()
    for t in ts:
        t.join(timeout=600)
    assert not errors, errors
    for w in range(2):
        for k, v in exp.items():
            same(results[w][k], v, f"context {w}: {k}")
