#!/usr/bin/env python
"""The two-way pipeline kernel with and without its merge work (UKM_SETOP_NULL=1: load, scan, offset hand-off and
copy-out only -- results are not valid) on one pass F0 x F1 of the C3 files: what the tile machinery alone costs.
Needs a measurement build of the library:  make -C unikmer_b200/csrc clean all EXTRA=-DUKM_MEASURE"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tools.exp_nway import timed  # noqa: E402
from unikmer_b200 import Engine  # noqa: E402

eng = Engine(0)
stream = torch.cuda.Stream()
eng.use_stream(stream.cuda_stream)
with torch.cuda.stream(stream):
    U = 10**9
    files = [eng.synth_member_file(0, U, U, 3, 4, f).clone() for f in range(2)]
    out = torch.empty(int(files[0].shape[0]) * 2 + 16, dtype=torch.int64, device="cuda")
    for null in ("0", "1"):
        os.environ["UKM_SETOP_NULL"] = null
        for pipe in ("1", "0", "4"):
            os.environ["UKM_SETOP_PIPE"] = pipe
            for name, fn in (("inter", eng.inter), ("union", eng.union), ("merge", eng.merge)):
                eng.stats_reset(); eng.stats_enable(True)
                ms = timed(stream, lambda: fn(files) if name == "merge" else fn(files, out=out), reps=3)
                eng.stats_enable(False)
                print(json.dumps({"null": null, "pipe_cfg": pipe, "op": name, "ms": round(ms, 3),
                                  "k": {k: round(v["ms"] / max(v["launches"], 1), 3) for k, v in eng.stats().items()}}), flush=True)
