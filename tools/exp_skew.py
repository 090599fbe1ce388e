#!/usr/bin/env python
"""inter / diff on the C3 inputs for several values of UKM_SETOP_SKEW (|B| >= skew * |A| -> look A up in B instead of
walking both).  One JSON line per value."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tools.exp_nway import timed  # noqa: E402
from unikmer_b200 import Engine  # noqa: E402


def main():
    U = int(float(sys.argv[1])) if len(sys.argv) > 1 else 10**9
    eng = Engine(0)
    stream = torch.cuda.Stream()
    eng.use_stream(stream.cuda_stream)
    with torch.cuda.stream(stream):
        files = [eng.synth_member_file(0, U, U, 3, 4, f).clone() for f in range(8)]
        out = torch.empty(int(files[0].shape[0]) + 16, dtype=torch.int64, device="cuda")
        for skew in ("3", "6", "10", "12", "16", "24", "40", "0"):
            os.environ["UKM_SETOP_SKEW"] = skew
            res = {}
            for name, fn in (("inter", eng.inter), ("diff", eng.diff)):
                eng.stats_reset(); eng.stats_enable(True)
                ms = timed(stream, lambda: fn(files, out=out), reps=3)
                eng.stats_enable(False)
                res[name] = round(ms, 3)
            print(json.dumps({"bench": "skew", "skew": skew, **res}), flush=True)
    eng.close()


if __name__ == "__main__":
    main()
