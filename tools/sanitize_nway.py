#!/usr/bin/env python
"""Small N-way union / filter cases for compute-sanitizer runs (memcheck, racecheck, synccheck, initcheck):

    compute-sanitizer --tool racecheck python tools/sanitize_nway.py
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import oracle  # noqa: E402
from unikmer_b200 import Engine  # noqa: E402


def main():
    eng = Engine(0)
    os.environ["UKM_NWAY_FORCE"] = "1"
    bad = 0
    for cfg in ("0", "1", "4"):
        os.environ["UKM_NWAY_CFG"] = cfg
        for nf, N in ((2, 6000), (5, 20000), (8, 40000), (8, 300)):
            files = [oracle.member_file(0, N, N, 3, 4, f) for f in range(nf)]
            u = eng.union(files)[0]
            exp = np.unique(np.concatenate(files))
            ok = np.array_equal(u, exp)
            bad += not ok
            print(f"union cfg {cfg} nf {nf} N {N}: {'ok' if ok else 'MISMATCH'}", flush=True)
            if nf >= 3:
                for name, fn, orc in (("inter", eng.inter, oracle.inter), ("diff", eng.diff, oracle.diff)):
                    ok = np.array_equal(fn(files)[0], orc(files)[0])
                    bad += not ok
                    print(f"{name} cfg {cfg} nf {nf} N {N}: {'ok' if ok else 'MISMATCH'}", flush=True)
    eng.close()
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
