#!/usr/bin/env python
"""inter / diff on the C3 inputs for every look-up mode of setop_search_kernel (UKM_SEARCH_MODE: 0 bisection,
1 interpolated start + gallop, 2 the same anchored on the thread's first item).  One JSON line per mode, with a
checksum of the result so that the modes can be compared."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tools.exp_nway import timed  # noqa: E402
from unikmer_b200 import Engine  # noqa: E402


def main():
    once = "--once" in sys.argv  # one untimed inter per mode (for an ncu launch list of the look-up kernels)
    args = [a for a in sys.argv[1:] if a != "--once"]
    U = int(float(args[0])) if args else 10**9
    eng = Engine(0)
    stream = torch.cuda.Stream()
    eng.use_stream(stream.cuda_stream)
    with torch.cuda.stream(stream):
        files = [eng.synth_member_file(0, U, U, 3, 4, f).clone() for f in range(8)]
        out = torch.empty(int(files[0].shape[0]) + 16, dtype=torch.int64, device="cuda")
        if once:
            for mode in ("0", "1", "2"):
                os.environ["UKM_SEARCH_MODE"] = mode
                eng.inter(files, out=out)
            stream.synchronize()
            eng.close()
            return
        cfgs = [(skew, mode) for skew in ("6", "16") for mode in ("0", "1", "2")]  # (skew threshold, look-up mode)
        for skew, mode in cfgs:
            os.environ["UKM_SETOP_SKEW"] = skew
            os.environ["UKM_SEARCH_MODE"] = mode
            res = {}
            for name, fn in (("inter", eng.inter), ("diff", eng.diff)):
                ms = timed(stream, lambda: fn(files, out=out), reps=5)
                r = fn(files, out=out)[0]
                stream.synchronize()
                res[name] = round(ms, 3)
                res[name + "_n"] = int(r.shape[0])
                res[name + "_sum"] = int(r.sum().item())
            print(json.dumps({"bench": "search_mode", "skew": skew, "mode": mode, **res}), flush=True)
    eng.close()


if __name__ == "__main__":
    main()
