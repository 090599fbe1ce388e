#!/usr/bin/env python
"""A/B timing of inter / diff / union on the C3 inputs (8 sorted files x ~universe/2 k-mers, one B200) under
environment-variable variants of the library.  One JSON line per (op, variant).

    python tools/exp_ops.py --ops inter,diff --variants "new:;old:UKM_NFILTER=0;cfg1:UKM_NFILTER_CFG=1"
"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from unikmer_b200 import Engine  # noqa: E402


def timed(stream, fn, reps, warm=1):
    for _ in range(warm):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record(stream)
    for _ in range(reps):
        fn()
    e1.record(stream)
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--universe", type=float, default=1e9)
    ap.add_argument("--files", type=int, default=8)
    ap.add_argument("--ops", default="inter,diff,union")
    ap.add_argument("--variants", default="default:")
    ap.add_argument("--reps", type=int, default=5)
    args = ap.parse_args()
    U = int(args.universe)
    eng = Engine(0)
    stream = torch.cuda.Stream()
    eng.use_stream(stream.cuda_stream)
    variants = []
    for v in args.variants.split(";"):
        name, _, envs = v.partition(":")
        variants.append((name, dict(e.split("=", 1) for e in envs.split(",") if e)))
    all_keys = sorted({k for _, d in variants for k in d})
    with torch.cuda.stream(stream):
        files = [eng.synth_member_file(0, U, U, 3, 4, f).clone() for f in range(args.files)]
        total = sum(int(f.shape[0]) for f in files)
        out = torch.empty(min(total, U) + 16, dtype=torch.int64, device="cuda")
        for op in args.ops.split(","):
            fn = getattr(eng, op)
            ref = None
            for name, envs in variants:
                for k in all_keys:
                    os.environ.pop(k, None)
                os.environ.update(envs)
                res = {}

                def run():
                    res["r"] = fn(files, out=out)[0]
                eng.stats_reset(); eng.stats_enable(True)
                try:
                    ms = timed(stream, run, args.reps)
                except Exception as e:  # keep going: one broken variant must not cost the whole GPU call
                    print(json.dumps({"op": op, "variant": name, "error": str(e)}), flush=True)
                    continue
                eng.stats_enable(False)
                st = eng.stats()
                n_out = int(res["r"].shape[0])
                chk = int(res["r"].sum().item())
                if ref is None:
                    ref = (n_out, chk)
                print(json.dumps({"op": op, "variant": name, "env": envs, "ms": round(ms, 3), "kmers_in_per_s": total / ms * 1e3,
                                  "algo_GBps": round((total + n_out) * 8 / ms / 1e6, 1), "n_out": n_out, "same_as_first": (n_out, chk) == ref,
                                  "kernels": {k: {"launches": v["launches"], "ms_per_launch": round(v["ms"] / max(v["launches"], 1), 3)}
                                              for k, v in st.items()}}), flush=True)
    eng.close()


if __name__ == "__main__":
    main()
