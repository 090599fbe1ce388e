#!/usr/bin/env python
"""A/B of the union paths on the C3 inputs (8 sorted files x ~5e8 k-mers, one B200): the single-pass N-way
kernel in each tile shape (UKM_NWAY_CFG) against the two-way merge tree (UKM_NWAY=0), plus the host->device
copy rate the end-to-end number is bound by.  One JSON line per measurement."""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from unikmer_b200 import Engine  # noqa: E402


def timed(stream, fn, reps=3, warm=1):
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
    ap.add_argument("--cfgs", default="off,0,1,2,3,4")
    ap.add_argument("--h2d", action="store_true")
    ap.add_argument("--only", default="", help="union | inter: run just that part (ncu captures)")
    args = ap.parse_args()
    U = int(args.universe)
    eng = Engine(0)
    stream = torch.cuda.Stream()
    eng.use_stream(stream.cuda_stream)
    with torch.cuda.stream(stream):
        files = [eng.synth_member_file(0, U, U, 3, 4, f).clone() for f in range(8)]
        total = sum(int(f.shape[0]) for f in files)
        out = torch.empty(min(total, U) + 16, dtype=torch.int64, device="cuda")
        ref = None
        for cfg in ([] if args.only == "inter" else args.cfgs.split(",")):
            if cfg == "off":
                os.environ["UKM_NWAY"] = "0"
            else:
                os.environ["UKM_NWAY"] = "1"
                os.environ["UKM_NWAY_CFG"] = cfg
            res = {}

            def run():
                res["u"] = eng.union(files, out=out)[0]
            eng.stats_reset(); eng.stats_enable(True)
            ms = timed(stream, run)
            eng.stats_enable(False)
            st = eng.stats()
            n_out = int(res["u"].shape[0])
            chk = int(res["u"].sum().item())
            if ref is None:
                ref = (n_out, chk)
            print(json.dumps({"bench": "union8", "cfg": cfg, "ms": ms, "kmers_in_per_s": total / ms * 1e3,
                              "algo_GBps": (total + n_out) * 8 / ms / 1e6, "n_out": n_out, "same_as_first": (n_out, chk) == ref,
                              "kernels": {k: {"launches": v["launches"], "ms_per_launch": round(v["ms"] / max(v["launches"], 1), 3)}
                                          for k, v in st.items()}}), flush=True)
        # inter / diff: N-way hash filter against the file-by-file passes
        oi = torch.empty(int(files[0].shape[0]) + 16, dtype=torch.int64, device="cuda")
        if args.only == "union":
            eng.close()
            return
        for name, fn in ((("inter8", eng.inter),) if args.only == "inter" else (("inter8", eng.inter), ("diff8", eng.diff))):
            ref = None
            for cfg in args.cfgs.split(","):
                if cfg == "off":
                    os.environ["UKM_NWAY"] = "0"
                else:
                    os.environ["UKM_NWAY"] = "1"
                    os.environ["UKM_NWAY_FILTER"] = "1"
                    os.environ["UKM_NWAY_CFG"] = cfg
                res = {}

                def run():
                    res["r"] = fn(files, out=oi)[0]
                eng.stats_reset(); eng.stats_enable(True)
                ms = timed(stream, run)
                eng.stats_enable(False)
                st = eng.stats()
                n_out = int(res["r"].shape[0])
                chk = int(res["r"].sum().item())
                if ref is None:
                    ref = (n_out, chk)
                print(json.dumps({"bench": name, "cfg": cfg, "ms": ms, "kmers_in_per_s": total / ms * 1e3,
                                  "algo_GBps": (total + n_out) * 8 / ms / 1e6, "n_out": n_out, "same_as_first": (n_out, chk) == ref,
                                  "kernels": {k: {"launches": v["launches"], "ms_per_launch": round(v["ms"] / max(v["launches"], 1), 3)}
                                              for k, v in st.items()}}), flush=True)
        os.environ["UKM_NWAY"] = "1"
        for sub in (2, 4):
            os.environ["UKM_NWAY"] = "1"
            os.environ["UKM_NWAY_CFG"] = "0"
            ms = timed(stream, lambda: eng.union(files[:sub], out=out))
            tin = sum(int(f.shape[0]) for f in files[:sub])
            print(json.dumps({"bench": f"union{sub}", "cfg": "0", "ms": ms, "kmers_in_per_s": tin / ms * 1e3}), flush=True)
        if args.h2d:
            h = torch.empty(1 << 29, dtype=torch.int64, pin_memory=True)  # 4 GiB
            d = torch.empty_like(h, device="cuda")
            ms = timed(stream, lambda: d.copy_(h, non_blocking=True))
            print(json.dumps({"bench": "h2d_pinned_4GiB", "ms": ms, "GBps": h.numel() * 8 / ms / 1e6}), flush=True)
            ms = timed(stream, lambda: h.copy_(d, non_blocking=True))
            print(json.dumps({"bench": "d2h_pinned_4GiB", "ms": ms, "GBps": h.numel() * 8 / ms / 1e6}), flush=True)
    eng.close()


if __name__ == "__main__":
    main()
