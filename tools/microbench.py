#!/usr/bin/env python
"""Per-kernel micro-benchmarks on one B200 (not the headline bench): radix sort configs, pair sort,
k-mer/ntHash generation, count pipeline, folds.  Prints one JSON line per measurement."""
import argparse
import json
import os
import sys

import numpy as np
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
    ap.add_argument("--n", type=float, default=1e9)
    ap.add_argument("--what", default="sort,pairs,kmers,count,fold")
    args = ap.parse_args()
    n = int(args.n)
    eng = Engine(0)
    stream = torch.cuda.Stream()
    eng.use_stream(stream.cuda_stream)
    what = args.what.split(",")
    with torch.cuda.stream(stream):
        if "sort" in what or "sort1" in what:
            src = eng.synth_random_keys(0, n, 2)
            work = torch.empty_like(src)
            variants = [("2", "ballot")] if "sort1" in what else [(c, m) for m in ("any", "ballot") for c in ("0", "1", "2", "3")]
            for cfg, match in variants:
                os.environ["UKM_SORT_CFG"] = cfg
                os.environ["UKM_SORT_MATCH"] = match

                def run():
                    work.copy_(src)
                    eng.sort(work, key_bits=62)
                copy_ms = timed(stream, lambda: work.copy_(src))
                eng.stats_reset(); eng.stats_enable(True)
                ms = timed(stream, run) - copy_ms
                eng.stats_enable(False)
                st = eng.stats()
                ok = bool((work[1:] >= work[:-1]).all().item())
                print(json.dumps({"bench": "sort_u64", "n": n, "cfg": cfg, "match": match, "ms": ms, "keys_per_s": n / ms * 1e3,
                                  "GBps_136B": 136 * n / ms / 1e6, "GBps_actual": (1 + 2 * 8) * 8 * n / ms / 1e6, "sorted": ok,
                                  "kernels": {k: round(v["ms"] / max(v["launches"], 1), 3) for k, v in st.items()}}), flush=True)
            os.environ.pop("UKM_SORT_CFG", None)
            os.environ.pop("UKM_SORT_MATCH", None)
            del src, work
        if "cub" in what:
            # YARDSTICK ONLY: cub::DeviceRadixSort on the same keys (tools/yardstick/libcubyard.so, never part of libukm.so)
            import ctypes as C
            yl = C.CDLL(os.path.join(os.path.dirname(os.path.abspath(__file__)), "yardstick", "libcubyard.so"))
            yl.cub_sort_u64.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_size_t,
                                        C.POINTER(C.c_size_t)]
            src = eng.synth_random_keys(0, n, 2)
            dst = torch.empty_like(src)
            need = C.c_size_t(0)
            assert yl.cub_sort_u64(src.data_ptr(), dst.data_ptr(), n, 0, 62, stream.cuda_stream, None, 0, C.byref(need)) == 0
            temp = torch.empty(need.value + 256, dtype=torch.uint8, device="cuda")

            def run_cub():
                r = yl.cub_sort_u64(src.data_ptr(), dst.data_ptr(), n, 0, 62, stream.cuda_stream, temp.data_ptr(), temp.numel(), None)
                assert r == 0, r
            ms = timed(stream, run_cub)
            ok = bool((dst[1:] >= dst[:-1]).all().item())
            print(json.dumps({"bench": "cub_DeviceRadixSort_u64_YARDSTICK", "n": n, "bits": 62, "ms": ms, "keys_per_s": n / ms * 1e3,
                              "GBps_136B": 136 * n / ms / 1e6, "sorted": ok, "temp_bytes": need.value,
                              "note": "library sort, out of place, no input copy; comparison only, never on the product path"}), flush=True)
            del src, dst, temp
        if "setops" in what:
            U = n
            files = [eng.synth_member_file(0, U, U, 3, 4, f).clone() for f in range(8)]
            tot = sum(int(f.shape[0]) for f in files)
            for pipe, vt, skew in [("0", "15", "3"), ("1", "15", "3"), ("2", "15", "3"), ("3", "15", "3"), ("4", "15", "3"), ("5", "15", "3")]:
                os.environ["UKM_SETOP_PIPE"], os.environ["UKM_SETOP_VT"], os.environ["UKM_SETOP_SKEW"] = pipe, vt, skew
                res = {}
                for name in ("inter", "diff", "union"):
                    fn = getattr(eng, name)
                    eng.stats_reset(); eng.stats_enable(True)
                    ms = timed(stream, lambda: fn(files), reps=3)
                    eng.stats_enable(False)
                    st = eng.stats()
                    res[name] = {"ms": round(ms, 3), "kmers_per_s": tot / ms * 1e3,
                                 "launch_ms": [round(v["ms"] / 4, 3) for k, v in st.items() if k.startswith("setop")]}
                print(json.dumps({"bench": "setops_C3", "universe": U, "pipe": pipe, "vt": vt, "skew": skew, **res}), flush=True)
            os.environ.pop("UKM_SETOP_VT", None); os.environ.pop("UKM_SETOP_SKEW", None); os.environ.pop("UKM_SETOP_PIPE", None)
            del files
        if "pairs" in what:
            m = n // 2
            src = eng.synth_random_keys(0, m, 2)
            vals = torch.arange(m, dtype=torch.int32, device="cuda")
            wk, wv = torch.empty_like(src), torch.empty_like(vals)

            def run():
                wk.copy_(src); wv.copy_(vals)
                eng.sort(wk, wv, key_bits=62)
            copy_ms = timed(stream, lambda: (wk.copy_(src), wv.copy_(vals)))
            ms = timed(stream, run) - copy_ms
            print(json.dumps({"bench": "sort_pairs", "n": m, "ms": ms, "keys_per_s": m / ms * 1e3}), flush=True)
            del src, vals, wk, wv
        if "kmers" in what or "count" in what or "minimizer" in what:
            L = min(n, 10**9)
            nrec = 10
            bases = torch.cat([eng.synth_bases(r, 0, L // nrec, 5) for r in range(nrec)])
            off = torch.arange(0, nrec + 1, dtype=torch.int64, device="cuda") * (L // nrec)
            for hashed in (True, False):
                if "kmers" in what:
                    eng.stats_reset(); eng.stats_enable(True)
                    ms = timed(stream, lambda: eng.kmers(bases, off, 31, canonical=True, hashed=hashed), reps=2)
                    eng.stats_enable(False)
                    st = eng.stats()
                    kk = [v for k, v in st.items() if k.startswith("kmer_")][0]
                    kms = kk["ms"] / kk["launches"]
                    print(json.dumps({"bench": "kmers", "hashed": hashed, "bases": L, "call_ms": ms, "kernel_ms": kms,
                                      "kernel_GBps": 9 * L / kms / 1e6, "kmers_per_s": L / kms * 1e3}), flush=True)
            if "count" in what:
                eng.stats_reset(); eng.stats_enable(True)
                ms = timed(stream, lambda: eng.count(bases, off, 31, canonical=True, hashed=True), reps=1)
                eng.stats_enable(False)
                print(json.dumps({"bench": "count_k31_KH", "bases": L, "ms": ms, "kmers_per_s": L / ms * 1e3,
                                  "kernels": {k: round(v["ms"], 2) for k, v in eng.stats().items()}}), flush=True)
            if "minimizer" in what:
                eng.stats_reset(); eng.stats_enable(True)
                ms = timed(stream, lambda: eng.count_minimizer(bases, off, 31, 15, canonical=True), reps=1)
                eng.stats_enable(False)
                print(json.dumps({"bench": "count_minimizer_k31_w15", "bases": L, "ms": ms, "kmers_per_s": L / ms * 1e3,
                                  "kernels": {k: round(v["ms"], 2) for k, v in eng.stats().items()}}), flush=True)
            del bases
        if "fold" in what:
            m = n // 2
            keys = eng.synth_random_keys(0, m, 2)
            eng.sort(keys, key_bits=62)
            eng.stats_reset(); eng.stats_enable(True)
            ms = timed(stream, lambda: eng.fold(1, keys), reps=2)
            eng.stats_enable(False)
            st = eng.stats()["fold"]
            kms = st["ms"] / st["launches"]
            print(json.dumps({"bench": "fold_unique", "n": m, "call_ms": ms, "kernel_ms": kms, "kernel_GBps": 16 * m / kms / 1e6}), flush=True)


if __name__ == "__main__":
    main()
