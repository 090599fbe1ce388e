#!/usr/bin/env python
"""bench.py -- the headline benchmark: sorted uint64 k-mers/sec for union / inter / diff.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...

Workload (BASELINE.json configs[2], "C3" in SURVEY.md 8d): 8 sorted duplicate-free files of
~5e8 k=31 k-mers each -- universe U(j; 1e9, S=3), file f holds U_j iff bit f of sm64(4+j).
One STEP = `inter` + `diff` + `union` over the 8 files (each op consumes all ~4e9 input
k-mers); value = (3 * sum |F_i|) / step time, inputs resident in HBM.

A step is ONE call of the C ABI (ukm_setops_stream): the three results come out of one pass over
the inputs; the same step as three calls is reported beside it.

N > 1 (strong scaling: same total work): file f lives on rank f mod N; every rank owns a key range,
pulls its slice of the remote files over NVLink (CUDA-IPC peer copies on the copy engines, double
buffered, the plan made once: unikmer_b200/dist.py) while the kernels run on the slices that have
arrived, and makes the same single ABI call on them; results stay sharded in rank order.

--impl reference: the reference's CPU algorithms for the same path (hash-map union,
two-pointer inter/diff; oracle/oracle.c -- the Go reference cannot be built in this image)
on a bounded sample of the same workload, on the host cores.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

S_SEED, T_SEED = 3, 4
N_FILES = 8
METRIC = "sorted_uint64_kmers_per_sec_union_inter_diff"
UNIT = "k-mers/s"


def hbm_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "25",
                                          "-i", str(self.gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        top = sorted(sm)[len(sm) // 2:] if sm else []  # samples under load = upper half
        return {"sm_mhz": statistics.median(top) if top else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ---------------------------------------------------------------------------------------
# reference arm / cpu baseline: the oracle port on host cores, bounded sample
# ---------------------------------------------------------------------------------------
def cpu_pass(universe: int, threads: int):
    """One inter + diff + union over the 8 files of a `universe`-sized sample, reference algorithms."""
    import oracle
    files = [oracle.member_file(0, universe, universe, S_SEED, T_SEED, f) for f in range(N_FILES)]
    total = sum(len(f) for f in files)
    t0 = time.perf_counter()
    i, _ = oracle.inter(files)
    t1 = time.perf_counter()
    d, _ = oracle.diff(files, threads=threads)
    t2 = time.perf_counter()
    u, _ = oracle.union(files, threads=threads)
    t3 = time.perf_counter()
    return total, (t1 - t0, t2 - t1, t3 - t2), (len(i), len(d), len(u))


def cpu_baseline(sample_universe: int):
    threads = os.cpu_count() or 1
    # the port's speed at three sample sizes brackets how it moves with the size (the GPU arm runs universe 1e9: the
    # hash-map union only gets slower as its table outgrows the caches, the two-pointer loops stay flat)
    by_size = {}
    for uni in (sample_universe // 4, sample_universe // 2):
        t_, (a_, b_, c_), _ = cpu_pass(uni, threads)
        by_size[f"{uni:.1e}"] = {"kmers_per_s": 3 * t_ / (a_ + b_ + c_), "inter": t_ / a_, "diff": t_ / b_, "union": t_ / c_}
    total, (ti, td, tu), sizes = cpu_pass(sample_universe, threads)
    secs = ti + td + tu
    by_size[f"{sample_universe:.1e}"] = {"kmers_per_s": 3 * total / secs, "inter": total / ti, "diff": total / td, "union": total / tu}
    return {"value": 3 * total / secs, "unit": UNIT, "cores": threads, "kind": "port", "by_sample_universe": by_size,
            "sample": f"C3 scaled to universe {sample_universe:.0e} (8 files x ~{sample_universe // 2:.1e} k-mers): one inter+diff+union pass, "
                      f"{secs:.1f} s (inter {ti:.2f} s, diff {td:.2f} s, union {tu:.2f} s); C restatement of the Go algorithms "
                      f"(hash-map union, two-pointer inter/diff; single goroutine each, the sort in union/diff uses {threads} threads)",
            "per_op_kmers_per_s": {"inter": total / ti, "diff": total / td, "union": total / tu}}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    uni = args.ref_universe
    vals, last = [], None
    for it in range(args.warmup + args.steps):
        t0 = time.perf_counter()
        total, ts, sizes = cpu_pass(uni, os.cpu_count() or 1)
        dt = sum(ts)
        if it >= args.warmup:
            vals.append((3 * total / dt, dt))
        last = (total, ts, sizes)
    value = statistics.mean(v for v, _ in vals)
    ms = 1e3 * statistics.mean(d for _, d in vals)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "u64", "data": "synthetic",
        "config": {"workload": f"C3 sample: inter+diff+union over 8 sorted files, universe {uni:.0e} (~{uni // 2:.1e} k-mers per file), k=31",
                   "note": "Go reference cannot be built here (no Go toolchain, un-vendored modules): C restatement of its algorithms (oracle/oracle.c)"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": os.cpu_count() or 1, "kind": "port",
                         "sample": f"universe {uni:.0e}; per step inter {last[1][0]:.2f} s, diff {last[1][1]:.2f} s, union {last[1][2]:.2f} s"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


# ---------------------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--universe", type=float, default=1e9, help="universe size N (files hold ~N/2 k-mers each)")
    ap.add_argument("--ref-universe", type=float, default=0,
                    help="sample universe for the CPU arm; default: 8e7 (about 22 s per step on 16 host threads), less when "
                         "steps + warmup would not fit about 520 s at that size")
    ap.add_argument("--cpu-universe", type=float, default=4e7, help="largest sample universe for the cpu_baseline object (also run at 1/2 and 1/4)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--e2e-steps", type=int, default=2)
    ap.add_argument("--exchange-chunks", type=int, default=1, help="N>1: pieces of a rank's key range; piece c+1 is pulled over NVLink while piece c is computed")
    ap.add_argument("--exchange", default="peer", choices=["peer", "nccl"], help="N>1: NVLink peer pulls (copy engines, overlapped) or one NCCL all-to-all-v")
    args = ap.parse_args()
    if not args.ref_universe:
        # a pass costs ~24 s of wall time per 8e7 of universe on the GPU box's host (measured: 22.2 s timed + generation),
        # about linearly; keep the whole --steps K --warmup W run near 520 s (the driver's budget per run is 870 s for both arms)
        per_pass = 520.0 / max(1, args.steps + args.warmup)
        args.ref_universe = max(1e7, min(8e7, 8e7 * per_pass / 24.0))
        args.ref_universe = float(int(args.ref_universe / 1e6) * 1e6)
    args.universe, args.ref_universe, args.cpu_universe = int(args.universe), int(args.ref_universe), int(args.cpu_universe)
    if args.impl == "reference":
        return run_reference(args)

    # Use the first N GPUs of the box and nothing else (at N = 2 / 4 on an 8-GPU box the other GPUs showed
    # activity -- contexts created by peer-access probing); a launcher that already restricts the devices is left alone.
    if "CUDA_VISIBLE_DEVICES" not in os.environ:
        n_local = int(os.environ.get("LOCAL_WORLD_SIZE", "0") or 0) or max(1, args.gpus)
        os.environ["CUDA_VISIBLE_DEVICES"] = ",".join(str(i) for i in range(n_local))

    import torch
    import torch.distributed as dist

    from unikmer_b200 import Engine
    from unikmer_b200.dist import KeyRangeExchange, PeerPullExchange, equal_width_splitters, owner_of_file

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    dev = torch.device(f"cuda:{local}")
    eng = Engine(local)
    stream = torch.cuda.Stream(device=dev)
    eng.use_stream(stream.cuda_stream)
    U = args.universe

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    with torch.cuda.stream(stream):
        # ---- inputs: generated on the device, resident in HBM before anything is timed ----
        local_files = {}
        for f in range(N_FILES):
            if owner_of_file(f, world) == rank:
                local_files[f] = eng.synth_member_file(0, U, U, S_SEED, T_SEED, f).clone()
        torch.cuda.synchronize()
        sizes = torch.zeros(N_FILES, dtype=torch.int64, device=dev)
        for f, t in local_files.items():
            sizes[f] = t.shape[0]
        if world > 1:
            dist.all_reduce(sizes)
        total_in = int(sizes.sum().item())
        splitters = equal_width_splitters(world, 62)
        ex = KeyRangeExchange(eng, rank, world)
        pex = None
        if world > 1 and args.exchange == "peer":
            try:
                pex = PeerPullExchange(eng, rank, world, local_files, N_FILES)
            except Exception as e:  # no peer access / IPC: fall back to the NCCL all-to-all-v
                if rank == 0:
                    print(f"peer exchange unavailable ({e}); using NCCL", file=sys.stderr)
                pex = None
            ok = torch.tensor([1 if pex is not None else 0], device=dev)
            dist.all_reduce(ok, op=dist.ReduceOp.MIN)
            if int(ok.item()) == 0:
                pex = None

        if pex is not None:
            # the plan of the pipelined exchange is made ONCE (the resident files do not change between steps): where every
            # file is cut at the G * K piece boundaries; per step no collective and no host round trip is needed for it
            K = max(1, args.exchange_chunks)
            pex.plan_chunks(splitters, K)
            cb = pex.chunk_bounds
            n0_rank = int(cb[0, (rank + 1) * K] - cb[0, rank * K])
            nall_rank = int((cb[:, (rank + 1) * K] - cb[:, rank * K]).sum())
            oi = torch.empty(n0_rank + 16, dtype=torch.int64, device=dev)
            od = torch.empty(n0_rank + 16, dtype=torch.int64, device=dev)
            ou = torch.empty(nall_rank + 16, dtype=torch.int64, device=dev)

        OPS = ("inter", "diff", "union")
        if world == 1:
            oi = torch.empty(int(sizes[0]) + 16, dtype=torch.int64, device=dev)
            od = torch.empty(int(sizes[0]) + 16, dtype=torch.int64, device=dev)
            # capacity of a union's output span = the sum of the input sizes (the ABI's worst-case contract): with less the
            # library merges into a temporary and copies the result over
            ou = torch.empty(total_in + 16, dtype=torch.int64, device=dev)

        def step():
            # ONE call of the C ABI per step (ukm_setops_stream on device spans): the single-pass N-way union kernel sees, in
            # its last merge level, how many files hold every key -- which is all inter (every file) and diff (file 0 only)
            # need -- so all three results come out of one pass over the inputs; all three are written in full every step
            if world == 1:
                return tuple(eng.setops([local_files[f] for f in range(N_FILES)], OPS, outs=[oi, od, ou]))
            if pex is None:
                files = ex.exchange(local_files, N_FILES, splitters)
                return tuple(eng.setops(files, OPS, shard=True))
            # The rank's key range in K pieces: the copy engines pull the next piece of every remote file over NVLink (peer
            # mappings, no SMs, no NCCL kernels) while the kernels run on the current one; results are written piece after
            # piece into the rank's output buffers (key order = piece order).
            wi = wd = wu = 0
            for slices, ev in pex.exchange_chunks():
                pex.wait(ev, range(N_FILES))
                a, b, c = eng.setops(slices, OPS, outs=[oi[wi:], od[wd:], ou[wu:]], shard=True)
                wi, wd, wu = wi + a.shape[0], wd + b.shape[0], wu + c.shape[0]
            return oi[:wi], od[:wd], ou[:wu]

        def step_separate_calls():
            # the same step as three ABI calls (ukm_inter, ukm_diff, ukm_union), each reading every input
            files = [local_files[f] for f in range(N_FILES)]
            return eng.inter(files, out=oi)[0], eng.diff(files, out=od)[0], eng.union(files, out=ou)[0]

        # clocks / throttle reasons are sampled from before the warm-up until the end of the timed region (nvidia-smi
        # needs ~0.1 s before its first line; at N = 8 the timed region itself is shorter than that)
        sampler = ClockSampler(local)
        if rank == 0:
            sampler.start()
        res = None
        for _ in range(args.warmup):
            res = step()
        del res
        eng.stats_reset()
        eng.stats_enable(True)
        launches0 = eng.launch_count()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(args.steps):
            res = step()
        e1.record(stream)
        barrier()
        clocks = sampler.stop() if rank == 0 else None
        ms_total = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(ms_total, op=dist.ReduceOp.MAX)
        ms_step = float(ms_total.item()) / args.steps
        launches = eng.launch_count() - launches0
        eng.stats_enable(False)
        stats = eng.stats()
        separate = None
        if world == 1:
            step_separate_calls()
            torch.cuda.synchronize()
            e0.record(stream)
            for _ in range(3):
                step_separate_calls()
            e1.record(stream)
            torch.cuda.synchronize()
            ms_sep = e0.elapsed_time(e1) / 3
            separate = {"ms_per_step": ms_sep, "value": 3.0 * total_in / (ms_sep * 1e-3),
                        "note": "the same step as three ABI calls (ukm_inter, ukm_diff, ukm_union), each reading every input"}
            res = step()  # the results checked below are those of the timed path
        value = 3.0 * total_in / (ms_step * 1e-3)

        # ---- result checks ----
        # (1) FULL results: {count, sum mod 2^64, xor} of every result on the device (summed / xor-ed over the ranks) against
        #     the digests the generator's membership bits give for the whole universe (oracle.c3_digest: no set operation
        #     involved, 1e9 universe keys in about a second on the host);
        # (2) an exact key window of every result against the oracle's set operations.
        inter, diff, union = res

        def dev_digest(t):
            if t.shape[0] == 0:
                return [0, 0, 0]
            x = t.view(torch.int64)
            xo = x
            while xo.shape[0] > 1:  # xor-reduce by halving (torch has no bitwise reduction)
                h = xo.shape[0] // 2
                y = xo[:h] ^ xo[h:2 * h]
                xo = torch.cat([y, xo[2 * h:]]) if xo.shape[0] % 2 else y
            return [int(t.shape[0]), int(x.sum().item()) & (2**64 - 1), int(xo.item()) & (2**64 - 1)]
        dg = torch.tensor([[v - 2**64 if v >= 2**63 else v for v in dev_digest(t)] for t in (inter, diff, union)],
                          dtype=torch.int64, device=dev)
        if world > 1:
            cs = dg[:, :2].clone()
            dist.all_reduce(cs)  # count and sum add up (mod 2^64: int64 wraps)
            xs = [torch.zeros_like(dg[:, 2]) for _ in range(world)]
            dist.all_gather(xs, dg[:, 2].contiguous())
            xr = xs[0]
            for t in xs[1:]:
                xr = xr ^ t
            dg = torch.cat([cs, xr[:, None]], dim=1)
        got = [[int(v) & (2**64 - 1) for v in row] for row in dg.tolist()]
        n_inter, n_diff, n_union = got[0][0], got[1][0], got[2][0]
        check = {"n_inter": n_inter, "n_diff": n_diff, "n_union": n_union}
        if rank == 0:
            import oracle
            exp = oracle.c3_digest(0, U, U, S_SEED, T_SEED, N_FILES)
            for name, g in zip(("inter", "diff", "union"), got):
                ok = tuple(g) == tuple(exp[name])
                check[f"{name}_full_digest_exact"] = ok
                if not ok:
                    raise SystemExit(f"bench self-check failed: {name} digest {g} != expected {exp[name]}")
            w = 2_000_000  # j-window [0, w): keys below U(w) -- exact parity of that window
            W = (1 << 62) // U
            bound = w * W
            ofiles = [oracle.member_file(0, w, U, S_SEED, T_SEED, f) for f in range(N_FILES)]
            for name, got_t, exp_k in (("inter", inter, oracle.inter(ofiles)[0]), ("diff", diff, oracle.diff(ofiles)[0]),
                                       ("union", union, oracle.union(ofiles)[0])):
                g = got_t[: len(exp_k) + 8].cpu().numpy().view(np.uint64)
                g = g[g < bound]
                ok = len(g) == len(exp_k) and bool(np.array_equal(g, exp_k))
                check[f"{name}_window_exact"] = ok
                if not ok:
                    raise SystemExit(f"bench self-check failed: {name} window differs from the oracle")

        # ---- roofline of the dominant kernel family, from CUDA events on the stream (library-side ukm_stats) ----
        # algorithmic bytes: every input key read once + every output key written once per operation
        # (per two-way pass for the chained inter / diff, whose passes are separate launches)
        peak, peak_src = hbm_peak()
        so = {k: v for k, v in stats.items() if k.startswith("setop_")}
        so_ms = sum(v["ms"] for v in so.values())
        so_bytes = sum(v["algo_bytes"] for v in so.values())
        dom_name, dom = max(so.items(), key=lambda kv: kv[1]["ms"]) if so else ("none", {"ms": 0.0, "algo_bytes": 0.0, "launches": 0})
        achieved = dom["algo_bytes"] / (dom["ms"] * 1e-3) / 1e9 if dom["ms"] else 0.0
        traffic = None
        prof = os.path.join(ROOT, "profiles", "setop_ncu_traffic.json")
        if os.path.exists(prof):
            try:
                traffic = json.load(open(prof)).get(dom_name, {}).get("dram_bytes_per_launch")
            except Exception:
                traffic = None
        kernel_of = {"setop_union_nway": "nway_kernel<UNION> (single-pass 8-way union: TMA tile loads, in-smem merge levels) + its partition",
                     "setop_inter_nway": "nfilter_kernel<INTER> (single-pass 8-way filter over file-0 chunks) + partition + mask gather",
                     "setop_diff_nway": "nfilter_kernel<DIFF> (single-pass 8-way filter over file-0 chunks) + partition + mask gather",
                     "setop_inter_diff_union_nway": "nway_kernel<UNION> with inter / diff riding along (ONE pass for the three results: the last "
                                                    "merge level sees how many files hold each key) + its partition",
                     "setop_mask_gather": "inter / diff masks -> compact results (count / scan / gather)",
                     "setop_inter_diff_nway": "nfilter_kernel<BOTH> (inter AND diff in one pass over file-0 chunks) + partition + mask gathers",
                     "setop_union": "setop_pipe_kernel<UNION> (two-way merge-path passes)",
                     "setop_inter": "setop_pipe_kernel<INTER> + setop_search_kernel (two-way passes in file order)",
                     "setop_diff": "setop_pipe_kernel<DIFF> + setop_search_kernel (two-way passes in file order)"}
        # per OPERATION on SURVEY.md 8(d) bytes (every input key read once + every output key written once per operation):
        # with the single-pass kernels an operation is one stats family, so these are also the per-kernel numbers
        per_op = {k: {"ms_per_op": v["ms"] / v["launches"], "algo_GB_per_op": v["algo_bytes"] / v["launches"] / 1e9,
                      "achieved_GBps": v["algo_bytes"] / (v["ms"] * 1e-3) / 1e9, "frac": v["algo_bytes"] / (v["ms"] * 1e-3) / 1e9 / peak}
                  for k, v in so.items() if v["ms"] and v["launches"]}
        # bytes the dominant launch really moves (DRAM): inputs once + the union's output for the fused call; = algorithmic otherwise
        moved = None
        if dom_name == "setop_inter_diff_union_nway" and dom["launches"]:
            moved = (dom["algo_bytes"] / dom["launches"]) - 2.0 * 8.0 * (total_in / world if world > 1 else total_in)
        roofline = {"bound": "hbm", "kernel": kernel_of.get(dom_name, dom_name), "family": dom_name,
                    "algorithmic_bytes_definition": "SURVEY.md 8(d), per operation: every input key read once + every output key written once; "
                                                    "one launch of the fused kernel does three operations (union, inter, diff)",
                    "bytes_moved_per_launch": moved,
                    "achieved_on_bytes_moved": (moved / (dom["ms"] / dom["launches"] * 1e-3) / 1e9) if moved else None,
                    "frac_on_bytes_moved": (moved / (dom["ms"] / dom["launches"] * 1e-3) / 1e9 / peak) if moved and peak else None,
                    "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak if peak else None,
                    "traffic": traffic, "peak_source": peak_src, "launches": dom["launches"],
                    "avg_launch_ms": dom["ms"] / dom["launches"] if dom["launches"] else None,
                    "algo_bytes_per_launch": dom["algo_bytes"] / dom["launches"] if dom["launches"] else None,
                    "share_of_step": dom["ms"] / (ms_step * args.steps) if ms_step else None,
                    "per_operation": per_op,
                    "all_setop_kernels": {"achieved": so_bytes / (so_ms * 1e-3) / 1e9 if so_ms else 0.0,
                                          "frac": (so_bytes / (so_ms * 1e-3) / 1e9 / peak) if so_ms and peak else None,
                                          "share_of_step": so_ms / (ms_step * args.steps) if ms_step else None}}

        # ---- e2e: same step through the C ABI with HOST buffers (pinned), H2D + D2H in the timed region ----
        # ---- N > 1: how fast the copy engines pull a step's remote slices, under the kernels and alone ----
        exchange = None
        if pex is not None:
            torch.cuda.synchronize()
            barrier()
            pex.time_pulls, pex.pull_marks = True, []
            step()  # the pulls recorded are those of the NEXT step's pieces, queued while this step's kernels run
            loaded = pex.pull_timing()
            barrier()
            pex.pull_marks = []
            gen = pex.exchange_chunks()
            next(gen)  # hands out the piece pulled above and queues the next pulls: nothing else runs on the GPU
            idle = pex.pull_timing()
            gen.close()
            barrier()
            pex.time_pulls, pex.pull_marks = False, []
            exchange = {"copy_streams": len(pex.streams), "pieces_per_slice": pex.split, "under_kernels": loaded, "gpu_idle": idle,
                        "note": "rank 0's peer copies of one step (CUDA events around every copy); span_ms = first start to last end"}
            for d in (loaded, idle):
                if d:
                    d["GBps_over_span"] = d["bytes"] / max(d["span_ms"], 1e-9) / 1e6

        e2e = None
        if not args.no_e2e:
            hfiles = {f: torch.empty(t.shape[0], dtype=torch.int64, pin_memory=True) for f, t in local_files.items()}
            for f, t in local_files.items():
                hfiles[f].copy_(t)
            torch.cuda.synchronize()
            extra = {}
            if world == 1:
                order = [hfiles[f] for f in range(N_FILES)]
                local_files.clear()  # free the device copies: the e2e path starts from host memory
                torch.cuda.empty_cache()
                ho_i = torch.empty(order[0].shape[0], dtype=torch.int64, pin_memory=True)
                ho_d = torch.empty(order[0].shape[0], dtype=torch.int64, pin_memory=True)
                ho_u = torch.empty(min(total_in, U) + 16, dtype=torch.int64, pin_memory=True)

                def e2e_step():
                    # ONE call of the C ABI (ukm_setops_stream) with host spans in, host spans out: the library cuts the key
                    # space into ranges and overlaps the upload of range c+1, the three operations on range c and the download
                    # of range c-1; every input byte crosses PCIe once per step
                    r = eng.setops(order, ("inter", "diff", "union"), outs=[ho_i, ho_d, ho_u])
                    return tuple(int(x.shape[0]) for x in r)

                def e2e_step_whole():
                    # the step's inputs cross PCIe once (ukm_copy into device spans), the three operations chain on the
                    # device copies, every result is delivered into pinned host memory: nothing overlaps
                    d = [eng.upload(h) for h in order]
                    a = eng.inter(d, out=ho_i)[0]
                    b = eng.diff(d, out=ho_d)[0]
                    c = eng.union(d, out=ho_u)[0]
                    return a.shape[0], b.shape[0], c.shape[0]

                def e2e_step_host_spans():
                    # three separate ABI calls, each handed the HOST spans: each streams its own upload (3 x 32 GB H2D)
                    a = eng.inter(order, out=ho_i)[0]
                    b = eng.diff(order, out=ho_d)[0]
                    c = eng.union(order, out=ho_u)[0]
                    return a.shape[0], b.shape[0], c.shape[0]
                path = ("ONE C-ABI call (ukm_setops_stream) from pinned HOST spans to pinned HOST spans: inter + diff + union streamed as "
                        "key ranges inside the library (upload of range c+1 | kernels on range c | download of range c-1); every input "
                        "byte crosses PCIe once per step")
            else:
                ho = []  # pinned host buffers for the rank's results, sized on the first (warm-up) call
                if pex is not None:
                    pex.prefetch_next_step = False  # an end-to-end step pulls nothing before its own upload has finished
                    pex.chunk_prefetched = None

                def e2e_step():
                    # every rank uploads the files it owns into its resident buffers, then the pipelined NVLink exchange +
                    # the three operations per piece (the same step() as above), then the rank's results go back to the host
                    for f, h in hfiles.items():
                        local_files[f].copy_(h, non_blocking=True)
                    torch.cuda.current_stream().synchronize()
                    dist.barrier()  # a peer must not pull a slice its owner is still uploading
                    outs = step()
                    if not ho:
                        ho.extend(torch.empty(o.shape[0] + 16, dtype=torch.int64, pin_memory=True) for o in outs)
                    for o, h in zip(outs, ho):
                        h[:o.shape[0]].copy_(o, non_blocking=True)
                    torch.cuda.current_stream().synchronize()
                    return tuple(int(o.shape[0]) for o in outs)
                path = ("per rank: H2D of the files it owns from pinned host memory -> barrier -> pipelined NVLink peer-pull exchange + "
                        "inter/diff/union per piece -> D2H of the rank's results into pinned host memory")
            h2d = total_in * 8

            def timed(fn, reps):
                fn()  # warm-up
                barrier()
                t0 = time.perf_counter()
                for _ in range(reps):
                    ns_ = fn()
                barrier()
                wall = (time.perf_counter() - t0) / reps
                wt = torch.tensor([wall], dtype=torch.float64, device=dev)
                nt = torch.tensor(list(ns_), dtype=torch.int64, device=dev)
                if world > 1:
                    dist.all_reduce(wt, op=dist.ReduceOp.MAX)
                    dist.all_reduce(nt)
                assert nt.tolist() == [n_inter, n_diff, n_union], "e2e results differ from the device-resident run"
                return float(wt.item())
            w1 = timed(e2e_step, args.e2e_steps)
            d2h = (n_inter + n_diff + n_union) * 8
            e2e = {"value": 3.0 * total_in / w1, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                   "ms_per_step": 1e3 * w1, "steps": args.e2e_steps, "path": path}
            if world == 1:
                # the result of the single call, checked as a whole: digests of the pinned host outputs
                import oracle
                exp = oracle.c3_digest(0, U, U, S_SEED, T_SEED, N_FILES)
                for name, h, n in (("inter", ho_i, n_inter), ("diff", ho_d, n_diff), ("union", ho_u, n_union)):
                    assert oracle.digest3(h[:n].numpy().view(np.uint64)) == tuple(exp[name]), f"e2e {name} digest differs"
                e2e["host_outputs_digest_exact"] = True
                w2 = timed(e2e_step_whole, 1)
                e2e["whole_files_uploaded_then_computed"] = {
                    "value": 3.0 * total_in / w2, "ms_per_step": 1e3 * w2,
                    "path": "ukm_copy of the 8 files to the device once per step, ukm_inter/ukm_diff/ukm_union on the device spans, "
                            "results delivered into pinned host buffers (no overlap)"}
                w3 = timed(e2e_step_host_spans, 1)
                e2e["host_spans_every_call"] = {"value": 3.0 * total_in / w3, "unit": UNIT, "ms_per_step": 1e3 * w3,
                                                "h2d_bytes_per_step": 3 * total_in * 8, "d2h_bytes_per_step": d2h,
                                                "path": "ukm_inter, ukm_diff, ukm_union called one after the other with HOST spans: each "
                                                        "call streams its own upload of all inputs"}

        cpu = None
        if rank == 0 and not args.no_cpu:
            cpu = cpu_baseline(args.cpu_universe)

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "u64",
            "data": "synthetic",
            "config": {"workload": f"C3: inter+diff+union over 8 sorted duplicate-free files x ~{U // 2:.1e} k=31 uint64 k-mers "
                                   f"(universe {U:.0e}, {total_in} k-mers in); each op reads all inputs",
                       "inputs": "device-resident, 32 GB >> 126 MB L2 (no L2 flush needed)" if U >= 10**8 else "device-resident",
                       "step": "one C-ABI call (ukm_setops_stream, device spans) per step and rank: union, inter and diff from ONE pass over "
                               "the inputs (single-pass 8-way merge; inter / diff from the run lengths of its last level); all three "
                               "results written in full; the same step as three separate calls is reported in three_separate_calls",
                       "parallelism": ("1 GPU" if world == 1 else f"key-range shards x{world}, " +
                                       (f"NVLink peer pulls on the copy engines (CUDA IPC), {args.exchange_chunks} piece(s) per rank and step, double-buffered: "
                                        "the next piece (the next step's first piece after the last one) is pulled while the single-pass N-way "
                                        "kernels run on the current one; the exchange plan is made once"
                                        if pex is not None else "one NCCL all-to-all-v per step")),
                       "kmers_per_step": 3 * total_in},
            "roofline": roofline, "three_separate_calls": separate, "cpu_baseline": cpu, "e2e": e2e, "exchange": exchange, "gpu_launches": int(launches), "clocks": clocks,
            "check": check,
            "per_kernel": {k: {"launches": v["launches"], "ms": round(v["ms"], 3),
                               "GBps": round(v["algo_bytes"] / (v["ms"] * 1e-3) / 1e9, 1) if v["ms"] else None}
                           for k, v in stats.items()},
        }
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


def _quiet_stdout():
    """Libraries (NCCL's version banner) print to fd 1; the contract is ONE JSON line on stdout.  Point fd 1 at
    stderr for the run and keep the real stdout for the result line."""
    sys.stdout.flush()
    real = os.dup(1)
    os.dup2(2, 1)
    return os.fdopen(real, "w")


if __name__ == "__main__":
    _out = _quiet_stdout()
    _print = print

    def print(*a, **k):  # noqa: A001 -- result lines go to the real stdout
        k.setdefault("file", _out)
        _print(*a, **k)
        _out.flush()
    sys.exit(main())
