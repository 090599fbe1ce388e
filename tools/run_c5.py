#!/usr/bin/env python
"""C5 (BASELINE.json configs[4]): `common -n 32` with TaxId LCA over 64 files x ~1e8 k-mers, key-range sharded
across the GPUs of one box.  Launch with torchrun (one rank per GPU) or plain python for one GPU:

    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 tools/run_c5.py [--universe 2e8]

Generator: universe U(j; N, S=6); file f holds U_j iff bit f of sm64(7+j); taxonomy: 1e4-node synthetic
tree parent[t] = 1 + sm64(8+t) % (t-1); file f carries the GLOBAL taxid leaf_f = 10000 - f (README.md:169-171 workflow).
Checks an exact key window against the CPU oracle on rank 0 and prints one JSON line."""
import argparse
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from unikmer_b200 import Engine, KmerSet  # noqa: E402
from unikmer_b200.dist import KeyRangeExchange, equal_width_splitters, owner_of_file  # noqa: E402

S, T, Q, NFILES, NTAX = 6, 7, 8, 64, 10_000


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--universe", type=float, default=2e8)
    ap.add_argument("--threshold", type=int, default=32)
    ap.add_argument("--steps", type=int, default=2)
    args = ap.parse_args()
    U = int(args.universe)
    rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", "0"), ("WORLD_SIZE", "1"), ("LOCAL_RANK", "0")))
    torch.cuda.set_device(local)
    dev = torch.device(f"cuda:{local}")
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    import oracle
    parent = np.zeros(NTAX + 1, dtype=np.uint32)
    parent[1] = 1
    for t in range(2, NTAX + 1):
        parent[t] = 1 + oracle.sm64(Q + t) % (t - 1)
    eng = Engine(local)
    eng.set_taxonomy(parent)
    stream = torch.cuda.Stream(device=dev)
    eng.use_stream(stream.cuda_stream)
    leaf = [NTAX - f for f in range(NFILES)]
    with torch.cuda.stream(stream):
        local_files = {f: eng.synth_member_file(0, U, U, S, T, f).clone() for f in range(NFILES) if owner_of_file(f, world) == rank}
        torch.cuda.synchronize()
        n_in = torch.tensor([sum(int(t.shape[0]) for t in local_files.values())], dtype=torch.int64, device=dev)
        if world > 1:
            dist.all_reduce(n_in)
        ex = KeyRangeExchange(eng, rank, world)
        spl = equal_width_splitters(world, 62)

        def step():
            files = ex.exchange(local_files, NFILES, spl) if world > 1 else [local_files[f] for f in range(NFILES)]
            sets = [KmerSet(k, None, global_taxid=leaf[f]) for f, k in enumerate(files)]
            return eng.common(sets, args.threshold, has_taxid=True)

        res = step()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(args.steps):
            res = step()
        e1.record(stream)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1) / args.steps], dtype=torch.float64, device=dev)
        n_out = torch.tensor([res[0].shape[0]], dtype=torch.int64, device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
            dist.all_reduce(n_out)
    if rank == 0:
        w = 300_000
        W = (1 << 62) // U
        otax = oracle.Taxonomy(parent)
        ofiles = [(k, np.full(len(k), leaf[f], dtype=np.uint32)) for f in range(NFILES) for k in [oracle.member_file(0, w, U, S, T, f)]]
        ek, et = oracle.common(ofiles, args.threshold, has_taxid=True, tax=otax)
        gk = res[0][: len(ek) + 8].cpu().numpy().view(np.uint64)
        gt = res[1][: len(ek) + 8].cpu().numpy().view(np.uint32)
        m = gk < w * W
        ok = bool(np.array_equal(gk[m], ek) and np.array_equal(gt[m], et))
        print(json.dumps({"config": "C5 common -n %d, 64 files, universe %.0e, global taxids, LCA over a 1e4-node tree" % (args.threshold, U),
                          "n_gpus": world, "kmers_in": int(n_in.item()), "kmers_out": int(n_out.item()), "ms_per_step": float(ms.item()),
                          "kmers_per_s": int(n_in.item()) / float(ms.item()) * 1e3, "window_exact_vs_oracle": ok}))
        if not ok:
            raise SystemExit("C5 self-check failed")
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
