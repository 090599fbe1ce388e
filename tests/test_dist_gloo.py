"""world_size-2 gloo test of the multi-GPU plumbing (unikmer_b200/dist.py): key-range plan,
one grouped all-to-all-v, per-rank operation, rank-order concatenation == single-process result.
The local operation is stood in by the CPU oracle here (test only): the exchange logic is what
is under test; on GPUs the backend is unikmer_b200.Engine (bench.py --gpus N)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


class CpuBackend:
    """partition_sorted with the ABI's semantics (offsets[0]=0, lower_bound per splitter, n)."""

    def partition_sorted(self, keys, splitters):
        k = keys.numpy().view(np.uint64)
        return np.concatenate([[0], np.searchsorted(k, splitters, side="left"), [len(k)]]).astype(np.uint64)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_files, N, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import oracle
    from unikmer_b200.dist import KeyRangeExchange, equal_width_splitters, gather_rank_order, owner_of_file
    files = {f: torch.from_numpy(oracle.member_file(0, N, N, 3, 4, f).view(np.int64))
             for f in range(n_files) if owner_of_file(f, world) == rank}
    ex = KeyRangeExchange(CpuBackend(), rank, world)
    splitters = equal_width_splitters(world, 62)
    slices = ex.exchange(files, n_files, splitters)
    assert len(slices) == n_files
    lo = 0 if rank == 0 else int(splitters[rank - 1])
    hi = (1 << 62) if rank == world - 1 else int(splitters[rank])
    np_slices = [s.numpy().view(np.uint64) for s in slices]
    for s in np_slices:  # every slice lies in this rank's key range and is still sorted
        assert len(s) == 0 or (int(s[0]) >= lo and int(s[-1]) < hi)
        assert (np.diff(s.astype(np.int64)) > 0).all()
    out = {}
    for name in ("inter", "diff", "union"):
        piece = getattr(oracle, name)(np_slices)[0]
        full = gather_rank_order(torch.from_numpy(piece.view(np.int64)), rank, world)
        out[name] = full.numpy().view(np.uint64).copy()
    if rank == 0:
        q.put(out)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,n_files", [(2, 8), (2, 3)])
def test_key_range_exchange_matches_single_process(world, n_files):
    import oracle
    N = 200_000
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_files, N, q)) for r in range(world)]
    for p in procs:
        p.start()
    out = q.get(timeout=180)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    files = [oracle.member_file(0, N, N, 3, 4, f) for f in range(n_files)]
    for name in ("inter", "diff", "union"):
        assert np.array_equal(out[name], getattr(oracle, name)(files)[0]), name


def test_splitters_and_ownership():
    from unikmer_b200.dist import equal_width_splitters, owner_of_file
    s = equal_width_splitters(8, 62)
    assert len(s) == 7 and int(s[0]) == (1 << 62) // 8 and int(s[-1]) == 7 * ((1 << 62) // 8)
    assert len(equal_width_splitters(1)) == 0
    assert [owner_of_file(f, 4) for f in range(8)] == [0, 1, 2, 3, 0, 1, 2, 3]


def test_pipelined_exchange_plan_pieces_concatenate_to_the_whole_result():
    """The pipelined exchange (PeerPullExchange.plan_chunks / exchange_chunks) cuts every rank's key range into K pieces and
    runs the operations piece by piece with UKM_F_SHARD semantics; the concatenation over (rank, piece) must equal the
    whole-file result.  Checked here on the CPU with the plan's splitters and the oracle as the per-piece operation
    (inter: an empty slice empties the piece's result -- the shard rule -- instead of the whole-file quirk B-3)."""
    import oracle
    from unikmer_b200.dist import equal_width_splitters, fine_splitters
    N, n_files = 200_000, 8
    files = [oracle.member_file(0, N, N, 3, 4, f) for f in range(n_files)]
    files[5] = files[5][files[5] < (1 << 59)]  # a file that has nothing in most pieces
    for world, K in ((2, 4), (8, 3), (4, 1)):
        fine = fine_splitters(equal_width_splitters(world, 62), world, K)
        assert len(fine) == world * K - 1 and (np.diff(fine.astype(np.float64)) > 0).all()
        cuts = [np.concatenate([[0], np.searchsorted(f, fine, side="left"), [len(f)]]) for f in files]
        res = {"inter": [], "diff": [], "union": []}
        for p in range(world * K):
            sl = [f[c[p]:c[p + 1]] for f, c in zip(files, cuts)]
            if all(len(x) for x in sl):
                res["inter"].append(oracle.inter(sl)[0])
            sub = [x for x in sl[1:] if len(x)]
            res["diff"].append(oracle.diff([sl[0]] + sub)[0] if len(sl[0]) and sub else sl[0])
            res["union"].append(oracle.union(sl)[0])
        exp_i = files[0]
        for f in files[1:]:
            exp_i = exp_i[np.isin(exp_i, f)]
        assert np.array_equal(np.concatenate(res["inter"]) if res["inter"] else np.zeros(0, np.uint64), exp_i)
        assert np.array_equal(np.concatenate(res["diff"]), oracle.diff(files)[0])
        assert np.array_equal(np.concatenate(res["union"]), oracle.union(files)[0])
