"""Key-range sharding of the set operations across the GPUs of one box.

One process per GPU (torch.distributed; NCCL on GPUs, gloo in the CPU tests).  All six
operations are key-local -- the result for key x depends only on the occurrences of x --
so the key space is cut into G contiguous ranges, rank r receives every input file's
slice in its range with ONE grouped all-to-all-v (batched isend/irecv = one
ncclGroupStart/End), runs the single-GPU operation on its bucket, and the global result is
the concatenation of the per-rank outputs in rank order (already globally sorted).

For sorted inputs the partition is G-1 binary searches per file (ukm_partition_sorted) and
every payload is a contiguous slice: no scatter kernel, no second collective.

`backend` is whatever runs the local pieces: unikmer_b200.Engine on a GPU.  (The gloo
tests pass a stand-in so the exchange plan can be checked on CPU.)
"""
from __future__ import annotations

from typing import Dict, List, Sequence

import os

import numpy as np
import torch
import torch.distributed as dist


def equal_width_splitters(world: int, key_bits: int = 62) -> np.ndarray:
    """G-1 splitters cutting [0, 2^key_bits) into equal-width ranges (uniform keys: k-mer codes of
    random sequence, hashes).  Range r = [s[r-1], s[r])."""
    return np.array([((i << key_bits) // world) for i in range(1, world)], dtype=np.uint64)


def fine_splitters(splitters: np.ndarray, world: int, chunks: int, key_hi: int = 1 << 62) -> np.ndarray:
    """Cut every rank's key range [s[r-1], s[r]) into `chunks` equal-width pieces: G*K - 1 splitters, fine range r*K + c is
    piece c of rank r (the rank boundaries themselves are among them)."""
    edges = [0] + [int(x) for x in splitters] + [int(key_hi)]
    fine = []
    for r in range(world):
        lo, hi = edges[r], edges[r + 1]
        fine += [lo + (hi - lo) * c // chunks for c in range(chunks)]
    return np.array(fine[1:], dtype=np.uint64)


def owner_of_file(f: int, world: int) -> int:
    """Initial placement: file f lives on rank f mod G (files arrive from different readers)."""
    return f % world


class KeyRangeExchange:
    def __init__(self, backend, rank: int, world: int, group=None):
        self.backend = backend
        self.rank = rank
        self.world = world
        self.group = group

    def plan(self, local_files: Dict[int, torch.Tensor], n_files: int, splitters: np.ndarray):
        """Binary-search every local file at the splitters and share the slice sizes.
        Returns (offsets per local file, counts[f][r] = elements of file f in range r, for ALL files)."""
        G = self.world
        offsets = {}
        counts = torch.zeros(n_files, G, dtype=torch.int64)
        for f, t in local_files.items():
            off = np.asarray(self.backend.partition_sorted(t, splitters), dtype=np.int64)
            assert len(off) == G + 1 and off[0] == 0 and off[-1] == t.shape[0]
            offsets[f] = off
            counts[f] = torch.from_numpy(np.diff(off))
        if G > 1:
            dev = next(iter(local_files.values())).device if local_files else torch.device("cpu")
            c = counts.to(dev)
            dist.all_reduce(c, op=dist.ReduceOp.SUM, group=self.group)  # every file has exactly one owner
            counts = c.cpu()
        return offsets, counts

    def exchange(self, local_files: Dict[int, torch.Tensor], n_files: int, splitters: np.ndarray,
                 local_values: Dict[int, torch.Tensor] = None):
        """One all-to-all-v.  Returns, for every file id in order, this rank's key-range slice of the keys
        (and, if `local_values` holds a per-key array for every local file -- taxids --, a second list with the
        matching value slices, moved in the same grouped call)."""
        G, me = self.world, self.rank
        offsets, counts = self.plan(local_files, n_files, splitters)
        with_vals = local_values is not None
        out: List[torch.Tensor] = [None] * n_files  # type: ignore
        outv: List[torch.Tensor] = [None] * n_files  # type: ignore
        ref = next(iter(local_files.values())) if local_files else None
        refv = next(iter(local_values.values())) if (with_vals and local_values) else None
        ops = []
        for f in range(n_files):
            o = owner_of_file(f, G)
            if o == me:
                t, off = local_files[f], offsets[f]
                out[f] = t[off[me]:off[me + 1]]  # own slice: a view, no copy
                if with_vals:
                    outv[f] = local_values[f][off[me]:off[me + 1]]
                for r in range(G):
                    if r != me and off[r + 1] > off[r]:
                        ops.append(dist.P2POp(dist.isend, t[off[r]:off[r + 1]], r, group=self.group))
                        if with_vals:
                            ops.append(dist.P2POp(dist.isend, local_values[f][off[r]:off[r + 1]], r, group=self.group))
            else:
                n = int(counts[f, me])
                buf = torch.empty(n, dtype=ref.dtype if ref is not None else torch.int64,
                                  device=ref.device if ref is not None else "cpu")
                out[f] = buf
                if with_vals:
                    outv[f] = torch.empty(n, dtype=refv.dtype if refv is not None else torch.int32,
                                          device=buf.device)
                if n:
                    ops.append(dist.P2POp(dist.irecv, buf, o, group=self.group))
                    if with_vals:
                        ops.append(dist.P2POp(dist.irecv, outv[f], o, group=self.group))
        if ops:
            for req in dist.batch_isend_irecv(ops):
                req.wait()
        return (out, outv) if with_vals else out

    @staticmethod
    def exchanged_bytes(counts: torch.Tensor, world: int) -> int:
        """Bytes that cross NVLink in one exchange (everything except the owners' own slices)."""
        total = int(counts.sum()) * 8
        own = sum(int(counts[f, owner_of_file(f, world)]) for f in range(counts.shape[0])) * 8
        return total - own


def gather_rank_order(piece: torch.Tensor, rank: int, world: int, group=None) -> torch.Tensor:
    """Concatenate per-rank outputs in rank order on every rank (tests / small results only)."""
    if world == 1:
        return piece
    n = torch.tensor([piece.shape[0]], dtype=torch.int64, device=piece.device)
    sizes = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(sizes, n, group=group)
    bufs = [torch.empty(int(s.item()), dtype=piece.dtype, device=piece.device) for s in sizes]
    dist.all_gather(bufs, piece.contiguous(), group=group) if len(set(int(s.item()) for s in sizes)) == 1 else _uneven_all_gather(bufs, piece, rank, world, group)
    return torch.cat(bufs)


def _uneven_all_gather(bufs: Sequence[torch.Tensor], piece: torch.Tensor, rank: int, world: int, group=None):
    for r in range(world):
        if r == rank:
            bufs[r].copy_(piece)
        dist.broadcast(bufs[r], src=r, group=group)


class PeerPullExchange:
    """Key-range exchange over NVLink peer mappings, driven by the copy engines.

    Set-up (once, like creating a communicator): every rank shares CUDA-IPC handles of its resident files, every
    other rank maps them.  Per step each rank PULLS its key-range slice of every remote file with an asynchronous
    peer copy on side streams -- no NCCL kernels, no SMs, so the transfers overlap the merge passes that already
    have their inputs -- and hands back (tensor, event) pairs; consumers wait on the event of the file they need.
    """

    def __init__(self, backend, rank: int, world: int, local_files: Dict[int, torch.Tensor], n_files: int, group=None,
                 n_streams: int = 0, split: int = 0):
        from torch.multiprocessing.reductions import reduce_tensor
        # copy-engine parallelism: one peer copy keeps one engine busy, several in flight on different streams add up
        # until the link is full.  UKM_PULL_STREAMS / UKM_PULL_SPLIT override (streams; pieces one file's slice is cut into)
        n_streams = int(n_streams or os.environ.get("UKM_PULL_STREAMS", "8"))
        self.split = max(1, int(split or os.environ.get("UKM_PULL_SPLIT", "1")))
        # a peer copy is queued on a stream of the OWNER's device (torch's rule for cross-device copies): with one stream per
        # owner its files travel one after the other; two per owner measured 12.7 instead of 13.9 ms per C3 step at N = 4,
        # where the 6 GB a rank pulls per step bound the step (profiles/r02_exp_pull_n4.md)
        self.per_src = max(1, int(os.environ.get("UKM_PULL_SRC_STREAMS", "2")))
        self.src_streams: Dict[tuple, torch.cuda.Stream] = {}
        self.src_next: Dict[int, int] = {}
        self.backend, self.rank, self.world, self.group, self.n_files = backend, rank, world, group, n_files
        self.local = local_files
        self.device = next(iter(local_files.values())).device
        shared = {f: reduce_tensor(t) for f, t in local_files.items()}
        everyone = [None] * world
        dist.all_gather_object(everyone, shared, group=group)
        self.peer: Dict[int, torch.Tensor] = {}
        for r, d in enumerate(everyone):
            if r == rank:
                continue
            for f, (fn, args) in d.items():
                self.peer[f] = fn(*args)  # a tensor on the owner's device, backed by the owner's memory
        self.streams = [torch.cuda.Stream(device=self.device) for _ in range(n_streams)]
        self.bufs: Dict[int, torch.Tensor] = {}

    def exchange_async(self, splitters: np.ndarray):
        """Returns (slices, events): slices[f] = this rank's key-range slice of file f; events[f] = CUDA event to wait
        on before reading it (None for local views)."""
        G, me, n_files = self.world, self.rank, self.n_files
        # slice boundaries: the owner binary-searches its files; one small all-reduce shares them
        bounds = torch.zeros(n_files, G + 1, dtype=torch.int64)
        for f, t in self.local.items():
            bounds[f] = torch.from_numpy(np.asarray(self.backend.partition_sorted(t, splitters), dtype=np.int64))
        b = bounds.to(self.device)
        dist.all_reduce(b, op=dist.ReduceOp.SUM, group=self.group)
        bounds = b.cpu()
        cur = torch.cuda.current_stream(self.device)
        ready = torch.cuda.Event()
        ready.record(cur)
        slices: List[torch.Tensor] = [None] * n_files  # type: ignore
        events: List[torch.cuda.Event] = [None] * n_files  # type: ignore
        for i, f in enumerate(range(n_files)):
            lo, hi = int(bounds[f, me]), int(bounds[f, me + 1])
            if f in self.local:
                slices[f] = self.local[f][lo:hi]
                continue
            n = hi - lo
            buf = self.bufs.get(f)
            if buf is None or buf.shape[0] < n:
                buf = torch.empty(int(n * 1.05) + 16, dtype=self.peer[f].dtype, device=self.device)
                self.bufs[f] = buf
            s = self.streams[i % len(self.streams)]
            s.wait_event(ready)  # do not overwrite a buffer the previous step may still read
            with torch.cuda.stream(s):
                buf[:n].copy_(self.peer[f][lo:hi], non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(s)
            slices[f] = buf[:n]
            events[f] = ev
        self.bytes_pulled = sum(int(slices[f].shape[0]) * 8 for f in range(n_files) if f not in self.local)
        return slices, events

    def wait(self, events, files):
        cur = torch.cuda.current_stream(self.device)
        for f in files:
            if events[f] is None:
                continue
            for e in (events[f] if isinstance(events[f], (list, tuple)) else (events[f],)):
                cur.wait_event(e)

    # ---- pipelined form: the rank's key range in `chunks` pieces, piece c + 1 pulled while piece c is computed ----
    def plan_chunks(self, splitters: np.ndarray, chunks: int, key_hi: int = 1 << 62):
        """Cut every rank's key range [s[r-1], s[r]) into `chunks` equal-width pieces and find, ONCE, where every file
        is cut (the owners binary-search their files, one all-reduce shares the positions).  The plan stays valid as
        long as the resident files do not change; exchange_chunks() reuses it every step -- no per-step collective,
        no host round trip for the bounds."""
        G, K = self.world, int(chunks)
        fine = fine_splitters(splitters, G, K, key_hi)
        bounds = torch.zeros(self.n_files, G * K + 1, dtype=torch.int64)
        for f, t in self.local.items():
            bounds[f] = torch.from_numpy(np.asarray(self.backend.partition_sorted(t, fine), dtype=np.int64))
        b = bounds.to(self.device)
        dist.all_reduce(b, op=dist.ReduceOp.SUM, group=self.group)
        self.chunk_bounds = b.cpu()
        self.chunks = K
        self.chunk_bufs: Dict[tuple, torch.Tensor] = {}
        self.chunk_done = [None, None]  # event after the compute that last read buffer set q
        self.chunk_prefetched = None
        self.prefetch_next_step = True
        self.chunk_seq = 0
        self.time_pulls = False  # measurement: keep (start, end, bytes) events of every peer copy in pull_marks
        self.pull_marks: list = []

    def pull_timing(self):
        """(bytes, ms, GB/s) of the peer copies recorded since time_pulls was set: first start to last end."""
        if not self.pull_marks:
            return None
        torch.cuda.synchronize(self.device)
        t_ref = self.pull_marks[0][0]
        lo = min(t_ref.elapsed_time(t0) for t0, _, _ in self.pull_marks)
        hi = max(t_ref.elapsed_time(e) for _, e, _ in self.pull_marks)
        busy = sum(t0.elapsed_time(e) for t0, e, _ in self.pull_marks)
        nbytes = sum(b for _, _, b in self.pull_marks)
        return {"bytes": nbytes, "copies": len(self.pull_marks), "span_ms": hi - lo, "sum_copy_ms": busy,
                "GBps_per_copy": nbytes / max(busy, 1e-9) / 1e6}

    def exchange_chunks(self):
        """Generator over the pieces of this rank's key range: yields (slices, events) like exchange_async.  The pulls
        of piece c + 1 are queued on the copy engines before piece c is handed out, into the other buffer set; a set is
        overwritten only after the compute that read it (everything the consumer queued before asking for the next
        piece) has finished."""
        G, me, K = self.world, self.rank, self.chunks
        cur = torch.cuda.current_stream(self.device)

        def issue(c):
            q = self.chunk_seq & 1  # buffer sets alternate piece after piece, across steps too
            self.chunk_seq += 1
            slices: List[torch.Tensor] = [None] * self.n_files  # type: ignore
            events: List[torch.cuda.Event] = [None] * self.n_files  # type: ignore
            i = 0
            for f in range(self.n_files):
                lo, hi = int(self.chunk_bounds[f, me * K + c]), int(self.chunk_bounds[f, me * K + c + 1])
                if f in self.local:
                    slices[f] = self.local[f][lo:hi]
                    continue
                n = hi - lo
                buf = self.chunk_bufs.get((f, q))
                if buf is None or buf.shape[0] < n:
                    buf = torch.empty(int(n * 1.05) + 16, dtype=self.peer[f].dtype, device=self.device)
                    self.chunk_bufs[(f, q)] = buf
                evs = []
                for part in range(self.split):  # a slice may be cut into pieces that travel on different copy engines
                    a, b = n * part // self.split, n * (part + 1) // self.split
                    if b <= a and part:
                        continue
                    s = self.streams[i % len(self.streams)]
                    i += 1
                    if self.chunk_done[q] is not None:
                        s.wait_event(self.chunk_done[q])
                    with torch.cuda.stream(s):
                        if self.time_pulls:
                            t0 = torch.cuda.Event(enable_timing=True)
                            t0.record(s)
                        if self.per_src <= 1:  # torch: the copy runs on the OWNER device's current stream of this process
                            buf[a:b].copy_(self.peer[f][lo + a:lo + b], non_blocking=True)
                        else:  # ... on one of per_src streams of the owner device, so copies from one owner overlap
                            od = self.peer[f].device
                            j = self.src_next.get(od.index, 0)
                            self.src_next[od.index] = j + 1
                            key = (od.index, j % self.per_src)
                            if key not in self.src_streams:
                                self.src_streams[key] = torch.cuda.Stream(device=od)
                            with torch.cuda.stream(self.src_streams[key]):
                                buf[a:b].copy_(self.peer[f][lo + a:lo + b], non_blocking=True)
                        e = torch.cuda.Event(enable_timing=self.time_pulls)
                        e.record(s)
                    if self.time_pulls:
                        self.pull_marks.append((t0, e, (b - a) * 8))
                    evs.append(e)
                ev = evs
                slices[f] = buf[:n]
                events[f] = ev
            return slices, events, q

        # piece 0 of this step was queued at the end of the step before (a stream of steps over resident files: the pull of
        # the next step's first piece overlaps this step's last piece, so only the very first step waits for a transfer)
        nxt = self.chunk_prefetched if getattr(self, "chunk_prefetched", None) is not None else issue(0)
        self.chunk_prefetched = None
        for c in range(K):
            now = nxt
            # queue the next pulls BEFORE handing this piece out: the operations the consumer runs on it block the host
            if c + 1 < K:
                nxt = issue(c + 1)
            elif self.prefetch_next_step:
                self.chunk_prefetched = issue(0)
            yield now[0], now[1]
            done = torch.cuda.Event()
            done.record(cur)
            self.chunk_done[now[2]] = done
