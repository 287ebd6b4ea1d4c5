"""Multi-GPU layouts of the scoring path (one process per GPU, ``torch.distributed``).

Two ways the path shards:

* **read sharding** — every rank holds the whole index in its HBM and scores a contiguous slice of the
  records; hit counts are independent per record, so there is no collective on the data path.  Totals (one
  value per document) are summed with one tiny all-reduce when a caller wants file-level scores.
* **document-column sharding** — for an index that does not fit one GPU, rank g holds the byte columns of
  its document range of every row, all ranks score all records against their columns, and the per-record
  score tiles are combined with an all-gather along the document axis (NCCL over NVLink on GPUs; the same
  code runs on ``gloo`` for the CPU tests).  Tiles are double-buffered: the all-gather of tile t runs on a
  side stream while tile t+1 is being queried.
"""

from __future__ import annotations

from typing import Callable, Iterator

import numpy as np
import torch
import torch.distributed as dist


def read_shard(n_records: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous slice ``[lo, hi)`` of the records scored by ``rank`` (i * N / G boundaries)."""
    return n_records * rank // world, n_records * (rank + 1) // world


def column_shards(n_docs: int, world: int, align: int = 128) -> list[tuple[int, int]]:
    """Document ranges per rank, boundaries on multiples of ``align`` documents (128 documents = one 16-byte
    column chunk of a row; xs_cobs_open needs multiples of 8).  Trailing ranks may be empty for tiny indices."""
    units = -(-n_docs // align)
    out = []
    for g in range(world):
        lo = min(n_docs, (units * g // world) * align)
        hi = min(n_docs, (units * (g + 1) // world) * align)
        out.append((lo, hi))
    return out


def choose_column_groups(index_bytes: int, hbm_bytes: int, world: int, budget: float = 0.5) -> int:
    """Fewest document-column groups ``C`` (a divisor of ``world``) whose column shard takes at most ``budget`` of one
    GPU's HBM; the other ``world / C`` factor shards the records.  Column sharding costs every rank the hashing of
    every record and — for rows of a few 128-byte DRAM lines — whole lines per probe whatever the shard width, so
    columns are split only as far as memory demands (180 GB per B200) and the rest of the machine shards records."""
    for c in range(1, world + 1):
        if world % c == 0 and index_bytes / c <= budget * hbm_bytes:
            return c
    return world


def grid_layout(rank: int, world: int, col_groups: int) -> tuple[int, int, list[int]]:
    """``(read_group, column_rank, ranks of this rank's column group)`` of a ``col_groups x world / col_groups`` grid:
    the ``col_groups`` ranks that together hold all document columns are neighbours, ``rank = read_group * C + column_rank``."""
    if col_groups < 1 or world % col_groups:
        raise ValueError("col_groups must divide the world size")
    rg, cr = divmod(rank, col_groups)
    return rg, cr, list(range(rg * col_groups, (rg + 1) * col_groups))


def allreduce_totals(local_totals: np.ndarray | torch.Tensor, group=None) -> torch.Tensor:
    """Sum of per-document totals over read-sharded ranks (int64)."""
    t = torch.as_tensor(local_totals).to(torch.int64).clone()
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        if dist.get_backend(group) == "nccl" and not t.is_cuda:
            t = t.cuda()
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return t


def allgather_columns(local: torch.Tensor, shards: list[tuple[int, int]], group=None) -> torch.Tensor:
    """``local`` = [n, width of this rank's shard] -> [n, n_docs]: all-gather along the document axis.
    Shards may differ in width: tiles are padded to the widest shard for the collective, then trimmed."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world == 1:
        return local
    if local.dtype != torch.uint8:   # collectives on the byte view: every count type travels the same way
        item = local.element_size()
        full = allgather_columns(local.contiguous().view(torch.uint8), [(lo * item, hi * item) for lo, hi in shards], group)
        return full.contiguous().view(local.dtype)
    widths = [hi - lo for lo, hi in shards]
    wmax = max(widths)
    n = local.shape[0]
    send = local
    if local.shape[1] != wmax:
        send = torch.zeros((n, wmax), dtype=local.dtype, device=local.device)
        send[:, : local.shape[1]] = local
    recv = torch.empty((world * n, wmax), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(recv, send.contiguous(), group=group)
    recv = recv.view(world, n, wmax)
    return torch.cat([recv[g, :, : widths[g]] for g in range(world)], dim=1)


def merge_local_best(best: torch.Tensor, count: torch.Tensor, n_best: torch.Tensor, shards: list[tuple[int, int]],
                     rank: int, group=None) -> tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """Combine per-rank (best local document, its count, tie multiplicity) of column-sharded scoring into the
    global best document index, its count and the tie flag — an all-gather of 3 integers per record instead of
    the whole score row.  Ranks hold ascending document ranges, so the first rank reaching the maximum owns the
    first best document."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    local = torch.stack([best.to(torch.int64) + shards[rank][0], count.to(torch.int64), n_best.to(torch.int64)], dim=1).contiguous()
    if world == 1:
        return local[:, 0], local[:, 1], local[:, 2] > 1
    n = local.shape[0]
    recv = torch.empty((world * n, 3), dtype=torch.int64, device=local.device)
    dist.all_gather_into_tensor(recv, local, group=group)
    recv = recv.view(world, n, 3)
    counts = recv[:, :, 1]
    mx, first_rank = counts.max(dim=0)                  # torch returns the first maximal rank
    at_max = counts == mx.unsqueeze(0)
    multiplicity = (recv[:, :, 2] * at_max).sum(dim=0)
    gbest = recv[:, :, 0].gather(0, first_rank.unsqueeze(0)).squeeze(0)
    return gbest, mx, multiplicity > 1


class ColumnShardedIndex:
    """This rank's document-column shard of a COBS classic index plus the score all-gather."""

    def __init__(self, path, rank: int | None = None, world: int | None = None, device: int | None = None, group=None):
        from . import engine

        self.group = group
        self.rank = dist.get_rank(group) if rank is None else rank
        self.world = dist.get_world_size(group) if world is None else world
        self.n_docs = _doc_count(path)
        self.shards = column_shards(self.n_docs, self.world)
        lo, hi = self.shards[self.rank]
        if hi <= lo:
            raise ValueError(f"rank {self.rank} would hold no documents: use at most {-(-self.n_docs // 128)} ranks")
        self.index = engine.CobsIndex(path, device=self.rank if device is None else device, doc_begin=lo, doc_end=hi)

    def query_tiles(self, tiles: Iterator[tuple[int, int, int, int, int]], step: int, dtype: int,
                    consume: Callable[[int, torch.Tensor], None]) -> None:
        """For every tile ``(d_bases, n_bases, d_begin, d_end, n_seq)`` of device pointers: query the local
        columns, all-gather the score tile, hand ``[n_seq, n_docs]`` to ``consume(tile_index, scores)``.
        The collective of tile t overlaps the query of tile t + 1."""
        from ._abi import XS_U8, XS_U16

        tdt = {XS_U8: torch.uint8, XS_U16: torch.uint16}.get(dtype, torch.uint32)
        dev = torch.device("cuda", self.index.device)
        comm = torch.cuda.Stream(device=dev)
        compute = torch.cuda.current_stream(dev)
        pending = None
        for t, (d_bases, n_bases, d_begin, d_end, n_seq) in enumerate(tiles):
            local = torch.empty((n_seq, self.index.n_docs), dtype=tdt, device=dev)
            self.index.query_device(d_bases, n_bases, d_begin, d_end, n_seq, step, dtype, local.data_ptr(), compute.cuda_stream)
            done = torch.cuda.Event()
            done.record(compute)
            if pending is not None:
                pt, full, ev = pending
                ev.synchronize()
                consume(pt, full)
            with torch.cuda.stream(comm):
                comm.wait_event(done)
                full = allgather_columns(local, self.shards, self.group)
                ev = torch.cuda.Event()
                ev.record(comm)
            local.record_stream(comm)
            pending = (t, full, ev)
        if pending is not None:
            pt, full, ev = pending
            ev.synchronize()
            consume(pt, full)


    def classify_tiles(self, tiles: Iterator[tuple[int, int, int, int, int]], step: int, dtype: int,
                       consume: Callable[[int, torch.Tensor, torch.Tensor, torch.Tensor], None]) -> torch.Tensor:
        """Like ``query_tiles`` but the exchange is reduced first: every rank takes the per-record maximum over its
        own columns on the device (xs_scores_reduce_device), and only (best, count, multiplicity) travel.
        ``consume(tile, best_doc_index, best_count, is_tie)``; returns this rank's per-document totals (uint64)."""
        from . import engine
        from ._abi import XS_U8, XS_U16

        tdt = {XS_U8: torch.uint8, XS_U16: torch.uint16}.get(dtype, torch.uint32)
        dev = torch.device("cuda", self.index.device)
        stream = torch.cuda.current_stream(dev)
        totals = torch.zeros(self.index.n_docs, dtype=torch.int64, device=dev)
        for t, (d_bases, n_bases, d_begin, d_end, n_seq) in enumerate(tiles):
            local = torch.empty((n_seq, self.index.n_docs), dtype=tdt, device=dev)
            best = torch.empty(n_seq, dtype=torch.int32, device=dev)
            cnt = torch.empty(n_seq, dtype=torch.int32, device=dev)
            nb = torch.empty(n_seq, dtype=torch.int32, device=dev)
            self.index.query_device(d_bases, n_bases, d_begin, d_end, n_seq, step, dtype, local.data_ptr(), stream.cuda_stream)
            engine.scores_reduce_device(local.data_ptr(), n_seq, self.index.n_docs, dtype, self.index.device, best.data_ptr(),
                                        cnt.data_ptr(), nb.data_ptr(), totals.data_ptr(), stream.cuda_stream)
            consume(t, *merge_local_best(best, cnt, nb, self.shards, self.rank, self.group))
        return totals


def make_comm(rank: int, world: int, device: int, group=None):
    """This rank's ``engine.Comm`` (the library's own NCCL communicator, ``xs_comm_init``): rank 0 creates the id and
    ``torch.distributed`` — already up for the launch plumbing — carries the 128 bytes to the others."""
    from . import engine

    box = [engine.Comm.unique_id() if rank == 0 else None]
    dist.broadcast_object_list(box, src=0, group=group)
    return engine.Comm(box[0], rank, world, device)


def make_grid_comm(rank: int, world: int, device: int, col_groups: int):
    """The library communicator of this rank's column group (``grid_layout``): the leader of every group creates an id,
    one ``all_gather_object`` over the launch group hands every rank its leader's."""
    from . import engine

    rg, cr, members = grid_layout(rank, world, col_groups)
    ids = [None] * world
    dist.all_gather_object(ids, engine.Comm.unique_id() if cr == 0 else None)
    return engine.Comm(ids[members[0]], cr, col_groups, device)


class ShardedScorer:
    """Document-column sharded scoring with the exchange behind the C ABI (BASELINE config 5):
    every rank scores every record tile against its columns into ``[n, w]`` rows of one common padded width
    (``xs_cobs_query_device_ld``), ``xs_allgather_scores`` (ncclAllGather over NVLink) delivers ``[world][n][w]`` on a
    side stream while the next tile is being scored, and ``xs_sharded_reduce_device`` consumes that buffer in place:
    per-record first best document, its count and the tie multiplicity (the reference benchmark's per-read call,
    scripts/benchmark/main.nf:417-436).  No pad, concat or eager pass over the tile."""

    def __init__(self, index, shards: list[tuple[int, int]], comm, dtype: int, max_tile: int):
        from ._abi import XS_U8, XS_U16

        self.index, self.shards, self.comm, self.dtype = index, shards, comm, dtype
        self.world = len(shards)
        self.widths = [hi - lo for lo, hi in shards]
        item = int(dtype)
        unit = 16 // item
        self.w = -(-max(self.widths) // unit) * unit                 # rows stay 16-byte multiples for the vector loads
        self.tdt = {XS_U8: torch.uint8, XS_U16: torch.uint16}.get(dtype, torch.uint32)
        self.dev = torch.device("cuda", index.device)
        self.max_tile = max_tile
        # three tile slots: the scoring kernel is persistent and fills every SM, so the collective of tile t only gets its
        # CTAs placed when the kernel of tile t + 1 drains; with two slots the kernel of tile t + 2 would wait for exactly
        # that collective and the exchange would never overlap anything (measured: +1.1 ms per tile)
        self.n_slots = 3
        self.local = [torch.empty((max_tile, self.w), dtype=self.tdt, device=self.dev) for _ in range(self.n_slots)]
        self.all = [torch.empty((self.world, max_tile, self.w), dtype=self.tdt, device=self.dev) for _ in range(self.n_slots)]
        self.comm_stream = torch.cuda.Stream(device=self.dev, priority=-1)
        self.exchange_ms = 0.0
        self.stall_ms = 0.0        # time the scoring stream waited for a tile slot (time_exchange=True)
        self.query_ms = 0.0        # span of the scoring calls on the scoring stream (all their kernels)

    def run(self, tiles: Iterator[tuple[int, int, int, int, int]], step: int,
            consume: Callable[[int, torch.Tensor, torch.Tensor, torch.Tensor], None], time_exchange: bool = False) -> None:
        """``tiles`` = ``(d_bases, n_bases, d_begin, d_end, n_seq)`` device pointers; ``consume(tile, best, count,
        n_best)`` is called in tile order once a tile's epilogue has finished."""
        from . import engine

        compute = torch.cuda.current_stream(self.dev)
        free = [None] * self.n_slots  # event: the exchange that read local[slot] has finished
        pending = []
        timers = []
        stalls = []
        for t, (d_bases, n_bases, d_begin, d_end, n_seq) in enumerate(tiles):
            if n_seq > self.max_tile:
                raise ValueError("tile larger than max_tile")
            slot = t % self.n_slots
            if time_exchange:
                w0, w1, q1 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
                w0.record(compute)
            if free[slot] is not None:
                compute.wait_event(free[slot])
            if time_exchange:
                w1.record(compute)
            loc, al = self.local[slot], self.all[slot]
            self.index.query_device(d_bases, n_bases, d_begin, d_end, n_seq, step, self.dtype, loc.data_ptr(),
                                    compute.cuda_stream, ld=self.w)
            scored = torch.cuda.Event(enable_timing=time_exchange)
            scored.record(compute)
            if time_exchange:
                stalls.append((w0, w1, scored))
            best = torch.empty(n_seq, dtype=torch.int32, device=self.dev)
            cnt = torch.empty(n_seq, dtype=torch.int32, device=self.dev)
            nb = torch.empty(n_seq, dtype=torch.int32, device=self.dev)
            cs = self.comm_stream
            cs.wait_event(scored)
            if time_exchange:
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(cs)
            # blocks of [world][n_seq][w] inside the slot's buffer (contiguous for this n_seq)
            self.comm.allgather_scores(loc.data_ptr(), n_seq, self.w * int(self.dtype), al.data_ptr(), cs.cuda_stream)
            if time_exchange:
                e1.record(cs)
                timers.append((e0, e1))
            engine.sharded_reduce_device(al.data_ptr(), n_seq, self.dtype, self.index.device, self.w, self.widths,
                                         best.data_ptr(), cnt.data_ptr(), nb.data_ptr(), 0, cs.cuda_stream)
            done = torch.cuda.Event()
            done.record(cs)
            free[slot] = done
            pending.append((t, best, cnt, nb, done))
            while len(pending) > self.n_slots - 1:
                pt, b_, c_, n_, ev = pending.pop(0)
                ev.synchronize()
                consume(pt, b_, c_, n_)
        for pt, b_, c_, n_, ev in pending:
            ev.synchronize()
            consume(pt, b_, c_, n_)
        if time_exchange:
            self.exchange_ms += sum(a.elapsed_time(b) for a, b in timers)
            self.stall_ms += sum(a.elapsed_time(b) for a, b, _ in stalls)
            self.query_ms += sum(b.elapsed_time(c) for _, b, c in stalls)


def _doc_count(path) -> int:
    """num_documents from a COBS classic header (A.1) without loading the index."""
    import struct

    with open(path, "rb") as f:
        head = f.read(18 + 4 + 4 + 1 + 4)
    if head[:18] != b"COBS:CLASSIC_INDEX":
        raise ValueError("document-column sharding needs a COBS classic index")
    return struct.unpack_from("<I", head, 27)[0]
