"""Two-stage classification without the file round trip: genus Bloom filter -> species index (+ SVM).

The reference's ``xspect all`` (main.py:84-188) runs ``filter_genus`` — which writes the records whose genus score
reaches the threshold to ``filtered_sequences/genus_filtered_<uuid>.fasta`` (filter_sequences.py:71-124,
file_io.py:166-191) — and then ``classify_species`` on that directory, so every stage re-parses its input.  Here
the reads are uploaded once: the Bloom kernel scores all records, the threshold mask is taken on the device,
and the species index is queried for the surviving records through their offsets into the same device buffer.
The file-based pipeline (``main.py all``) stays available and produces the same numbers; this module is the
GPU-resident equivalent for large read sets (BASELINE.json config 3).

Threshold semantics are the reference's: a record is kept when ``round(hits / num_kmers, 2) >= threshold``
(models/result.py:59,92-123), evaluated exactly through a per-``num_kmers`` table of minimal hit counts.
"""

from __future__ import annotations

import numpy as np
import torch

from . import engine
from .seqio import SequenceBatch


def min_hits_table(max_kmers: int, threshold: float) -> np.ndarray:
    """``t[n]`` = smallest hit count h with ``round(h / n, 2) >= threshold`` (n + 1 when none), n = 0..max_kmers.
    ``round(h / n, 2)`` is non-decreasing in h, so a binary search per n with Python's own ``round`` is exact."""
    t = np.zeros(max_kmers + 1, dtype=np.int64)
    t[0] = 1
    for n in range(1, max_kmers + 1):
        lo, hi = 0, n + 1
        while lo < hi:
            mid = (lo + hi) // 2
            if mid <= n and round(mid / n, 2) >= threshold:
                hi = mid
            else:
                lo = mid + 1
        t[n] = lo
    return t


_HOST_POOL = None


def _host_pool():
    """Host threads for the per-block NumPy passes over tens of millions of records (NumPy releases the GIL)."""
    global _HOST_POOL
    if _HOST_POOL is None:
        import os
        from concurrent.futures import ThreadPoolExecutor
        _HOST_POOL = ThreadPoolExecutor(max_workers=max(1, min(16, (os.cpu_count() or 2))))
    return _HOST_POOL


def _plan_blocks(begin: np.ndarray, end: np.ndarray, block_records: int, block_bytes: int) -> list[tuple[int, int, int, int, int, int]]:
    """Record ranges ``(i0, i1, byte_lo, byte_hi, min_len, max_len)`` of at most ``block_records`` records whose bases
    span at most about ``block_bytes`` (records in file order: the span of a range is what its records cover).  A
    batch whose records are not laid out in order (spans much larger than the records) is one block."""
    n = int(begin.size)
    if n == 0:
        return []
    sample = slice(0, n, max(1, n // 4096))
    avg = max(1.0, float((end[sample] - begin[sample]).mean()))
    per = max(1, min(block_records, int(block_bytes / avg)))

    def stats(i0: int):
        i1 = min(n, i0 + per)
        lens = end[i0:i1] - begin[i0:i1]
        return i0, i1, int(begin[i0:i1].min()), int(end[i0:i1].max()), int(lens.min()), int(lens.max())

    starts = range(0, n, per)
    out = list(_host_pool().map(stats, starts)) if len(starts) > 1 else [stats(0)]
    if len(out) > 1 and any(hi - lo > 4 * block_bytes + (1 << 20) for _, _, lo, hi, _, _ in out):
        return [(0, n, min(o[2] for o in out), max(o[3] for o in out), min(o[4] for o in out), max(o[5] for o in out))]
    return out


def genus_then_species(genus_model, species_model, sequence_input, threshold: float = 0.7, step: int = 1,
                       predict: bool = True, block_records: int | None = None, block_bytes: int = 320 << 20) -> dict:
    """Score every record against the genus filter, keep those reaching ``threshold`` and classify them with the
    species model; returns per-record genus hits, the kept mask, read-level species calls for the kept records,
    file-level species totals / scores and (for an SVM species model) the prediction.

    The records go through the device in blocks on three streams: while block b is scored (Bloom kernel -> threshold
    from the exact table -> kept offsets compacted to the front of the block, the other slots empty records -> species
    kernel -> argmax / totals), block b + 1 is copied in and the results of block b - 1 are copied out.  No host
    synchronisation between blocks (the number of kept records per block is read back with the results); host arrays
    that are not page-locked go through page-locked staging so that no copy blocks the enqueueing thread, and the
    host's own passes over the records (length checks, ``num_kmers``, result compaction) run block-wise on a thread
    pool, those that need no result while the device works.  ``out["timing"]`` says where the call's time went."""
    import time

    if threshold < 0 or threshold > 1:
        raise ValueError("The filter threshold must be between 0 and 1.")
    stamps = [("start", time.perf_counter())]
    mark = lambda name: stamps.append((name, time.perf_counter()))
    batch = sequence_input if isinstance(sequence_input, SequenceBatch) else genus_model._to_batch(sequence_input)
    bf = genus_model.bf.filter
    ix = species_model.index.index
    if bf.device != ix.device:
        raise ValueError("genus filter and species index must live on the same GPU")
    dev = torch.device("cuda", ix.device)
    n = len(batch)
    k = genus_model.k
    n_docs = ix.n_docs
    if block_records is None:
        # about eight blocks so that the first copy-in and the last copy-out are a small part of the pass, but blocks of at
        # least 500 k records (a 150-bp block then still has the 32 Mi windows the bucketed kernels want) and at most 2 M
        block_records = min(2_000_000, max(500_000, -(-n // 8)))
    blocks = _plan_blocks(batch.begin, batch.end, block_records, block_bytes)
    if blocks and min(bl[4] for bl in blocks) <= k:       # ProbabilisticFilterModel._check_lengths, from the block statistics
        raise ValueError("Invalid sequence, must be longer than k")
    mark("plan_and_check")
    # results land in page-locked arrays at the block's own record offset (a block keeps at most its own records)
    h_hits = engine.pinned_empty(n, np.int32)
    h_keep = engine.pinned_empty(n, np.uint8)
    h_best = engine.pinned_empty(n, np.int32)
    h_cnt = engine.pinned_empty(n, np.int32)
    h_nb = engine.pinned_empty(n, np.int32)
    h_kept_n = engine.pinned_empty(max(1, len(blocks)), np.int64)
    t_hits, t_keep, t_best, t_cnt, t_nb, t_kept_n = (torch.from_numpy(a) for a in (h_hits, h_keep, h_best, h_cnt, h_nb, h_kept_n))
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")                   # read-only NumPy buffers (bytes input): they are only read
        bases_t = torch.from_numpy(batch.bases)
        begin_t = torch.from_numpy(batch.begin.view(np.int64))
        end_t = torch.from_numpy(batch.end.view(np.int64))
    n_slots = min(3, len(blocks))
    max_rec = max((bl[1] - bl[0] for bl in blocks), default=0)
    max_span = max((bl[3] - bl[2] for bl in blocks), default=0)

    def staging(t: torch.Tensor, count: int, dtype):
        """Page-locked staging per slot for a host array the driver would copy synchronously."""
        if count == 0 or t.is_pinned():
            return None
        return [torch.from_numpy(engine.pinned_empty(count, dtype)) for _ in range(n_slots)]

    st_bases, st_begin, st_end = staging(bases_t, max_span, np.uint8), staging(begin_t, max_rec, np.int64), staging(end_t, max_rec, np.int64)
    table_np = np.zeros(0, np.int64)
    mark("host_buffers")
    ev_pairs = {"genus": [], "threshold": [], "species": []}

    def timed_ev(stream):
        e = torch.cuda.Event(enable_timing=True)
        e.record(stream)
        return e

    with torch.cuda.device(dev):
        compute = torch.cuda.current_stream(dev)
        copy_in, copy_out = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)
        slots = [(torch.empty(max_span, dtype=torch.uint8, device=dev), torch.empty(max_rec, dtype=torch.int64, device=dev),
                  torch.empty(max_rec, dtype=torch.int64, device=dev)) for _ in range(n_slots)]
        slot_free: list = [None] * n_slots        # compute has finished with the slot's device buffers
        slot_sent: list = [None] * n_slots        # the slot's staging has been copied to the device
        totals = torch.zeros(n_docs, dtype=torch.int64, device=dev)
        total_kmers_d = torch.zeros((), dtype=torch.int64, device=dev)
        table = None
        start = torch.cuda.Event()
        start.record(compute)
        copy_in.wait_event(start)          # the slots were allocated on the compute stream

        def upload(bi: int):
            i0, i1, lo, hi = blocks[bi][:4]
            sl = bi % n_slots
            d_bases, d_begin, d_end = slots[sl]
            src = [bases_t[lo:hi], begin_t[i0:i1], end_t[i0:i1]]
            if st_bases is not None or st_begin is not None or st_end is not None:
                if slot_sent[sl] is not None:
                    slot_sent[sl].synchronize()            # three blocks back: long done
                for j, (st, cnt) in enumerate(((st_bases, hi - lo), (st_begin, i1 - i0), (st_end, i1 - i0))):
                    if st is not None:
                        st[sl][:cnt].copy_(src[j])
                        src[j] = st[sl][:cnt]
            with torch.cuda.stream(copy_in):
                if slot_free[sl] is not None:
                    copy_in.wait_event(slot_free[sl])
                d_bases[: hi - lo].copy_(src[0], non_blocking=True)
                d_begin[: i1 - i0].copy_(src[1], non_blocking=True)
                d_end[: i1 - i0].copy_(src[2], non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(copy_in)
            slot_sent[sl] = ev
            return ev

        ready = [upload(bi) for bi in range(min(2, len(blocks)))]
        for bi, (i0, i1, lo, hi, _, max_len) in enumerate(blocks):
            if bi + 2 < len(blocks):
                ready.append(upload(bi + 2))
            nb_rec = i1 - i0
            sl = bi % n_slots
            d_bases, d_begin, d_end = slots[sl]
            max_nk = -((max_len - k + 1) // -step)
            compute.wait_event(ready[bi])
            b_rel = d_begin[:nb_rec] - lo
            e_rel = d_end[:nb_rec] - lo
            # ---- stage 1: genus Bloom filter
            d_hits = torch.empty(nb_rec, dtype=torch.int32, device=dev)
            e0 = timed_ev(compute)
            bf.query_device(d_bases.data_ptr(), hi - lo, b_rel.data_ptr(), e_rel.data_ptr(), nb_rec, step, d_hits.data_ptr(),
                            compute.cuda_stream)
            e1 = timed_ev(compute)
            nk = torch.div(e_rel - b_rel - k + step, step, rounding_mode="floor")
            if max_nk <= (1 << 20):
                if max_nk >= table_np.size:
                    table_np = min_hits_table(max_nk, threshold)
                    table = torch.from_numpy(table_np).to(dev)
                d_keep = d_hits.to(torch.int64) >= table[nk]
            else:   # a few very long records: evaluate the rounded score on the host
                hh = d_hits.cpu().numpy()
                nkh = nk.cpu().numpy()
                d_keep = torch.from_numpy(np.fromiter((round(int(a) / int(b), 2) >= threshold for a, b in zip(hh, nkh)),
                                                      dtype=bool, count=nb_rec)).to(dev)
            # ---- stage 2: species index on the kept records: their offsets move to the front of the block (stable), the
            # remaining slots are empty records, so the launch needs no count from the device
            rank_ = torch.cumsum(d_keep, dim=0)
            dst = torch.where(d_keep, rank_ - 1, torch.full_like(rank_, nb_rec))
            s_begin = torch.zeros(nb_rec + 1, dtype=torch.int64, device=dev)
            s_end = torch.zeros(nb_rec + 1, dtype=torch.int64, device=dev)
            s_begin.scatter_(0, dst, b_rel)
            s_end.scatter_(0, dst, e_rel)
            total_kmers_d += (nk * d_keep).sum()
            dt = 1 if max_nk <= 255 else (2 if max_nk <= 65535 else 4)
            d_counts = torch.empty((nb_rec, n_docs), dtype={1: torch.uint8, 2: torch.uint16, 4: torch.uint32}[dt], device=dev)
            best = torch.empty(nb_rec, dtype=torch.int32, device=dev)
            cnt = torch.empty(nb_rec, dtype=torch.int32, device=dev)
            nbest = torch.empty(nb_rec, dtype=torch.int32, device=dev)
            e2 = timed_ev(compute)
            ix.query_device(d_bases.data_ptr(), hi - lo, s_begin.data_ptr(), s_end.data_ptr(), nb_rec, step, dt,
                            d_counts.data_ptr(), compute.cuda_stream)
            engine.scores_reduce_device(d_counts.data_ptr(), nb_rec, n_docs, dt, ix.device, best.data_ptr(), cnt.data_ptr(),
                                        nbest.data_ptr(), totals.data_ptr(), compute.cuda_stream)
            e3 = timed_ev(compute)
            ev_pairs["genus"].append((e0, e1)); ev_pairs["threshold"].append((e1, e2)); ev_pairs["species"].append((e2, e3))
            done = torch.cuda.Event()
            done.record(compute)
            slot_free[sl] = done
            keep_u8 = d_keep.to(torch.uint8)
            kept_n = rank_[-1:].clone()
            scored = torch.cuda.Event()
            scored.record(compute)
            # ---- results of this block to the host on their own stream
            with torch.cuda.stream(copy_out):
                copy_out.wait_event(scored)
                t_hits[i0:i1].copy_(d_hits, non_blocking=True)
                t_keep[i0:i1].copy_(keep_u8, non_blocking=True)
                t_best[i0:i1].copy_(best, non_blocking=True)
                t_cnt[i0:i1].copy_(cnt, non_blocking=True)
                t_nb[i0:i1].copy_(nbest, non_blocking=True)
                t_kept_n[bi : bi + 1].copy_(kept_n, non_blocking=True)
            for t in (d_hits, keep_u8, best, cnt, nbest, kept_n):
                t.record_stream(copy_out)
        mark("enqueue")
        # host work that needs no result overlaps the device
        num_kmers = np.empty(n, dtype=np.int64)

        def fill_num_kmers(bl):
            i0, i1 = bl[0], bl[1]
            lens = (batch.end[i0:i1] - batch.begin[i0:i1]).astype(np.int64)
            num_kmers[i0:i1] = lens - (k - 1) if step == 1 else (lens - k + step) // step

        list(_host_pool().map(fill_num_kmers, blocks))
        mark("num_kmers")
        copy_out.synchronize()
        compute.synchronize()
        mark("device_wait")
        totals_h = totals.cpu().numpy()
        total_kmers = int(total_kmers_d.item())
    keep = h_keep.view(np.bool_)
    kept_n = [int(c) for c in h_kept_n[: len(blocks)]]
    kept_off = np.concatenate([[0], np.cumsum(kept_n)]).astype(np.int64)
    m = int(kept_off[-1])
    best_h, cnt_h, nb_h = np.empty(m, np.int32), np.empty(m, np.int32), np.empty(m, np.int32)
    kept_index = np.empty(m, np.int64)

    def collect(j: int):
        i0, i1 = blocks[j][0], blocks[j][1]
        o0, o1 = int(kept_off[j]), int(kept_off[j + 1])
        best_h[o0:o1] = h_best[i0 : i0 + kept_n[j]]
        cnt_h[o0:o1] = h_cnt[i0 : i0 + kept_n[j]]
        nb_h[o0:o1] = h_nb[i0 : i0 + kept_n[j]]
        kept_index[o0:o1] = np.flatnonzero(keep[i0:i1]) + i0

    list(_host_pool().map(collect, range(len(blocks))))
    out = {
        "batch": batch, "genus_label": next(iter(genus_model.display_names)), "genus_hits": h_hits.view(np.uint32), "num_kmers": num_kmers,
        "kept": keep, "kept_index": kept_index, "labels": ix.names,
        "best": best_h.view(np.uint32), "best_hits": cnt_h.view(np.uint32), "ambiguous": nb_h > 1,
    }
    mark("collect")
    out["total_hits"] = {name: int(v) for name, v in zip(ix.names, totals_h)}
    out["total_scores"] = {name: round(v / total_kmers, 2) for name, v in out["total_hits"].items()} if total_kmers else {}
    out["total_kmers"] = total_kmers
    out["prediction"] = None
    if predict and hasattr(species_model, "_get_svm") and total_kmers:
        x = [list(dict(sorted(out["total_scores"].items())).values())]
        out["prediction"] = str(species_model._get_svm(None).predict(x)[0])
    mark("scores")
    # where the call's time went: host phases (seconds) and the device time of the stages summed over the blocks
    out["timing"] = {"host_s": {b[0]: round(b[1] - a[1], 4) for a, b in zip(stamps, stamps[1:])}, "blocks": len(blocks),
                     "staged_through_pinned": [name for name, st in (("bases", st_bases), ("begin", st_begin), ("end", st_end)) if st is not None],
                     "device_ms": {name: round(sum(a.elapsed_time(b) for a, b in pairs), 2) for name, pairs in ev_pairs.items()}}
    return out


def genus_then_species_sharded(genus_model, species_model, sequence_input, threshold: float = 0.7, step: int = 1, group=None) -> dict:
    """``genus_then_species`` over read-sharded ranks (BASELINE config 3: one process per GPU, both model files
    replicated in every GPU's HBM, each rank holding a contiguous slice of the reads): per-record results stay on the
    rank that scored them; the per-document species totals and the k-mer count of the kept reads are summed with one
    small all-reduce (the only collective — hit counts are additive over reads), and the file-level scores and the SVM
    prediction are computed from the global totals exactly as ``ProbabilisticFilterSVMModel.predict`` does from one
    file's totals (probabilistic_filter_svm_model.py:209-223)."""
    import torch.distributed as dist

    from .distributed import allreduce_totals

    out = genus_then_species(genus_model, species_model, sequence_input, threshold, step, predict=False)
    names = out["labels"]
    local = np.array([out["total_hits"][n] for n in names] + [out["total_kmers"], int(out["kept"].sum()), len(out["kept"])], dtype=np.int64)
    glob = allreduce_totals(local, group).cpu().numpy() if dist.is_initialized() else local
    total_kmers = int(glob[len(names)])
    out["global_total_hits"] = {n: int(v) for n, v in zip(names, glob[: len(names)])}
    out["global_total_scores"] = {n: round(v / total_kmers, 2) for n, v in out["global_total_hits"].items()} if total_kmers else {}
    out["global_kept"], out["global_records"] = int(glob[len(names) + 1]), int(glob[len(names) + 2])
    out["prediction"] = None
    if hasattr(species_model, "_get_svm") and total_kmers:
        x = [list(dict(sorted(out["global_total_scores"].items())).values())]
        out["prediction"] = str(species_model._get_svm(None).predict(x)[0])
    return out
