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
from ._abi import XS_U32
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


def genus_then_species(genus_model, species_model, sequence_input, threshold: float = 0.7, step: int = 1,
                       predict: bool = True) -> dict:
    """Score every record against the genus filter, keep those reaching ``threshold`` and classify them with the
    species model; returns per-record genus hits, the kept mask, read-level species calls for the kept records,
    file-level species totals / scores and (for an SVM species model) the prediction."""
    if threshold < 0 or threshold > 1:
        raise ValueError("The filter threshold must be between 0 and 1.")
    batch = sequence_input if isinstance(sequence_input, SequenceBatch) else genus_model._to_batch(sequence_input)
    genus_model._check_lengths(batch)
    bf = genus_model.bf.filter
    ix = species_model.index.index
    if bf.device != ix.device:
        raise ValueError("genus filter and species index must live on the same GPU")
    dev = torch.device("cuda", ix.device)
    n = len(batch)
    k = genus_model.k
    with torch.cuda.device(dev):
        stream = torch.cuda.current_stream(dev)
        d_bases = torch.from_numpy(batch.bases).to(dev, non_blocking=True)
        d_begin = torch.from_numpy(batch.begin.view(np.int64)).to(dev, non_blocking=True)
        d_end = torch.from_numpy(batch.end.view(np.int64)).to(dev, non_blocking=True)
        # ---- stage 1: genus Bloom filter
        d_hits = torch.empty(n, dtype=torch.int32, device=dev)
        bf.query_device(d_bases.data_ptr(), batch.bases.size, d_begin.data_ptr(), d_end.data_ptr(), n, step, d_hits.data_ptr(),
                        stream.cuda_stream)
        lengths = batch.lengths
        num_kmers = -((lengths - k + 1) // -step)
        max_nk = int(num_kmers.max()) if n else 0
        if max_nk <= (1 << 20):
            table = torch.from_numpy(min_hits_table(max_nk, threshold)).to(dev)
            d_keep = d_hits.to(torch.int64) >= table[torch.from_numpy(num_kmers).to(dev)]
        else:   # a few very long records: evaluate the rounded score on the host
            h = d_hits.cpu().numpy()
            d_keep = torch.from_numpy(np.fromiter((round(int(a) / int(b), 2) >= threshold for a, b in zip(h, num_kmers)),
                                                  dtype=bool, count=n)).to(dev)
        # ---- stage 2: species index on the kept records (their offsets into the same device buffer)
        kept_idx = torch.nonzero(d_keep).squeeze(1)
        m = int(kept_idx.numel())
        s_begin = d_begin[kept_idx].contiguous()
        s_end = d_end[kept_idx].contiguous()
        d_counts = torch.empty((m, ix.n_docs), dtype=torch.uint32, device=dev)
        best = torch.empty(m, dtype=torch.int32, device=dev)
        cnt = torch.empty(m, dtype=torch.int32, device=dev)
        nb = torch.empty(m, dtype=torch.int32, device=dev)
        totals = torch.zeros(ix.n_docs, dtype=torch.int64, device=dev)
        if m:
            ix.query_device(d_bases.data_ptr(), batch.bases.size, s_begin.data_ptr(), s_end.data_ptr(), m, step, XS_U32,
                            d_counts.data_ptr(), stream.cuda_stream)
            engine.scores_reduce_device(d_counts.data_ptr(), m, ix.n_docs, XS_U32, ix.device, best.data_ptr(), cnt.data_ptr(),
                                        nb.data_ptr(), totals.data_ptr(), stream.cuda_stream)
        keep = d_keep.cpu().numpy()
        genus_hits = d_hits.cpu().numpy().astype(np.uint32)
        totals_h = totals.cpu().numpy()
        out = {
            "batch": batch, "genus_label": next(iter(genus_model.display_names)), "genus_hits": genus_hits, "num_kmers": num_kmers,
            "kept": keep, "kept_index": kept_idx.cpu().numpy(), "labels": ix.names,
            "best": best.cpu().numpy().view(np.uint32), "best_hits": cnt.cpu().numpy().view(np.uint32),
            "ambiguous": nb.cpu().numpy() > 1,
        }
    total_kmers = int(num_kmers[keep].sum())
    out["total_hits"] = {name: int(v) for name, v in zip(ix.names, totals_h)}
    out["total_scores"] = {name: round(v / total_kmers, 2) for name, v in out["total_hits"].items()} if total_kmers else {}
    out["total_kmers"] = total_kmers
    out["prediction"] = None
    if predict and hasattr(species_model, "_get_svm") and total_kmers:
        x = [list(dict(sorted(out["total_scores"].items())).values())]
        out["prediction"] = str(species_model._get_svm(None).predict(x)[0])
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
