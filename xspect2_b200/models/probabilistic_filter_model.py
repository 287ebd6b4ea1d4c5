"""Species-level model: a COBS classic index queried on the GPU.

API of the reference's ``ProbabilisticFilterModel`` (models/probabilistic_filter_model.py:28-601): same
constructor, ``to_dict`` / ``slug`` / ``save`` / ``load`` / ``calculate_hits`` / ``predict`` / ``_count_kmers``,
same model directory layout (``<slug>.json`` + ``<slug>/index.cobs_classic``), same exceptions.  What differs
is underneath: ``predict`` does not loop over records calling ``cobs_index.Search.search``
(:291-310, :227); all records of the input go to the GPU in one batched query of the HBM-resident index.
"""

from __future__ import annotations

import json
import os
import warnings
from math import ceil
from pathlib import Path
from typing import Any

import numpy as np

from .. import engine, seqio
from ..file_io import get_record_iterator
from ..model_management import slugify
from ..seqio import Seq, SeqRecord, SequenceBatch
from .result import ColumnarModelResult, ModelResult


def _abi_kind(kind: str) -> int:
    from .._abi import XS_COBS_CLASSIC, XS_COBS_COMPACT
    return XS_COBS_CLASSIC if kind == "classic" else XS_COBS_COMPACT


def default_device() -> int:
    """GPU the models load onto: ``XSPECT_B200_DEVICE`` or, under torchrun, ``LOCAL_RANK``; else 0."""
    return int(os.environ.get("XSPECT_B200_DEVICE", os.environ.get("LOCAL_RANK", 0)))


class BatchHits:
    """Columnar result of one batched query: ``counts[i, j]`` = hits of record ``ids[i]`` on document
    ``names[j]``; ``num_kmers[i]`` = sampled windows of record i.  ``to_hits()`` yields the nested dicts the
    reference's ModelResult holds."""

    def __init__(self, ids: list[str], names: list[str], counts: np.ndarray, num_kmers: np.ndarray, step: int):
        self.ids, self.names, self.counts, self.num_kmers, self.step = ids, names, counts, num_kmers, step

    def total_hits(self) -> dict[str, int]:
        t = self.counts.sum(axis=0, dtype=np.int64)
        return {n: int(v) for n, v in zip(self.names, t)}

    def argmax(self) -> tuple[np.ndarray, np.ndarray]:
        """Per record: index of the best document and whether the maximum is tied (ties = ambiguous, the
        read-level rule of scripts/benchmark/main.nf:417-436)."""
        best = self.counts.argmax(axis=1)
        mx = self.counts.max(axis=1)
        return best, (self.counts == mx[:, None]).sum(axis=1) > 1

    def to_hits(self) -> tuple[dict[str, dict[str, int]], dict[str, int]]:
        order = engine.result_order_batch(self.counts)
        names = np.array(self.names, dtype=object)
        hits, num_kmers = {}, {}
        sorted_counts = np.take_along_axis(np.asarray(self.counts), order.astype(np.int64), axis=1)
        for i, rid in enumerate(self.ids):
            hits[rid] = dict(zip(names[order[i]].tolist(), sorted_counts[i].tolist()))
            num_kmers[rid] = int(self.num_kmers[i])
        return hits, num_kmers


class ProbabilisticFilterModel:
    """Probabilistic filter model for sequence data (COBS classic index, one document per label)."""

    _exclude_ids_apply = True   # the single-filter model ignores exclude_ids (probabilistic_single_filter_model.py:98-125)

    def __init__(
        self,
        k: int,
        model_display_name: str,
        author: str | None,
        author_email: str | None,
        model_type: str,
        base_path: Path,
        fpr: float = 0.01,
        num_hashes: int = 7,
        training_accessions: dict[str, list[str]] | None = None,
    ) -> None:
        if k < 1:
            raise ValueError("Invalid k value, must be greater than 0")
        if not model_display_name:
            raise ValueError("Invalid filter display name, must be a non-empty string")
        if not model_type:
            raise ValueError("Invalid filter type, must be a non-empty string")
        if not isinstance(base_path, Path):
            raise ValueError("Invalid base path, must be a pathlib.Path object")
        self.k = k
        self.model_display_name = model_display_name
        self.author = author
        self.author_email = author_email
        self.model_type = model_type
        self.base_path = base_path
        self.display_names = {}
        self.fpr = fpr
        self.num_hashes = num_hashes
        self.index = None
        self.training_accessions = training_accessions

    # ------------------------------------------------------------------ metadata / persistence
    def get_cobs_index_path(self) -> str:
        return str(self.base_path / self.slug() / "index.cobs_classic")

    def to_dict(self) -> dict:
        return {
            "model_slug": self.slug(),
            "k": self.k,
            "model_display_name": self.model_display_name,
            "author": self.author,
            "author_email": self.author_email,
            "model_type": self.model_type,
            "model_class": self.__class__.__name__,
            "display_names": self.display_names,
            "fpr": self.fpr,
            "num_hashes": self.num_hashes,
            "training_accessions": self.training_accessions,
        }

    def slug(self) -> str:
        return slugify(self.model_display_name + "-" + str(self.model_type))

    @staticmethod
    def _gather_documents(files: list[Path]):
        """All records of the given files as one buffer; sequence i belongs to document ``seq_doc[i]`` = its file."""
        parts, begins, ends, docs = [], [], [], []
        off = 0
        for d, file in enumerate(files):
            b = SequenceBatch.from_file(file)
            parts.append(b.bases)
            begins.append(b.begin + np.uint64(off))
            ends.append(b.end + np.uint64(off))
            docs.append(np.full(len(b), d, dtype=np.uint32))
            off += b.bases.size
        cat = lambda xs, dt: np.concatenate(xs) if xs else np.zeros(0, dt)   # noqa: E731
        return cat(parts, np.uint8), cat(begins, np.uint64), cat(ends, np.uint64), cat(docs, np.uint32)

    def fit(self, dir_path: Path, display_names: dict | None = None, training_accessions: dict[str, list[str]] | None = None,
            device: int | None = None) -> None:
        """Build ``<slug>/index.cobs_classic`` from the fasta/fastq files of a directory, one document per file
        (reference :138-194, which hands the files to cobs classic_construct_list) — here the construction runs on
        the GPU (xs_cobs_build).  Document names are the file names up to the first '.', as cobs derives them;
        files are taken in sorted order."""
        if display_names is None:
            display_names = {}
        if not isinstance(dir_path, Path):
            raise ValueError("Invalid directory path, must be a pathlib.Path object")
        if not dir_path.exists():
            raise ValueError("Directory path does not exist")
        if not dir_path.is_dir():
            raise ValueError("Directory path must be a directory")
        self.training_accessions = training_accessions
        from ..definitions import fasta_endings, fastq_endings
        files = [f for f in sorted(dir_path.iterdir()) if f.is_file() and f.suffix[1:] in fasta_endings + fastq_endings]
        if not files:
            raise ValueError("No valid files found in directory. Must be fasta or fastq")
        names = []
        for file in files:
            doc_name = file.stem.split(".")[0]
            self.display_names[doc_name] = display_names.get(file.stem, file.stem)
            names.append(doc_name)
        bases, begin, end, seq_doc = self._gather_documents(files)
        index_path = Path(self.get_cobs_index_path())
        index_path.parent.mkdir(exist_ok=True, parents=True)
        dev = default_device() if device is None else device
        engine.build_cobs(index_path, _abi_kind("classic"), self.k, self.num_hashes, self.fpr, names, bases, begin, end, seq_doc, device=dev)
        self._open_index(dev)

    def save(self) -> None:
        json_path = self.base_path / f"{self.slug()}.json"
        (self.base_path / self.slug()).mkdir(exist_ok=True, parents=True)
        with open(json_path, "w", encoding="utf-8") as file:
            file.write(json.dumps(self.to_dict(), indent=4))

    @staticmethod
    def load(path: Path, device: int | None = None) -> "ProbabilisticFilterModel":
        """Read ``<slug>.json`` and put ``<slug>/index.cobs_classic`` into HBM (reference :351-391)."""
        with open(path, "r", encoding="utf-8") as file:
            model_json = json.loads(file.read())
        model = ProbabilisticFilterModel(
            model_json["k"],
            model_json["model_display_name"],
            model_json["author"],
            model_json["author_email"],
            model_json["model_type"],
            path.parent,
            model_json["fpr"],
            model_json["num_hashes"],
            model_json["training_accessions"],
        )
        model.display_names = model_json["display_names"]
        model._open_index(device)
        return model

    def _open_index(self, device: int | None) -> None:
        index_path = self.get_cobs_index_path()
        if not Path(index_path).exists():
            raise FileNotFoundError(f"Index file not found at {index_path}")
        self.index = engine.Search(index_path, True, device=default_device() if device is None else device)
        if self.index.index.k != self.k:
            raise ValueError(f"Index term size {self.index.index.k} differs from the model's k {self.k}")

    # ------------------------------------------------------------------ scoring
    def calculate_hits(self, sequence: Seq, exclude_ids: list[str] | None = None, step: int = 1) -> dict:
        """Hits of one sequence per document, in cobs result order (reference :196-235)."""
        if not seqio.is_seq(sequence):
            raise ValueError("Invalid sequence, must be a Bio.Seq or a Bio.SeqRecord object")
        if not len(sequence) > self.k:
            raise ValueError("Invalid sequence, must be longer than k")
        result_dict = self._convert_cobs_result_to_dict(self.index.search(str(sequence), step=step))
        if exclude_ids:
            return {doc: score for doc, score in result_dict.items() if doc not in exclude_ids}
        return result_dict

    def _to_batch(self, sequence_input, keep_records: bool = False) -> SequenceBatch:
        if seqio.is_record(sequence_input):
            sequence_input = [sequence_input]
        if self._is_sequence_list(sequence_input) or self._is_sequence_iterator(sequence_input):
            def checked(records):
                for r in records:
                    if not seqio.is_seq(r.seq):
                        raise ValueError("Invalid sequence, must be a Bio.Seq or a Bio.SeqRecord object")
                    yield r
            return SequenceBatch.from_records(checked(sequence_input), keep_records=keep_records)
        if isinstance(sequence_input, Path):
            get_record_iterator(sequence_input)  # same path / format errors as the reference
            return SequenceBatch.from_file(sequence_input)
        raise ValueError(
            "Invalid sequence input, must be a Seq object, a list of Seq objects, a"
            " SeqIO FastaIterator, a SeqIO FastqPhredIterator, or a Path object to a"
            " fasta/fastq file"
        )

    def _check_lengths(self, batch: SequenceBatch) -> None:
        if len(batch) and int(batch.lengths.min()) <= self.k:
            raise ValueError("Invalid sequence, must be longer than k")

    def predict_arrays(self, sequence_input, step: int = 1) -> BatchHits:
        """Batched scoring with a columnar result (no per-record Python objects)."""
        batch = sequence_input if isinstance(sequence_input, SequenceBatch) else self._to_batch(sequence_input)
        self._check_lengths(batch)
        ix = self.index.index
        counts = ix.query(batch.bases, batch.begin, batch.end, step=step)
        num_kmers = -((batch.lengths - self.k + 1) // -step)
        return BatchHits(batch.ids, ix.names, counts, num_kmers, step)

    def predict_summary(self, sequence_input, step: int = 1) -> dict:
        """Read-level calls and file-level scores without the per-record dictionaries: the device takes the argmax
        per record (ties flagged, the rule of scripts/benchmark/main.nf:417-436) and the per-document totals.
        Returns ``{"batch", "labels", "best", "best_hits", "ambiguous", "num_kmers", "total_hits", "total_scores"}``
        (record ids are ``result["batch"].ids``, decoded on first use);
        ``total_scores`` equals ``ModelResult.get_scores()["total"]`` for input without duplicate ids."""
        if isinstance(sequence_input, Path):
            streamed = self._predict_summary_streamed(sequence_input, step)
            if streamed is not None:
                return streamed
        batch = sequence_input if isinstance(sequence_input, SequenceBatch) else self._to_batch(sequence_input)
        self._check_lengths(batch)
        ix = self.index.index
        best, cnt, nb, totals = ix.classify(batch.bases, batch.begin, batch.end, step)
        num_kmers = -((batch.lengths - self.k + 1) // -step)
        total_kmers = int(num_kmers.sum())
        total_hits = {n: int(v) for n, v in zip(ix.names, totals)}
        return {
            "batch": batch, "labels": ix.names, "best": best, "best_hits": cnt, "ambiguous": nb > 1, "num_kmers": num_kmers,
            "total_hits": total_hits,
            "total_scores": {n: round(v / total_kmers, 2) for n, v in total_hits.items()} if total_kmers else {},
        }

    def _predict_summary_streamed(self, path: Path, step: int):
        """``predict_summary`` of a file with parsing, copies and kernels overlapped (``xs_cobs_classify_file``); None when
        the streaming reader does not take the file (wrapped FASTQ, unknown extension: the two-pass reader decides)."""
        from ..definitions import fasta_endings, fastq_endings
        suffix = path.suffix[1:]
        fmt = 2 if suffix in fastq_endings else 1 if suffix in fasta_endings else 0
        if not fmt or not path.is_file():
            return None
        ix = self.index.index
        try:
            r = ix.classify_file(path, fmt, step)
        except ValueError:
            return None
        if r["n_short"]:
            raise ValueError("Invalid sequence, must be longer than k")
        n = r["best"].size
        begin = np.zeros(n, np.uint64)
        lengths = r["seq_len"].astype(np.int64)
        batch = SequenceBatch(None, np.zeros(0, np.uint8), begin, r["seq_len"], None, r["id_buf"], r["id_end"])
        num_kmers = -((lengths - self.k + 1) // -step)
        total_kmers = int(num_kmers.sum())
        total_hits = {name: int(v) for name, v in zip(ix.names, r["totals"])}
        return {
            "batch": batch, "labels": ix.names, "best": r["best"], "best_hits": r["best_hits"], "ambiguous": r["n_best"] > 1,
            "num_kmers": num_kmers, "total_hits": total_hits,
            "total_scores": {name: round(v / total_kmers, 2) for name, v in total_hits.items()} if total_kmers else {},
            "streamed": {"parse_s": r["parse_s"], "total_s": r["total_s"]},
        }

    def predict(
        self,
        sequence_input: SeqRecord | list[SeqRecord] | Any | Path,
        exclude_ids: list[str] = None,
        step: int = 1,
        display_name: bool = False,
        validation: bool = False,
    ) -> ModelResult:
        """ModelResult for a record, a list of records, a record iterator or a fasta/fastq path
        (reference :237-331).  Later records with an id seen before overwrite earlier ones, one record not
        longer than k aborts the call — as in the reference's loop."""
        if validation:
            # the reference moves reads its alignment-based filter rejects into result.misclassified, which changes
            # hits, totals and the SVM prediction; returning unfiltered hits under the same flag would be a silently
            # different result
            raise NotImplementedError(
                "validation=True: the alignment-based misclassification filter (minimap2 mapping + Ripley's K, "
                "reference probabilistic_filter_model.py:508-601) is outside the GPU scoring path"
            )
        batch = self._to_batch(sequence_input)
        self._check_lengths(batch)
        cols = self.predict_arrays(batch, step)
        if display_name:
            doc_keys = [f"{key} -{self.display_names.get(key, 'Unknown').replace(self.model_display_name, '', 1)}" for key in cols.names]
        else:
            doc_keys = list(cols.names)
        drop = exclude_ids if (exclude_ids and self._exclude_ids_apply) else ()
        include = np.array([name not in drop for name in cols.names], dtype=bool)
        return ColumnarModelResult(self.slug(), batch.ids, cols.names, doc_keys, include, cols.counts, cols.num_kmers,
                                   sparse_sampling_step=step)

    def _convert_cobs_result_to_dict(self, cobs_result) -> dict:
        return {r.doc_name: r.score for r in cobs_result}

    def _count_kmers(self, sequence_input, step: int = 1) -> int:
        """``ceil((len - k + 1) / step)`` summed over the input (reference :411-469); windows with N count."""
        if seqio.is_seq(sequence_input):
            return self._count_kmers([sequence_input], step=step)
        if seqio.is_record(sequence_input):
            return self._count_kmers(sequence_input.seq, step=step)
        is_sequence_list = isinstance(sequence_input, list) and all(seqio.is_seq(s) for s in sequence_input)
        is_iterator = self._is_sequence_iterator(sequence_input)
        if is_sequence_list or is_iterator:
            total = 0
            for item in sequence_input:
                seq = item.seq if is_iterator else item
                total += ceil((len(seq) - self.k + 1) / step)
            return total
        raise ValueError(
            "Invalid sequence input, must be a Seq object, a list of Seq objects, a"
            " SeqIO FastaIterator, or a SeqIO FastqPhredIterator"
        )

    def _is_sequence_list(self, sequence_input: Any) -> bool:
        return isinstance(sequence_input, list) and all(seqio.is_record(s) for s in sequence_input)

    def _is_sequence_iterator(self, sequence_input: Any) -> bool:
        return seqio.is_record_iterator(sequence_input)

    def detecting_misclassification(self, hits, seq_records, min_reads: int = 10):
        """Alignment-based post-filter of the reference (:508-601; mappy + pysam): not part of this path."""
        raise NotImplementedError("alignment-based misclassification detection is outside the GPU scoring path")
