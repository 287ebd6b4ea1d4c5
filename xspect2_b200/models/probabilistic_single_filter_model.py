"""Genus-level model: one rbloom Bloom filter queried on the GPU.

API of the reference's ``ProbabilisticSingleFilterModel`` (models/probabilistic_single_filter_model.py:16-180).
The reference walks every k-mer in Python (``_generate_kmers`` :161-180, ``kmer in self.bf`` :122-124); here the
canonical form ``min(kmer, revcomp)``, XXH3-64, the 128-bit LCG and the bit probes all run in one kernel.
"""

from __future__ import annotations

import json
from math import ceil
from pathlib import Path

import numpy as np

from .. import engine, seqio
from ..seqio import Seq, SeqRecord, SequenceBatch
from .probabilistic_filter_model import BatchHits, ProbabilisticFilterModel, default_device


class ProbabilisticSingleFilterModel(ProbabilisticFilterModel):
    """Probabilistic filter model with a single Bloom filter (e.g. genus membership)."""

    _exclude_ids_apply = False

    def __init__(
        self,
        k: int,
        model_display_name: str,
        author: str | None,
        author_email: str | None,
        model_type: str,
        base_path: Path,
        fpr: float = 0.01,
        training_accessions: list[str] | None = None,
    ) -> None:
        super().__init__(
            k=k,
            model_display_name=model_display_name,
            author=author,
            author_email=author_email,
            model_type=model_type,
            base_path=base_path,
            fpr=fpr,
            num_hashes=1,
            training_accessions=training_accessions,
        )
        self.bf = None

    def fit(self, file_path: Path, display_name: str, training_accessions: list[str] | None = None, device: int | None = None) -> None:
        """Build ``<slug>/filter.bloom`` from every k-mer of the records of ``file_path`` (reference :63-96: a
        per-k-mer Python loop around rbloom.add) with one kernel launch; sized like the reference:
        ``Bloom(total_length - k + 1, fpr)``."""
        from ..file_io import get_record_iterator
        self.training_accessions = training_accessions
        get_record_iterator(file_path)     # same path / format errors as the reference
        batch = SequenceBatch.from_file(file_path)
        num_kmers = int(batch.lengths.sum()) - self.k + 1
        bloom_path = self.base_path / self.slug() / "filter.bloom"
        bloom_path.parent.mkdir(parents=True, exist_ok=True)
        dev = default_device() if device is None else device
        engine.build_bloom(bloom_path, self.k, num_kmers, self.fpr, batch.bases, batch.begin, batch.end, device=dev)
        self.display_names[file_path.stem] = display_name
        self.bf = engine.Bloom.load(str(bloom_path), self.k, device=dev)

    def calculate_hits(self, sequence: Seq | SeqRecord, exclude_ids=None, step: int = 1) -> dict:
        """``{first display name key: number of sampled k-mers found in the filter}`` (reference :98-125)."""
        if seqio.is_record(sequence):
            sequence = sequence.seq
        if not seqio.is_seq(sequence):
            raise ValueError("Invalid sequence, must be a Bio.Seq object")
        if not len(sequence) > self.k:
            raise ValueError("Invalid sequence, must be longer than k")
        return {next(iter(self.display_names)): self.bf.filter.hits(str(sequence), step)}

    def predict_arrays(self, sequence_input, step: int = 1) -> BatchHits:
        batch = sequence_input if isinstance(sequence_input, SequenceBatch) else self._to_batch(sequence_input)
        self._check_lengths(batch)
        hits = self.bf.filter.query(batch.bases, batch.begin, batch.end, step)
        num_kmers = -((batch.lengths - self.k + 1) // -step)
        return BatchHits(batch.ids, [next(iter(self.display_names))], np.asarray(hits).reshape(-1, 1), num_kmers, step)

    @staticmethod
    def load(path: Path, device: int | None = None) -> "ProbabilisticSingleFilterModel":
        """Read ``<slug>.json`` and put ``<slug>/filter.bloom`` into HBM (reference :127-159)."""
        with open(path, "r", encoding="utf-8") as file:
            model_json = json.loads(file.read())
        model = ProbabilisticSingleFilterModel(
            model_json["k"],
            model_json["model_display_name"],
            model_json["author"],
            model_json["author_email"],
            model_json["model_type"],
            path.parent,
            fpr=model_json["fpr"],
            training_accessions=model_json["training_accessions"],
        )
        model.display_names = model_json["display_names"]
        bloom_path = model.base_path / model.slug() / "filter.bloom"
        model.bf = engine.Bloom.load(str(bloom_path), model.k, device=default_device() if device is None else device)
        return model

    def _generate_kmers(self, sequence: Seq, step: int = 1):
        """The canonical k-mers the filter is probed with: ``min(kmer, revcomp)`` at positions 0, step, ...
        (reference :161-180).  Host-side restatement for inspection; the GPU kernel does this itself."""
        for i in range(ceil((len(sequence) - self.k + 1) / step)):
            kmer = sequence[i * step : i * step + self.k]
            yield str(min(kmer, str(kmer.reverse_complement())))
