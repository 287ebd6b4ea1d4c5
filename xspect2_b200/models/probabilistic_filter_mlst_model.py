"""MLST scheme model: one COBS compact index per locus, queried on the GPU.

API of the reference's ``ProbabilisticFilterMlstSchemeModel``
(models/probabilistic_filter_mlst_model.py:22-449): same constructor / ``to_dict`` / ``slug`` / ``load`` /
``calculate_hits`` / ``predict`` / ``get_cobs_result`` / ``sequence_splitter`` / ``has_sufficient_score`` and the
same result structure.  The reference searches every chunk of a long sequence separately per locus
(:236-243, ~1e3 ``Search.search`` calls per locus and genome); here all chunks of a sequence go to the GPU as
overlapping segments of one buffer, one batched query per locus, and the per-chunk ``score > 50`` filter and
per-allele sums (:243-256) run over the returned count matrix.
"""

from __future__ import annotations

import json
from pathlib import Path

import numpy as np

from .. import engine, seqio
from ..file_io import get_record_iterator
from ..handlers.pubmlst import PubMLSTHandler
from ..model_management import slugify
from ..seqio import Seq, SeqRecord
from .mlst_result import MlstResult
from .probabilistic_filter_model import ProbabilisticFilterModel, default_device


class ProbabilisticFilterMlstSchemeModel(ProbabilisticFilterModel):
    def __init__(
        self,
        k: int,
        model_display_name: str,
        base_path: Path,
        scheme_url: str,
        organism: str,
        fpr: float = 0.001,
        num_hashes: int = 1,
        author: str | None = None,
        author_email: str | None = None,
        model_type: str = "MLST",
    ) -> None:
        super().__init__(k, model_display_name, author, author_email, model_type, base_path, fpr, num_hashes, None)
        self.organism = organism
        self.scheme_url = scheme_url
        self.loci = {}
        self.avg_locus_bp_size = []
        self.indices = []
        self.pubmlst_handler = None   # injectable; defaults to PubMLSTHandler() at the call site

    def to_dict(self) -> dict:
        return super().to_dict() | {
            "organism": self.organism,
            "scheme_url": self.scheme_url,
            "loci": self.loci,
            "average_locus_base_pair_size": self.avg_locus_bp_size,
        }

    def slug(self) -> str:
        return slugify(self.organism + "-" + self.model_display_name + "-" + self.model_type)

    def get_cobs_index_path(self, locus: str) -> Path:
        return self.base_path / self.slug() / f"{locus}.cobs_compact"

    def fit(self, scheme_path: Path, device: int | None = None) -> None:
        """One compact index per locus directory of ``scheme_path``, one document per allele file (reference
        :100-142 via cobs compact_construct_list), constructed on the GPU.  ``avg_locus_bp_size`` is, like there, the
        length of the first allele file's record; loci and files are taken in sorted order."""
        if not scheme_path.exists():
            raise ValueError("Scheme not found. Please make sure to download the schemes prior!")
        from ..definitions import fasta_endings, fastq_endings
        from ..seqio import SequenceBatch
        from .probabilistic_filter_model import _abi_kind
        dev = default_device() if device is None else device
        for locus_path in sorted(scheme_path.iterdir()):
            if not locus_path.is_dir():
                continue
            locus = locus_path.name
            files = [p for p in sorted(locus_path.iterdir()) if p.is_file() and p.suffix[1:] in fasta_endings + fastq_endings]
            self.loci[locus] = sum(1 for p in files if p.suffix == ".fasta")
            first = next((p for p in files if p.suffix == ".fasta"), None)
            if first is None:
                raise ValueError(f"No allele fasta files found for locus {locus}")
            self.avg_locus_bp_size.append(int(SequenceBatch.from_file(first).lengths[0]))
            names = [p.stem.split(".")[0] for p in files]
            bases, begin, end, seq_doc = self._gather_documents(files)
            cobs_path = self.get_cobs_index_path(locus)
            cobs_path.parent.mkdir(exist_ok=True, parents=True)
            engine.build_cobs(cobs_path, _abi_kind("compact"), self.k, self.num_hashes, self.fpr, names, bases, begin, end, seq_doc, device=dev)
            self.indices.append(engine.Search(str(cobs_path), False, device=dev))

    def save(self) -> None:
        json_path = self.base_path / f"{self.slug()}.json"
        json_path.parent.mkdir(exist_ok=True, parents=True)
        with open(json_path, "w", encoding="utf-8") as file:
            file.write(json.dumps(self.to_dict(), indent=4))

    @staticmethod
    def load(path: Path, device: int | None = None) -> "ProbabilisticFilterMlstSchemeModel":
        if not path.exists():
            raise FileNotFoundError(f"Model JSON not found at {path}")
        with open(path, "r", encoding="utf-8") as file:
            model_json = json.loads(file.read())
        model = ProbabilisticFilterMlstSchemeModel(
            model_json["k"],
            model_json["model_display_name"],
            path.parent,
            model_json["scheme_url"],
            model_json["organism"],
            model_json["fpr"],
            model_json["num_hashes"],
            model_json.get("author"),
            model_json.get("author_email"),
            model_json.get("model_type"),
        )
        model.avg_locus_bp_size = model_json.get("average_locus_base_pair_size", [])
        model.loci = model_json.get("loci", {})
        dev = default_device() if device is None else device
        for locus in model.loci.keys():
            index_path = model.get_cobs_index_path(locus)
            if not index_path.exists():
                raise FileNotFoundError(f"Index file not found at {index_path}")
            model.indices.append(engine.Search(str(index_path), False, device=dev))
        return model

    # ------------------------------------------------------------------ chunking
    def _chunk_layout(self, n: int, allele_len: int) -> tuple[int, list[int], int | None]:
        """Chunk starts of ``sequence_splitter`` for a sequence of length n: (chunk length, starts of the full
        chunks, start of the remainder or None)."""
        if n < 1000000:
            sub = allele_len
        elif n < 10000000:
            sub = allele_len * 10
        else:
            sub = allele_len * 100
        stride = sub - self.k + 1
        starts = list(range(0, n - sub + 1, stride)) if (n >= sub and stride > 0) else []
        nxt = starts[-1] + stride if starts else 0
        return sub, starts, (nxt if nxt < n else None)

    def sequence_splitter(self, input_sequence: str, allele_len: int) -> list[str]:
        """Substrings of ``allele_len`` (x10 from 1 Mbp, x100 from 10 Mbp) overlapping by k-1 bases; a remainder
        shorter than k is appended to the last substring (reference :382-426)."""
        n = len(input_sequence)
        sub, starts, rem = self._chunk_layout(n, allele_len)
        chunks = [input_sequence[s : s + sub] for s in starts]
        if rem is not None:
            rest = input_sequence[rem:]
            if len(rest) < self.k:
                chunks[-1] += rest
            else:
                chunks.append(rest)
        return chunks

    def _chunk_segments(self, seq_bytes: np.ndarray, allele_len: int):
        """The same chunks as byte segments of one buffer.  The one non-contiguous case — a short remainder
        glued to the last chunk, which repeats the k-1 overlap bases — is materialised at the buffer's end."""
        n = seq_bytes.size
        sub, starts, rem = self._chunk_layout(n, allele_len)
        begin = np.asarray(starts, dtype=np.uint64)
        end = begin + np.uint64(sub)
        bases = seq_bytes
        if rem is not None:
            if n - rem < self.k:
                glued = np.concatenate((seq_bytes[starts[-1] : starts[-1] + sub], seq_bytes[rem:]))
                bases = np.concatenate((seq_bytes, glued))
                begin[-1], end[-1] = n, n + glued.size
            else:
                begin = np.concatenate((begin, [rem])).astype(np.uint64)
                end = np.concatenate((end, [n])).astype(np.uint64)
        return bases, begin, end

    # ------------------------------------------------------------------ scoring
    def get_cobs_result(self, cobs_result, kmer_threshold: bool) -> dict:
        """Allele -> score of one search, keeping scores > 50 only when ``kmer_threshold`` (reference :362-380)."""
        return {r.doc_name: r.score for r in cobs_result if not kmer_threshold or r.score > 50}

    def _score_records(self, seqs: list[np.ndarray], step: int) -> list[list[dict]]:
        """``out[record][locus]`` = allele -> score in the reference's dict order, for all records and all loci in
        one ``xs_mlst_query`` call: records >= 10000 bp are chunked (:236-256; chunk scores > 50 summed per allele,
        first-appearance order, stable sort by -score), shorter ones are searched whole (:272-286).  Chunking,
        threshold and sums run on the device; only the few chunk rows that pass the threshold come back."""
        sizes = np.fromiter((sb.size for sb in seqs), dtype=np.uint64, count=len(seqs))
        end = np.cumsum(sizes, dtype=np.uint64)
        begin = end - sizes
        bases = np.concatenate(seqs) if len(seqs) > 1 else seqs[0]
        indices = [search.index for search in self.indices]
        res = engine.mlst_query(indices, self.avg_locus_bp_size, bases, begin, end, step)
        out = [[None] * len(indices) for _ in seqs]
        for li, ix in enumerate(indices):
            names = ix.names
            for ri in range(len(seqs)):
                docs, scores = res[li][ri]
                out[ri][li] = {names[d]: v for d, v in zip(docs.tolist(), scores.tolist())}
        return out

    def _assemble(self, seq_len: int, per_locus: list[dict], limit: bool, limit_number: int) -> list[dict]:
        """The result structure of calculate_hits from per-locus allele scores (reference :257-303)."""
        loci = list(self.loci.keys())
        result_dict = {}
        highest_results = {}
        for counter, scores in enumerate(per_locus):
            if seq_len >= 10000:
                sorted_counts = dict(list(scores.items())[:limit_number]) if limit else scores
                if not sorted_counts:
                    # the reference replaces the whole result dict by this message (:259-260)
                    result_dict = "A Strain type could not be detected because of no kmer matches!"
                    highest_results[loci[counter]] = {"N/A": 0}
                else:
                    first_key = next(iter(sorted_counts))
                    result_dict[loci[counter]] = sorted_counts
                    highest_results[loci[counter]] = {first_key: sorted_counts[first_key]}
            else:
                result = dict(sorted(scores.items(), key=lambda x: -x[1])[:limit_number]) if limit else scores
                result_dict[loci[counter]] = result
                first_key, highest_result = next(iter(result.items()))
                highest_results[loci[counter]] = {first_key: highest_result}
        if not self.has_sufficient_score(highest_results, self.avg_locus_bp_size):
            highest_results["Attention:"] = "This strain type is not reliable due to low kmer hit rates!"
        else:
            handler = self.pubmlst_handler or PubMLSTHandler()
            flattened = {locus: int(list(allele_id.keys())[0].split("_")[-1]) for locus, allele_id in highest_results.items()}
            highest_results["ST_Name"] = handler.get_strain_type_name(flattened, self.scheme_url)
        return [{"Strain type": highest_results}, {"All results": result_dict}]

    def _check_sequence(self, sequence) -> None:
        if not seqio.is_seq(sequence):
            raise ValueError("Invalid sequence, must be a Bio.Seq object")
        if not len(sequence) > self.k:
            raise ValueError("Invalid sequence, must be longer than k")
        if not self.indices:
            raise ValueError("The model has not been trained yet")

    def calculate_hits(self, sequence: Seq, step: int = 1, limit: bool = False, limit_number: int = 5) -> list[dict]:
        """Best allele per locus and all allele scores for one sequence (reference :192-303)."""
        self._check_sequence(sequence)
        seq_bytes = np.frombuffer(str(sequence).encode("ascii", "replace"), dtype=np.uint8)
        return self._assemble(len(sequence), self._score_records([seq_bytes], step)[0], limit, limit_number)

    def _predict_records(self, records, step: int, limit: bool) -> dict:
        """All records of an input in one batched query per locus; same per-record results and dict order as the
        reference's record loop (:353-355), the first invalid record raises like there."""
        ids, seqs = [], []
        for record in records:
            self._check_sequence(record.seq)
            ids.append(record.id)
            seqs.append(np.frombuffer(str(record.seq).encode("ascii", "replace"), dtype=np.uint8))
        hits = {}
        if seqs:
            scored = self._score_records(seqs, step)
            for rid, sb, per_locus in zip(ids, seqs, scored):
                hits[rid] = self._assemble(int(sb.size), per_locus, limit, 5)
        return hits

    def predict(self, sequence_input, step: int = 1, limit: bool = False) -> MlstResult:
        """MlstResult for one record, a record iterator or a fasta/fastq path (reference :305-360)."""
        if seqio.is_record(sequence_input):
            if sequence_input.id == "<unknown id>":
                sequence_input.id = "test"
            hits = {sequence_input.id: self.calculate_hits(sequence_input.seq, step, limit)}
            return MlstResult(self.model_display_name, step, hits, None)
        if isinstance(sequence_input, Path):
            return ProbabilisticFilterMlstSchemeModel.predict(self, get_record_iterator(sequence_input), step=step, limit=limit)
        if seqio.is_record_iterator(sequence_input):
            return MlstResult(self.model_display_name, step, self._predict_records(sequence_input, step, limit), None)
        raise ValueError(
            "Invalid sequence input, must be a Seq object, a list of Seq objects, a"
            " SeqIO FastaIterator, or a SeqIO FastqPhredIterator"
        )

    def has_sufficient_score(self, highest_results: dict, locus_size: list[int]) -> bool:
        """True when any locus' best score reaches half its average allele length (reference :428-449)."""
        for i, allele_score in enumerate(highest_results.values()):
            if allele_score and next(iter(allele_score.values())) >= 0.5 * locus_size[i]:
                return True
        return False
