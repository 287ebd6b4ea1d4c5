"""FASTA / FASTQ input for the scoring path.

The reference feeds the path with Biopython (``SeqIO.parse`` in file_io.py:73-77; ``Seq``/``SeqRecord``
objects through predict, probabilistic_filter_model.py:237-331).  Biopython is not a dependency here: this
module provides small ``Seq`` / ``SeqRecord`` types and record iterators with Biopython's id semantics
(id = first whitespace-delimited word of the title line), and — the part the GPU path uses —
``SequenceBatch``: all records of an input laid out as one contiguous byte buffer plus offsets, which is what
the C ABI takes.  Biopython objects are accepted wherever records are (duck typing on ``.seq`` / ``.id``).
"""

from __future__ import annotations

import gzip
from pathlib import Path
from typing import Iterable, Iterator

import numpy as np

_COMPLEMENT = bytes.maketrans(b"ACGTMRWSYKVHDBXNUacgtmrwsykvhdbxnu", b"TGCAKYWSRMBDHVXNAtgcakywsrmbdhvxna")


class Seq:
    """A nucleotide sequence (immutable wrapper of ``str``) with the subset of Bio.Seq the path uses."""

    __slots__ = ("_data",)

    def __init__(self, data):
        if isinstance(data, Seq):
            data = data._data
        elif isinstance(data, (bytes, bytearray, memoryview)):
            data = bytes(data).decode("ascii")
        elif isinstance(data, np.ndarray):
            data = data.tobytes().decode("ascii")
        elif not isinstance(data, str):
            raise TypeError("data should be a string, bytes, bytearray, numpy uint8 array or Seq object")
        self._data = data

    def __str__(self) -> str:
        return self._data

    def __repr__(self) -> str:
        d = self._data
        return f"Seq({d!r})" if len(d) <= 60 else f"Seq('{d[:54]}...{d[-3:]}')"

    def __len__(self) -> int:
        return len(self._data)

    def __getitem__(self, item):
        r = self._data[item]
        return Seq(r) if isinstance(item, slice) else r

    def __iter__(self):
        return iter(self._data)

    def __hash__(self):
        return hash(self._data)

    def _cmp_key(self, other):
        if isinstance(other, Seq):
            return other._data
        if isinstance(other, str):
            return other
        if isinstance(other, (bytes, bytearray)):
            return bytes(other).decode("ascii")
        return NotImplemented

    def __eq__(self, other):
        o = self._cmp_key(other)
        return NotImplemented if o is NotImplemented else self._data == o

    def __lt__(self, other):
        o = self._cmp_key(other)
        return NotImplemented if o is NotImplemented else self._data < o

    def __le__(self, other):
        o = self._cmp_key(other)
        return NotImplemented if o is NotImplemented else self._data <= o

    def __gt__(self, other):
        o = self._cmp_key(other)
        return NotImplemented if o is NotImplemented else self._data > o

    def __ge__(self, other):
        o = self._cmp_key(other)
        return NotImplemented if o is NotImplemented else self._data >= o

    def complement(self) -> "Seq":
        return Seq(self._data.encode("ascii").translate(_COMPLEMENT))

    def reverse_complement(self) -> "Seq":
        return Seq(self._data.encode("ascii").translate(_COMPLEMENT)[::-1])

    def upper(self) -> "Seq":
        return Seq(self._data.upper())


class SeqRecord:
    """id / name / description / seq, like Bio.SeqRecord."""

    __slots__ = ("seq", "id", "name", "description")

    def __init__(self, seq, id: str = "<unknown id>", name: str = "<unknown name>", description: str = "<unknown description>"):
        self.seq = seq if (seq is None or hasattr(seq, "reverse_complement")) else Seq(seq)
        self.id = id
        self.name = name
        self.description = description

    def __len__(self) -> int:
        return len(self.seq)

    def __repr__(self) -> str:
        return f"SeqRecord(seq={self.seq!r}, id={self.id!r})"


def is_seq(obj) -> bool:
    """A Seq of this module or a Biopython Seq (anything str()-able with reverse_complement that is not a record)."""
    return isinstance(obj, Seq) or (hasattr(obj, "reverse_complement") and not hasattr(obj, "seq") and not isinstance(obj, str))


def is_record(obj) -> bool:
    return isinstance(obj, SeqRecord) or (hasattr(obj, "seq") and hasattr(obj, "id") and not isinstance(obj, (str, bytes)))


def _open_text(path: Path):
    if str(path).endswith(".gz"):
        return gzip.open(path, "rt", encoding="ascii", errors="replace")
    return open(path, "r", encoding="ascii", errors="replace")


class FastaIterator:
    """Iterates SeqRecords of a FASTA file (multi-line sequences, blanks and '\\r' stripped)."""

    def __init__(self, source):
        self._path = Path(source)
        self._it = self._records()

    def _records(self) -> Iterator[SeqRecord]:
        with _open_text(self._path) as f:
            title, parts = None, []
            for line in f:
                if line.startswith(">"):
                    if title is not None:
                        yield self._make(title, parts)
                    title, parts = line[1:].rstrip("\r\n"), []
                elif title is not None:
                    parts.append(line)
            if title is not None:
                yield self._make(title, parts)

    @staticmethod
    def _make(title: str, parts: list[str]) -> SeqRecord:
        seq = "".join(parts).replace("\n", "").replace("\r", "").replace(" ", "")
        words = title.split(None, 1)
        first = words[0] if words else ""
        return SeqRecord(Seq(seq), id=first, name=first, description=title)

    def __iter__(self):
        return self

    def __next__(self) -> SeqRecord:
        return next(self._it)


class FastqPhredIterator:
    """Iterates SeqRecords of a FASTQ file (wrapped sequence / quality lines allowed, qualities dropped)."""

    def __init__(self, source):
        self._path = Path(source)
        self._it = self._records()

    def _records(self) -> Iterator[SeqRecord]:
        with _open_text(self._path) as f:
            line = f.readline()
            while line:
                if not line.strip():
                    line = f.readline()
                    continue
                if not line.startswith("@"):
                    raise ValueError("Records in Fastq files should start with '@' character")
                title = line[1:].rstrip("\r\n")
                seq_parts = []
                line = f.readline()
                while line and not line.startswith("+"):
                    seq_parts.append(line.strip())
                    line = f.readline()
                if not line:
                    raise ValueError("End of file without quality information.")
                seq = "".join(seq_parts).replace(" ", "")
                qlen = 0
                line = f.readline()
                while line and (qlen < len(seq)):
                    qlen += len(line.strip())
                    line = f.readline()
                if qlen != len(seq):
                    raise ValueError(f"Lengths of sequence and quality values differs for {title} ({len(seq)} and {qlen}).")
                words = title.split(None, 1)
                first = words[0] if words else ""
                yield SeqRecord(Seq(seq), id=first, name=first, description=title)

    def __iter__(self):
        return self

    def __next__(self) -> SeqRecord:
        return next(self._it)


def parse(path, fmt: str):
    """``SeqIO.parse(path, "fasta" | "fastq")``."""
    if fmt == "fasta":
        return FastaIterator(path)
    if fmt == "fastq":
        return FastqPhredIterator(path)
    raise ValueError(f"Unknown format '{fmt}'")


def is_record_iterator(obj) -> bool:
    if isinstance(obj, (FastaIterator, FastqPhredIterator)):
        return True
    mod = type(obj).__module__ or ""
    return mod.startswith("Bio.SeqIO") and hasattr(obj, "__next__")


def write_fasta(records: Iterable, handle, wrap: int = 60) -> int:
    """``SeqIO.write(records, handle, "fasta")``: '>description' (id first) and 60-column lines."""
    n = 0
    for r in ([records] if is_record(records) else records):
        rid, desc = str(r.id), str(getattr(r, "description", "") or "")
        if desc and desc.split(None, 1)[0:1] == [rid]:
            title = desc
        elif desc and desc != "<unknown description>":
            title = f"{rid} {desc}"
        else:
            title = rid
        handle.write(f">{title}\n")
        s = str(r.seq)
        for i in range(0, len(s), wrap):
            handle.write(s[i : i + wrap] + "\n")
        n += 1
    return n


# --------------------------------------------------------------------------------------
# contiguous batches for the C ABI
# --------------------------------------------------------------------------------------
class SequenceBatch:
    """All records of an input as one byte buffer + offsets: ``bases[begin[i]:end[i]]`` is record i."""

    __slots__ = ("ids", "bases", "begin", "end", "records")

    def __init__(self, ids: list[str], bases: np.ndarray, begin: np.ndarray, end: np.ndarray, records=None):
        self.ids = ids
        self.bases = bases
        self.begin = begin
        self.end = end
        self.records = records

    def __len__(self) -> int:
        return len(self.ids)

    @property
    def lengths(self) -> np.ndarray:
        return (self.end - self.begin).astype(np.int64)

    def sequence(self, i: int) -> str:
        return self.bases[int(self.begin[i]) : int(self.end[i])].tobytes().decode("ascii")

    @classmethod
    def from_records(cls, records: Iterable, keep_records: bool = False) -> "SequenceBatch":
        ids, parts, lens, kept = [], [], [], []
        for r in records:
            s = str(r.seq).encode("ascii", "replace")
            ids.append(r.id)
            parts.append(s)
            lens.append(len(s))
            if keep_records:
                kept.append(r)
        bases = np.frombuffer(b"".join(parts), dtype=np.uint8) if parts else np.zeros(0, np.uint8)
        lens = np.asarray(lens, dtype=np.uint64)
        end = np.cumsum(lens, dtype=np.uint64)
        return cls(ids, bases, end - lens, end, kept if keep_records else None)

    @classmethod
    def from_file(cls, path: Path) -> "SequenceBatch":
        """Vectorised FASTA / FASTQ reader (format by extension, like file_io.get_record_iterator)."""
        from .definitions import fasta_endings, fastq_endings

        path = Path(path)
        suffix = path.suffix[1:]
        raw = np.fromfile(path, dtype=np.uint8)
        if suffix in fastq_endings:
            got = _fastq_fast(raw)
            if got is not None:
                return got
            return cls.from_records(FastqPhredIterator(path))
        if suffix in fasta_endings:
            return _fasta_fast(raw)
        raise ValueError("Invalid file format, must be a fasta or fastq file")


def _first_words(raw: np.ndarray, starts: np.ndarray, ends: np.ndarray) -> list[str]:
    """First whitespace-delimited word of each title line raw[starts[i]:ends[i]] (marker already skipped)."""
    buf = raw.tobytes()
    out = []
    for s, e in zip(starts.tolist(), ends.tolist()):
        w = buf[s:e].split(None, 1)
        out.append(w[0].decode("ascii", "replace") if w else "")
    return out


def _line_bounds(raw: np.ndarray):
    nl = np.flatnonzero(raw == 10)
    starts = np.concatenate(([0], nl + 1))
    ends = np.concatenate((nl, [raw.size]))
    if starts[-1] >= raw.size:          # file ends with '\n'
        starts, ends = starts[:-1], ends[:-1]
    # strip '\r'
    has_cr = (ends > starts) & (raw[np.maximum(ends - 1, 0)] == 13)
    ends = ends - has_cr.astype(ends.dtype)
    return starts.astype(np.int64), ends.astype(np.int64)


def _fastq_fast(raw: np.ndarray):
    """Strict 4-line FASTQ: sequences are used in place (offsets into the file buffer).  Returns None when the
    file is not in that shape (wrapped records, blank lines) so that the general parser takes over."""
    if raw.size == 0:
        return SequenceBatch([], raw, np.zeros(0, np.uint64), np.zeros(0, np.uint64))
    starts, ends = _line_bounds(raw)
    # drop trailing blank lines
    n = starts.size
    while n and ends[n - 1] == starts[n - 1]:
        n -= 1
    starts, ends = starts[:n], ends[:n]
    if n == 0 or n % 4:
        return None
    t, s, p, q = (slice(i, None, 4) for i in range(4))
    if not (np.all(raw[starts[t]] == ord("@")) and np.all(ends[p] > starts[p]) and np.all(raw[starts[p]] == ord("+"))):
        return None
    if not np.array_equal(ends[s] - starts[s], ends[q] - starts[q]):
        return None
    seq_b, seq_e = starts[s], ends[s]
    # embedded blanks inside a sequence line are rare; the general parser handles them
    ids = _first_words(raw, starts[t] + 1, ends[t])
    return SequenceBatch(ids, raw, seq_b.astype(np.uint64), seq_e.astype(np.uint64))


def _fasta_fast(raw: np.ndarray) -> SequenceBatch:
    """FASTA with wrapped lines: sequence bytes are compacted into one buffer (newlines, '\\r', blanks removed)."""
    if raw.size == 0:
        return SequenceBatch([], raw, np.zeros(0, np.uint64), np.zeros(0, np.uint64))
    starts, ends = _line_bounds(raw)
    is_title = (ends > starts) & (raw[np.minimum(starts, raw.size - 1)] == ord(">"))
    title_idx = np.flatnonzero(is_title)
    if title_idx.size == 0:
        return SequenceBatch([], np.zeros(0, np.uint8), np.zeros(0, np.uint64), np.zeros(0, np.uint64))
    # bytes that belong to sequences: not in a title line, after the first title, not whitespace
    keep = np.ones(raw.size, dtype=bool)
    keep[: starts[title_idx[0]]] = False
    title_mark = np.zeros(raw.size + 1, dtype=np.int32)
    np.add.at(title_mark, starts[title_idx], 1)
    np.add.at(title_mark, np.minimum(ends[title_idx] + 1, raw.size), -1)
    keep &= np.cumsum(title_mark[:-1]) == 0
    keep &= (raw != 10) & (raw != 13) & (raw != 32)
    bases = raw[keep]
    kept_before = np.concatenate(([0], np.cumsum(keep, dtype=np.int64)))
    rec_start = kept_before[starts[title_idx]]
    rec_end = np.concatenate((rec_start[1:], [bases.size]))
    ids = _first_words(raw, starts[title_idx] + 1, ends[title_idx])
    return SequenceBatch(ids, np.ascontiguousarray(bases), rec_start.astype(np.uint64), rec_end.astype(np.uint64))
