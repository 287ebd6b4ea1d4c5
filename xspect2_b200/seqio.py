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
    """All records of an input as one byte buffer + offsets: ``bases[begin[i]:end[i]]`` is record i.
    Record ids are decoded lazily (a 10 M-read FASTQ should not pay for 10 M Python strings unless asked)."""

    __slots__ = ("_ids", "_id_buf", "_id_end", "bases", "begin", "end", "records")

    def __init__(self, ids, bases: np.ndarray, begin: np.ndarray, end: np.ndarray, records=None, id_buf=None, id_end=None):
        self._ids = ids
        self._id_buf = id_buf
        self._id_end = id_end
        self.bases = bases
        self.begin = begin
        self.end = end
        self.records = records

    @property
    def ids(self) -> list[str]:
        if self._ids is None:
            text = self._id_buf.tobytes().decode("ascii", "replace")
            ends = self._id_end.tolist()
            self._ids = [text[a:b] for a, b in zip([0] + ends[:-1], ends)]
        return self._ids

    def __len__(self) -> int:
        return int(self.begin.size)

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
        """Native FASTA / FASTQ reader (xs_fastx_*; format by extension, like file_io.get_record_iterator).  The
        base buffer is page-locked when a GPU is present so the query's host-to-device copy is asynchronous."""
        import ctypes as C

        from . import _abi
        from .definitions import fasta_endings, fastq_endings

        path = Path(path)
        suffix = path.suffix[1:]
        if suffix in fastq_endings:
            fmt = 2
        elif suffix in fasta_endings:
            fmt = 1
        else:
            raise ValueError("Invalid file format, must be a fasta or fastq file")
        L = _abi.lib()
        h = C.c_void_p()
        _abi.check(L.xs_fastx_open(str(path).encode(), fmt, C.byref(h)))
        try:
            n_rec, n_bases, n_id = C.c_uint64(), C.c_uint64(), C.c_uint64()
            _abi.check(L.xs_fastx_stats(h, C.byref(n_rec), C.byref(n_bases), C.byref(n_id)))
            bases = _host_buffer(n_bases.value)
            begin = np.empty(n_rec.value, np.uint64)
            end = np.empty(n_rec.value, np.uint64)
            id_buf = np.empty(n_id.value, np.uint8)
            id_end = np.empty(n_rec.value, np.uint64)
            _abi.check(L.xs_fastx_read(h, bases.ctypes.data, begin.ctypes.data, end.ctypes.data, id_buf.ctypes.data, id_end.ctypes.data))
        finally:
            L.xs_fastx_close(h)
        return cls(None, bases, begin, end, None, id_buf, id_end)


def _host_buffer(n: int) -> np.ndarray:
    """Page-locked uint8 buffer when CUDA is available (large inputs only), ordinary memory otherwise."""
    if n >= (1 << 20):
        try:
            from .engine import pinned_empty

            return pinned_empty(n, np.uint8)
        except Exception:
            pass
    return np.empty(n, np.uint8)
