"""CPU oracle for the XspecT k-mer scoring hot path (python side).

TEST INFRASTRUCTURE ONLY — see the header of ``xs_oracle.cpp``.  Nothing under
``xspect2_b200/`` may import this module.  Allowed users: ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs.

PARITY STATUS: "parity unpinned" for the cobs-reloaded / rbloom pieces (file layouts, hash
seeds, non-ACGT handling, result order, LCG constants — SURVEY.md Appendix A, [UNVERIFIED-3P]);
pinned for XXH64 / XXH3-64 (python-xxhash known answers) and for every in-tree semantic
(reference file:line cited at each function): the restated loops at the end of this module
(``reference_predict*``, ``mlst_locus_scores``, ``sequence_splitter``) are checked against the
outputs of the reference's own Python code run over this module's ``CobsOracle`` / ``BloomOracle``
(tests/golden/make_reference_flows.py -> tests/golden/reference_flows.json,
tests/test_reference_golden.py).

Contents
  * readers for ``index.cobs_classic`` / ``<locus>.cobs_compact`` / ``filter.bloom`` (A.1, A.3, A.4)
  * format-faithful writers so fixtures can be generated offline (the reference's own tests
    train real models from NCBI downloads, tests/conftest.py:12-48 — impossible without network)
  * ``CobsOracle`` / ``BloomOracle``: the ``cobs_index.Search`` / ``rbloom.Bloom`` shaped objects
  * the python-level MLST epilogue restated from probabilistic_filter_mlst_model.py:236-303
"""

from __future__ import annotations

import ctypes as C
import math
import os
import struct
import subprocess
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent
_LIB = None

POLICY_SKIP = 0
POLICY_LITERAL = 1


class _CobsT(C.Structure):
    _fields_ = [
        ("data", C.c_void_p),
        ("n_pages", C.c_uint32),
        ("page_bytes", C.c_uint64),
        ("sig_size", C.c_void_p),
        ("page_off", C.c_void_p),
        ("n_docs", C.c_uint32),
        ("num_hashes", C.c_uint32),
        ("k", C.c_uint32),
        ("canonicalize", C.c_int),
        ("policy", C.c_int),
    ]


class _BloomT(C.Structure):
    _fields_ = [
        ("bits", C.c_void_p),
        ("n_bits", C.c_uint64),
        ("k_hashes", C.c_uint64),
        ("k", C.c_uint32),
    ]


def build(force: bool = False) -> Path:
    """Compile ``libxs_oracle.so`` with g++ (oracle/Makefile) if it is missing or stale."""
    so = _HERE / "libxs_oracle.so"
    src = _HERE / "xs_oracle.cpp"
    if force or not so.exists() or so.stat().st_mtime < src.stat().st_mtime:
        subprocess.run(["make", "-C", str(_HERE), "-B", "libxs_oracle.so"], check=True,
                       stdout=subprocess.PIPE, stderr=subprocess.STDOUT)
    return so


def lib() -> C.CDLL:
    global _LIB
    if _LIB is None:
        L = C.CDLL(str(build()))
        L.xso_xxh64.restype = C.c_uint64
        L.xso_xxh64.argtypes = [C.c_char_p, C.c_uint64, C.c_uint64]
        L.xso_xxh3_64.restype = C.c_uint64
        L.xso_xxh3_64.argtypes = [C.c_char_p, C.c_uint64]
        L.xso_cobs_term.restype = C.c_int
        L.xso_cobs_term.argtypes = [C.c_char_p, C.c_uint32, C.c_int, C.c_int, C.c_char_p]
        L.xso_bloom_term.restype = None
        L.xso_bloom_term.argtypes = [C.c_char_p, C.c_uint32, C.c_char_p]
        L.xso_bio_complement_table.argtypes = [C.c_void_p]
        L.xso_cobs_query_one.argtypes = [C.POINTER(_CobsT), C.c_void_p, C.c_uint64, C.c_uint32, C.c_void_p]
        L.xso_cobs_query_batch.argtypes = [C.POINTER(_CobsT), C.c_void_p, C.c_void_p, C.c_void_p,
                                           C.c_uint64, C.c_uint32, C.c_void_p, C.c_int]
        L.xso_cobs_rows.argtypes = [C.POINTER(_CobsT), C.c_void_p, C.c_uint64, C.c_uint32, C.c_void_p, C.c_void_p]
        L.xso_cobs_insert.argtypes = [C.POINTER(_CobsT), C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint64]
        L.xso_cobs_result_order.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p]
        L.xso_bloom_contains_hash.restype = C.c_int
        L.xso_bloom_contains_hash.argtypes = [C.POINTER(_BloomT), C.c_uint64]
        L.xso_bloom_indexes.argtypes = [C.POINTER(_BloomT), C.c_uint64, C.c_void_p]
        L.xso_bloom_hits_one.restype = C.c_uint32
        L.xso_bloom_hits_one.argtypes = [C.POINTER(_BloomT), C.c_void_p, C.c_uint64, C.c_uint32]
        L.xso_bloom_hits_batch.argtypes = [C.POINTER(_BloomT), C.c_void_p, C.c_void_p, C.c_void_p,
                                           C.c_uint64, C.c_uint32, C.c_void_p, C.c_int]
        L.xso_bloom_insert.argtypes = [C.POINTER(_BloomT), C.c_void_p, C.c_void_p, C.c_uint64]
        L.xso_max_threads.restype = C.c_int
        L.xso_synth_row.argtypes = [C.c_uint64, C.c_uint64, C.c_uint32, C.c_void_p]
        L.xso_cobs_query_batch_synth.argtypes = [C.c_uint64, C.c_uint32, C.c_uint64, C.c_uint32, C.c_uint32, C.c_int, C.c_int,
                                                 C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint32, C.c_void_p, C.c_int]
        _LIB = L
    return _LIB


# --------------------------------------------------------------------------------------
# hashes and canonical forms
# --------------------------------------------------------------------------------------
def xxh64(data: bytes, seed: int = 0) -> int:
    return lib().xso_xxh64(data, len(data), seed)


def xxh3_64(data: bytes) -> int:
    if len(data) > 128:
        raise ValueError("oracle XXH3 covers lengths 0..128 (k-mers)")
    return lib().xso_xxh3_64(data, len(data))


def cobs_term(window: bytes, canonicalize: int = 1, policy: int = POLICY_SKIP) -> bytes | None:
    out = C.create_string_buffer(len(window))
    ok = lib().xso_cobs_term(window, len(window), canonicalize, policy, out)
    return out.raw if ok else None


def bloom_term(window: bytes) -> bytes:
    """min(kmer, revcomp) exactly as probabilistic_single_filter_model.py:175-180."""
    out = C.create_string_buffer(len(window))
    lib().xso_bloom_term(window, len(window), out)
    return out.raw


def bio_complement_table() -> bytes:
    buf = (C.c_uint8 * 256)()
    lib().xso_bio_complement_table(buf)
    return bytes(buf)


def count_kmers(length: int, k: int, step: int = 1) -> int:
    """probabilistic_filter_model.py:462 — ceil((len - k + 1) / step), counts N windows too."""
    return math.ceil((length - k + 1) / step)


# --------------------------------------------------------------------------------------
# file formats — readers
# --------------------------------------------------------------------------------------
def _read_names(buf: memoryview, pos: int, n: int) -> tuple[list[str], int]:
    names = []
    for _ in range(n):
        end = pos
        while buf[end] != 0x0A:
            end += 1
        names.append(bytes(buf[pos:end]).decode("utf-8"))
        pos = end + 1
    return names, pos


def parse_cobs(path: str | os.PathLike) -> dict:
    """Parse a COBS classic or compact index file header; returns geometry + a uint8 memmap."""
    mm = np.memmap(path, dtype=np.uint8, mode="r")
    buf = memoryview(mm)
    if bytes(buf[:5]) != b"COBS:":
        raise ValueError("not a COBS file (missing 'COBS:' magic)")
    if bytes(buf[5:18]) == b"CLASSIC_INDEX":
        pos = 18
        version, term_size = struct.unpack_from("<II", buf, pos); pos += 8
        canonicalize = buf[pos]; pos += 1
        (n_docs,) = struct.unpack_from("<I", buf, pos); pos += 4
        sig, nh = struct.unpack_from("<QQ", buf, pos); pos += 16
        names, pos = _read_names(buf, pos, n_docs)
        if bytes(buf[pos:pos + 13]) != b"CLASSIC_INDEX":
            raise ValueError("classic header end magic missing")
        pos += 13
        row = (n_docs + 7) // 8
        if mm.size - pos != sig * row:
            raise ValueError(f"classic size identity fails: {mm.size - pos} != {sig}*{row}")
        return dict(kind="classic", version=version, k=term_size, canonicalize=int(canonicalize),
                    n_docs=n_docs, num_hashes=int(nh), page_bytes=row, sig_sizes=[int(sig)],
                    page_offs=[0], data_off=pos, names=names, mm=mm)
    if bytes(buf[5:18]) == b"COMPACT_INDEX":
        pos = 18
        version, term_size = struct.unpack_from("<II", buf, pos); pos += 8
        canonicalize = buf[pos]; pos += 1
        n_pages, n_docs = struct.unpack_from("<II", buf, pos); pos += 8
        (page_size,) = struct.unpack_from("<Q", buf, pos); pos += 8
        sigs, nhs = [], []
        for _ in range(n_pages):
            s, h = struct.unpack_from("<QQ", buf, pos); pos += 16
            sigs.append(int(s)); nhs.append(int(h))
        if len(set(nhs)) != 1:
            raise ValueError("compact pages with differing num_hashes are not supported")
        names, pos = _read_names(buf, pos, n_docs)
        pad = (page_size - ((pos + 13) % page_size)) % page_size
        # remainder 0: cobs may write no padding or one whole page (not checkable offline); accept either
        for cand in ([pad] if pad else [0, page_size]):
            if bytes(buf[pos + cand:pos + cand + 13]) == b"COMPACT_INDEX":
                pos += cand
                break
        else:
            raise ValueError("compact header end magic missing")
        pos += 13
        offs, o = [], 0
        for s in sigs:
            offs.append(o); o += s * page_size
        if mm.size - pos != o:
            raise ValueError(f"compact size identity fails: {mm.size - pos} != {o}")
        return dict(kind="compact", version=version, k=term_size, canonicalize=int(canonicalize),
                    n_docs=n_docs, num_hashes=nhs[0], page_bytes=int(page_size), sig_sizes=sigs,
                    page_offs=offs, data_off=pos, names=names, mm=mm)
    raise ValueError("unknown COBS magic word")


def parse_bloom(path: str | os.PathLike) -> dict:
    mm = np.memmap(path, dtype=np.uint8, mode="r")
    if mm.size < 9:
        raise ValueError("bloom file too small")
    (k_hashes,) = struct.unpack_from("<Q", memoryview(mm), 0)
    return dict(k_hashes=int(k_hashes), n_bits=(mm.size - 8) * 8, mm=mm)


# --------------------------------------------------------------------------------------
# writers (fixture construction; mirrors cobs classic/compact construction and rbloom.save)
# --------------------------------------------------------------------------------------
def cobs_signature_size(max_doc_kmers: int, num_hashes: int, fpr: float) -> int:
    """A.1.4 — ceil(n * (-h / ln(1 - fpr^(1/h))))."""
    ratio = -num_hashes / math.log(1.0 - fpr ** (1.0 / num_hashes))
    return max(1, math.ceil(max_doc_kmers * ratio))


def classic_header(k: int, canonicalize: int, names: list[str], sig: int, num_hashes: int) -> bytes:
    h = b"COBS:" + b"CLASSIC_INDEX" + struct.pack("<II", 1, k) + struct.pack("<B", canonicalize)
    h += struct.pack("<I", len(names)) + struct.pack("<QQ", sig, num_hashes)
    for n in names:
        h += n.encode("utf-8") + b"\n"
    return h + b"CLASSIC_INDEX"


def compact_header(k: int, canonicalize: int, names: list[str], page_size: int,
                   sigs: list[int], num_hashes: int, pad_full_page: bool = False) -> bytes:
    h = b"COBS:" + b"COMPACT_INDEX" + struct.pack("<II", 1, k) + struct.pack("<B", canonicalize)
    h += struct.pack("<II", len(sigs), len(names)) + struct.pack("<Q", page_size)
    for s in sigs:
        h += struct.pack("<QQ", s, num_hashes)
    for n in names:
        h += n.encode("utf-8") + b"\n"
    pad = (page_size - ((len(h) + 13) % page_size)) % page_size
    if pad == 0 and pad_full_page:      # the other reading of cobs' padding rule (range 1..page_size)
        pad = page_size
    return h + b"\0" * pad + b"COMPACT_INDEX"


def _as_u8(seq) -> np.ndarray:
    if isinstance(seq, str):
        seq = seq.encode("ascii")
    if isinstance(seq, (bytes, bytearray)):
        return np.frombuffer(bytes(seq), dtype=np.uint8)
    return np.ascontiguousarray(seq, dtype=np.uint8)


def write_classic(path, docs: dict[str, list], k: int, num_hashes: int = 7, fpr: float = 0.01,
                  canonicalize: int = 1, sig_size: int | None = None, policy: int = POLICY_SKIP) -> dict:
    """Build ``index.cobs_classic`` from ``docs = {name: [sequence, ...]}`` (dict order = doc order)."""
    names = list(docs)
    n_docs = len(names)
    if sig_size is None:
        mx = max(sum(max(0, len(s) - k + 1) for s in seqs) for seqs in docs.values())
        sig_size = cobs_signature_size(max(mx, 1), num_hashes, fpr)
    row = (n_docs + 7) // 8
    data = np.zeros(sig_size * row, dtype=np.uint8)
    sig = np.array([sig_size], dtype=np.uint64)
    off = np.array([0], dtype=np.uint64)
    ct = _CobsT(data.ctypes.data, 1, row, sig.ctypes.data, off.ctypes.data, n_docs, num_hashes, k,
                canonicalize, policy)
    for d, name in enumerate(names):
        for s in docs[name]:
            a = _as_u8(s)
            lib().xso_cobs_insert(C.byref(ct), data.ctypes.data, d, a.ctypes.data, a.size)
    with open(path, "wb") as f:
        f.write(classic_header(k, canonicalize, names, sig_size, num_hashes))
        f.write(data.tobytes())
    return dict(sig_size=sig_size, row=row)


def write_compact(path, docs: dict[str, list], k: int, num_hashes: int = 1, fpr: float = 0.001,
                  canonicalize: int = 1, page_size: int | None = None, policy: int = POLICY_SKIP) -> dict:
    """Build ``<locus>.cobs_compact``: documents sorted by size, 8*page_size per page (A.3)."""
    sizes = {n: sum(max(0, len(s) - k + 1) for s in seqs) for n, seqs in docs.items()}
    names = sorted(docs, key=lambda n: sizes[n])  # stable: ties keep given order
    n_docs = len(names)
    if page_size is None:
        page_size = max(1, int(math.sqrt(n_docs / 8)))
    per_page = 8 * page_size
    n_pages = (n_docs + per_page - 1) // per_page
    sigs = []
    for p in range(n_pages):
        grp = names[p * per_page:(p + 1) * per_page]
        sigs.append(cobs_signature_size(max(max(sizes[n] for n in grp), 1), num_hashes, fpr))
    offs, o = [], 0
    for s in sigs:
        offs.append(o); o += s * page_size
    data = np.zeros(o, dtype=np.uint8)
    sig = np.array(sigs, dtype=np.uint64)
    off = np.array(offs, dtype=np.uint64)
    ct = _CobsT(data.ctypes.data, n_pages, page_size, sig.ctypes.data, off.ctypes.data, n_docs,
                num_hashes, k, canonicalize, policy)
    for d, name in enumerate(names):
        for s in docs[name]:
            a = _as_u8(s)
            lib().xso_cobs_insert(C.byref(ct), data.ctypes.data, d, a.ctypes.data, a.size)
    with open(path, "wb") as f:
        f.write(compact_header(k, canonicalize, names, page_size, sigs, num_hashes))
        f.write(data.tobytes())
    return dict(page_size=page_size, sigs=sigs, names=names)


def bloom_params(expected_items: int, fpr: float) -> tuple[int, int]:
    """rbloom.Bloom.__new__ (A.4.1): bits = -n ln(fpr)/ln(2)^2 truncated; k = trunc(bits/n ln 2)."""
    size_in_bits = -1.0 * expected_items * math.log(fpr) / (math.log(2.0) ** 2)
    k = int((size_in_bits / expected_items) * math.log(2.0))
    n_bytes = (int(size_in_bits) + 7) // 8
    return n_bytes, k


def write_bloom(path, seqs: list, k: int, fpr: float = 0.01, n_bytes: int | None = None,
                k_hashes: int | None = None) -> dict:
    """Build ``filter.bloom`` like ProbabilisticSingleFilterModel.fit (:83-96) + rbloom.save."""
    total = sum(len(s) for s in seqs)
    nb, kh = bloom_params(max(total - k + 1, 1), fpr)
    n_bytes = nb if n_bytes is None else n_bytes
    k_hashes = kh if k_hashes is None else k_hashes
    bits = np.zeros(n_bytes, dtype=np.uint8)
    bt = _BloomT(bits.ctypes.data, n_bytes * 8, k_hashes, k)
    for s in seqs:
        a = _as_u8(s)
        lib().xso_bloom_insert(C.byref(bt), bits.ctypes.data, a.ctypes.data, a.size)
    with open(path, "wb") as f:
        f.write(struct.pack("<Q", k_hashes))
        f.write(bits.tobytes())
    return dict(n_bytes=n_bytes, k_hashes=k_hashes)


# --------------------------------------------------------------------------------------
# cobs_index.Search / rbloom.Bloom shaped oracles
# --------------------------------------------------------------------------------------
class SearchResult:
    """Shape of cobs_index SearchResult (doc_name, score) used at probabilistic_filter_model.py:406-409."""

    __slots__ = ("doc_name", "score")

    def __init__(self, doc_name: str, score: int):
        self.doc_name = doc_name
        self.score = score

    def __repr__(self):
        return f"SearchResult({self.doc_name!r}, {self.score})"


class CobsOracle:
    """``cobs_index.Search(path, load_complete)`` restated on the CPU."""

    def __init__(self, path, load_complete: bool = True, policy: int = POLICY_SKIP):
        self.meta = parse_cobs(path)
        m = self.meta
        self.names = m["names"]
        self.n_docs = m["n_docs"]
        self.k = m["k"]
        self.num_hashes = m["num_hashes"]
        data = m["mm"][m["data_off"]:]
        self._data = np.ascontiguousarray(data) if load_complete else data
        self._sig = np.array(m["sig_sizes"], dtype=np.uint64)
        self._off = np.array(m["page_offs"], dtype=np.uint64)
        self._ct = _CobsT(self._data.ctypes.data, len(m["sig_sizes"]), m["page_bytes"],
                          self._sig.ctypes.data, self._off.ctypes.data, self.n_docs,
                          self.num_hashes, self.k, m["canonicalize"], policy)

    def counts(self, seq, step: int = 1) -> np.ndarray:
        a = _as_u8(seq)
        out = np.zeros(self.n_docs, dtype=np.uint32)
        lib().xso_cobs_query_one(C.byref(self._ct), a.ctypes.data, a.size, step, out.ctypes.data)
        return out

    def counts_batch(self, bases: np.ndarray, seq_begin: np.ndarray, seq_end: np.ndarray,
                     step: int = 1, threads: int = 1) -> np.ndarray:
        bases = np.ascontiguousarray(bases, dtype=np.uint8)
        b = np.ascontiguousarray(seq_begin, dtype=np.uint64)
        e = np.ascontiguousarray(seq_end, dtype=np.uint64)
        out = np.zeros((b.size, self.n_docs), dtype=np.uint32)
        lib().xso_cobs_query_batch(C.byref(self._ct), bases.ctypes.data, b.ctypes.data, e.ctypes.data,
                                   b.size, step, out.ctypes.data, threads)
        return out

    def rows(self, seq, step: int = 1) -> tuple[np.ndarray, np.ndarray]:
        a = _as_u8(seq)
        n_w = max(0, (a.size - self.k) // step + 1) if a.size >= self.k else 0
        n_pages = len(self.meta["sig_sizes"])
        rows = np.zeros((n_w, self.num_hashes, n_pages), dtype=np.uint64)
        valid = np.zeros(n_w, dtype=np.uint8)
        if n_w:
            lib().xso_cobs_rows(C.byref(self._ct), a.ctypes.data, a.size, step, rows.ctypes.data, valid.ctypes.data)
        return rows, valid

    @staticmethod
    def result_order(scores: np.ndarray) -> np.ndarray:
        s = np.ascontiguousarray(scores, dtype=np.uint32)
        order = np.zeros(s.size, dtype=np.uint32)
        lib().xso_cobs_result_order(s.ctypes.data, s.size, order.ctypes.data)
        return order


    def search(self, query: str, step: int = 1) -> list[SearchResult]:
        """All documents, ordered like ClassicSearch::search (A.2.5c/d)."""
        if len(query) < self.k:
            raise RuntimeError("query too short for index term size")
        c = self.counts(query, step)
        return [SearchResult(self.names[i], int(c[i])) for i in self.result_order(c)]


class SynthCobsOracle:
    """The query of ``CobsOracle`` against the synthetic classic index of BASELINE config 5 (rows from the
    counter-based generator documented at ``xs_cobs_create_synthetic``; the index exists nowhere as a file)."""

    def __init__(self, n_docs: int, sig_size: int, k: int, num_hashes: int, seed: int, policy: int = POLICY_SKIP):
        self.n_docs, self.sig_size, self.k, self.num_hashes, self.seed, self.policy = n_docs, sig_size, k, num_hashes, seed, policy

    def row(self, r: int) -> np.ndarray:
        """Row ``r`` as bytes, little-endian 32-document words (the file layout of a classic index row)."""
        w = np.zeros((self.n_docs + 31) // 32, dtype=np.uint32)
        lib().xso_synth_row(self.seed, r, self.n_docs, w.ctypes.data)
        return w.view(np.uint8)[: (self.n_docs + 7) // 8].copy()

    def counts_batch(self, bases, seq_begin, seq_end, step: int = 1, threads: int = 1) -> np.ndarray:
        bases = np.ascontiguousarray(bases, dtype=np.uint8)
        b = np.ascontiguousarray(seq_begin, dtype=np.uint64)
        e = np.ascontiguousarray(seq_end, dtype=np.uint64)
        out = np.zeros((b.size, self.n_docs), dtype=np.uint32)
        lib().xso_cobs_query_batch_synth(self.seed, self.n_docs, self.sig_size, self.num_hashes, self.k, 1, self.policy,
                                         bases.ctypes.data, b.ctypes.data, e.ctypes.data, b.size, step, out.ctypes.data, threads)
        return out


class BloomOracle:
    """``rbloom.Bloom.load(path, hash_func=xxh3_64_intdigest)`` restated on the CPU."""

    def __init__(self, path, k: int):
        m = parse_bloom(path)
        self.k = k
        self.k_hashes = m["k_hashes"]
        self.n_bits = m["n_bits"]
        self._bits = np.ascontiguousarray(m["mm"][8:])
        self._bt = _BloomT(self._bits.ctypes.data, self.n_bits, self.k_hashes, k)

    def __contains__(self, kmer) -> bool:
        b = kmer.encode("utf-8") if isinstance(kmer, str) else bytes(kmer)
        return bool(lib().xso_bloom_contains_hash(C.byref(self._bt), xxh3_64(b)))

    def indexes(self, kmer: bytes) -> np.ndarray:
        out = np.zeros(self.k_hashes, dtype=np.uint64)
        lib().xso_bloom_indexes(C.byref(self._bt), xxh3_64(kmer), out.ctypes.data)
        return out

    def hits(self, seq, step: int = 1) -> int:
        a = _as_u8(seq)
        return int(lib().xso_bloom_hits_one(C.byref(self._bt), a.ctypes.data, a.size, step))

    def hits_batch(self, bases, seq_begin, seq_end, step: int = 1, threads: int = 1) -> np.ndarray:
        bases = np.ascontiguousarray(bases, dtype=np.uint8)
        b = np.ascontiguousarray(seq_begin, dtype=np.uint64)
        e = np.ascontiguousarray(seq_end, dtype=np.uint64)
        out = np.zeros(b.size, dtype=np.uint32)
        lib().xso_bloom_hits_batch(C.byref(self._bt), bases.ctypes.data, b.ctypes.data, e.ctypes.data,
                                   b.size, step, out.ctypes.data, threads)
        return out


def kmer_rows(seq, k: int, num_hashes: int, sig_size: int, canonicalize: int = 1, step: int = 1,
              policy: int = POLICY_SKIP) -> tuple[np.ndarray, np.ndarray]:
    """Row ids [n_windows, num_hashes] for an explicit classic geometry (no file needed)."""
    a = _as_u8(seq)
    n_w = (a.size - k) // step + 1 if a.size >= k else 0
    rows = np.zeros((n_w, num_hashes, 1), dtype=np.uint64)
    valid = np.zeros(n_w, dtype=np.uint8)
    sig = np.array([sig_size], dtype=np.uint64)
    off = np.array([0], dtype=np.uint64)
    ct = _CobsT(None, 1, 1, sig.ctypes.data, off.ctypes.data, 1, num_hashes, k, canonicalize, policy)
    if n_w:
        lib().xso_cobs_rows(C.byref(ct), a.ctypes.data, a.size, step, rows.ctypes.data, valid.ctypes.data)
    return rows[:, :, 0], valid


def max_threads() -> int:
    return lib().xso_max_threads()


# --------------------------------------------------------------------------------------
# MLST python-level epilogue (probabilistic_filter_mlst_model.py)
# --------------------------------------------------------------------------------------
def sequence_splitter(input_sequence: str, allele_len: int, k: int) -> list[str]:
    """probabilistic_filter_mlst_model.py:382-426, restated."""
    n = len(input_sequence)
    if n < 1000000:
        sub = allele_len
    elif n < 10000000:
        sub = allele_len * 10
    else:
        sub = allele_len * 100
    out, start = [], 0
    while start + sub <= n:
        out.append(input_sequence[start:start + sub])
        start += sub - k + 1
    if start < n:
        rest = input_sequence[start:]
        if len(rest) < k:
            out[-1] += rest
        else:
            out.append(rest)
    return out


def mlst_locus_scores(index: CobsOracle, sequence: str, allele_len: int, step: int = 1) -> dict[str, int]:
    """One locus of calculate_hits (:236-256 long branch, :272-286 short branch): allele → score,
    in the reference's dict order."""
    if len(sequence) >= 10000:
        all_counts: dict[str, int] = {}
        for chunk in sequence_splitter(sequence, allele_len, index.k):
            kept = {r.doc_name: r.score for r in index.search(chunk, step) if r.score > 50}
            for name, v in kept.items():
                all_counts[name] = all_counts.get(name, 0) + v
        return dict(sorted(all_counts.items(), key=lambda item: -item[1]))
    return {r.doc_name: r.score for r in index.search(sequence, step)}


# --------------------------------------------------------------------------------------
# the reference's per-record predict loops, restated over the oracle objects
# --------------------------------------------------------------------------------------
def reference_predict(index: CobsOracle, records: list[tuple[str, str]], k: int, exclude_ids=None, step: int = 1):
    """ProbabilisticFilterModel.predict's loop (probabilistic_filter_model.py:291-310): one search per record,
    dict in cobs result order, exclude filter, num_kmers = ceil((len-k+1)/step); later ids overwrite."""
    hits, num_kmers = {}, {}
    for rid, seq in records:
        if not len(seq) > k:
            raise ValueError("Invalid sequence, must be longer than k")
        d = {r.doc_name: r.score for r in index.search(seq, step)}
        if exclude_ids:
            d = {doc: s for doc, s in d.items() if doc not in exclude_ids}
        hits[rid] = d
        num_kmers[rid] = count_kmers(len(seq), k, step)
    return hits, num_kmers


def reference_predict_bloom(bf: BloomOracle, key: str, records: list[tuple[str, str]], k: int, step: int = 1):
    """ProbabilisticSingleFilterModel.calculate_hits (probabilistic_single_filter_model.py:98-125) per record."""
    hits, num_kmers = {}, {}
    for rid, seq in records:
        if not len(seq) > k:
            raise ValueError("Invalid sequence, must be longer than k")
        hits[rid] = {key: bf.hits(seq, step)}
        num_kmers[rid] = count_kmers(len(seq), k, step)
    return hits, num_kmers
