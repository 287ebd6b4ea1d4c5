"""Batched GPU engine behind the model classes: HBM-resident COBS indices and Bloom filters.

``CobsIndex`` / ``BloomFilter`` wrap the C-ABI handles; ``Search`` and ``Bloom`` present the
single-record shapes of the third-party objects the reference calls
(``cobs_index.Search.search`` at probabilistic_filter_model.py:227, ``kmer in rbloom.Bloom`` at
probabilistic_single_filter_model.py:122-124) on top of the same kernels.
"""

from __future__ import annotations

import ctypes as C
import os
import threading
from pathlib import Path

import numpy as np

from . import _abi
from ._abi import XS_U8, XS_U16, XS_U32, check, lib

_DT = {XS_U8: np.uint8, XS_U16: np.uint16, XS_U32: np.uint32}


# --------------------------------------------------------------------------------------
# pinned host memory
# --------------------------------------------------------------------------------------
_POOL: dict[int, list[int]] = {}      # capacity -> free page-locked pointers (cudaMallocHost costs ~0.5 ms per MB)
_POOL_BYTES = [0]
_POOL_LIMIT = int(os.environ.get("XSPECT_B200_PINNED_POOL_MB", 4096)) << 20
_POOL_LOCK = threading.Lock()


def trim_pinned_pool(keep_bytes: int = 0) -> int:
    """Release cached page-locked blocks until at most ``keep_bytes`` stay in the pool; returns the bytes freed."""
    freed = 0
    with _POOL_LOCK:
        for cap in sorted(_POOL, reverse=True):
            free = _POOL[cap]
            while free and _POOL_BYTES[0] > keep_bytes:
                lib().xs_host_free(free.pop())
                _POOL_BYTES[0] -= cap
                freed += cap
    return freed


def _capacity(nbytes: int) -> int:
    n = max(int(nbytes), 1)
    if n <= (1 << 20):
        return 1 << 20
    step = 1 << (n.bit_length() - 3)      # 8 size classes per power of two: at most 12.5 % slack
    return -(-n // step) * step


class _PinnedBlock:
    """Page-locked host allocation (xs_host_alloc) exposed through the array interface; recycled through a small
    pool when the array that owns it is collected."""

    def __init__(self, nbytes: int):
        self.cap = _capacity(nbytes)
        self.ptr = None
        with _POOL_LOCK:
            free = _POOL.get(self.cap)
            if free:
                self.ptr = free.pop()
                _POOL_BYTES[0] -= self.cap
        if self.ptr is None:
            p = C.c_void_p()
            check(lib().xs_host_alloc(self.cap, C.byref(p)))
            self.ptr = p.value
        self.nbytes = int(nbytes)
        self.__array_interface__ = {"shape": (max(self.nbytes, 1),), "typestr": "|u1", "data": (self.ptr, False), "version": 3}

    def __del__(self):
        ptr, self.ptr = getattr(self, "ptr", None), None
        if ptr:
            try:
                with _POOL_LOCK:
                    keep = _POOL_BYTES[0] + self.cap <= _POOL_LIMIT
                    if keep:
                        _POOL.setdefault(self.cap, []).append(ptr)
                        _POOL_BYTES[0] += self.cap
                if not keep:
                    lib().xs_host_free(ptr)
            except Exception:  # interpreter shutdown
                pass


def pinned_empty(shape, dtype=np.uint8) -> np.ndarray:
    """An uninitialised page-locked numpy array (fast, asynchronous host<->device copies)."""
    dtype = np.dtype(dtype)
    shape = (shape,) if np.isscalar(shape) else tuple(shape)
    n = int(np.prod(shape, dtype=np.int64)) if shape else 1
    block = _PinnedBlock(n * dtype.itemsize)
    flat = np.asarray(block)[: n * dtype.itemsize]
    return flat.view(dtype).reshape(shape)


def _ptr(a: np.ndarray) -> int:
    return a.ctypes.data


def _as_bases(bases) -> np.ndarray:
    if isinstance(bases, str):
        bases = bases.encode("utf-8")
    if isinstance(bases, (bytes, bytearray, memoryview)):
        return np.frombuffer(bases, dtype=np.uint8)
    a = np.asarray(bases)
    if a.dtype != np.uint8 or not a.flags.c_contiguous:
        a = np.ascontiguousarray(a, dtype=np.uint8)
    return a


def _as_u64(a) -> np.ndarray:
    a = np.asarray(a)
    if a.dtype != np.uint64 or not a.flags.c_contiguous:
        a = np.ascontiguousarray(a, dtype=np.uint64)
    return a


def device_trim(device: int = 0) -> None:
    """Return the library's cached device workspaces / bucketed scratch to the driver (``xs_device_trim``)."""
    check(lib().xs_device_trim(int(device)))


def pick_dtype(max_windows: int) -> int:
    """Smallest count type that cannot saturate for sequences of at most ``max_windows`` windows."""
    return XS_U8 if max_windows <= 255 else XS_U16 if max_windows <= 65535 else XS_U32


# --------------------------------------------------------------------------------------
# COBS
# --------------------------------------------------------------------------------------
class CobsIndex:
    """A COBS classic / compact index file resident in one GPU's HBM (optionally a column shard)."""

    def __init__(self, path, device: int = 0, doc_begin: int = 0, doc_end: int = 0, _handle=None):
        self._h = C.c_void_p()
        self.path = str(path)
        if _handle is not None:
            self._h = _handle
        else:
            if not os.path.isfile(self.path):
                raise FileNotFoundError(f"Index file not found at {self.path}")
            check(lib().xs_cobs_open(self.path.encode(), int(device), int(doc_begin), int(doc_end), C.byref(self._h)))
        info = _abi.CobsInfo()
        check(lib().xs_cobs_info(self._h, C.byref(info)))
        self.info = info
        need = C.c_uint64()
        check(lib().xs_cobs_doc_names(self._h, None, 0, C.byref(need)))
        buf = C.create_string_buffer(need.value + 1)
        check(lib().xs_cobs_doc_names(self._h, buf, need.value, C.byref(need)))
        names = buf.raw[: need.value].decode("utf-8").split("\n")
        self.all_names = names[:-1] if names and names[-1] == "" else names
        self.names = self.all_names[info.doc_begin : info.doc_end]

    @classmethod
    def synthetic(cls, n_docs: int, sig_size: int, k: int, num_hashes: int, seed: int, device: int = 0, doc_begin: int = 0,
                  doc_end: int = 0) -> "CobsIndex":
        """A classic index (or a document-column shard of it) whose rows come from the counter-based generator of
        ``xs_cobs_create_synthetic`` — BASELINE config 5's 120 GB index, generated straight into HBM."""
        h = C.c_void_p()
        check(lib().xs_cobs_create_synthetic(int(device), int(n_docs), int(doc_begin), int(doc_end), int(sig_size), int(k),
                                             int(num_hashes), int(seed), C.byref(h)))
        return cls(f"<synthetic D={n_docs} S={sig_size} seed={seed}>", _handle=h)

    # geometry
    k = property(lambda self: self.info.term_size)
    num_hashes = property(lambda self: self.info.num_hashes)
    n_docs = property(lambda self: self.info.doc_end - self.info.doc_begin)
    device = property(lambda self: self.info.device)

    @property
    def header_layout(self) -> str:
        """Which candidate reading of the file header matched (``xs_cobs_header_layout``)."""
        return lib().xs_cobs_header_layout(self._h).decode()

    @property
    def kernel(self) -> str:
        """The scoring kernel the handle's row layout selects for direct batches (``xs_cobs_kernel``)."""
        return lib().xs_cobs_kernel(self._h).decode()

    def doc_fill(self, sample_rows: int = 0) -> np.ndarray:
        """Fraction of set bits per local document over evenly spaced sample rows (``xs_cobs_doc_fill``)."""
        fill = np.zeros(self.n_docs, np.float64)
        check(lib().xs_cobs_doc_fill(self._h, int(sample_rows), _ptr(fill)))
        return fill

    def selfcheck(self, member_sequence=None, member_doc: str | int | None = None, step: int = 1) -> dict:
        """Structural self-check of a real model file (everything inside cobs-reloaded is restated
        from its published source and cannot be verified offline).  Reports the header reading that matched, the
        per-document fill, and — given a sequence known to be part of one document's training data (G1/G3: a training
        genome scores 1.0 on its own document) — its score there next to ``fill ** h``, the score a k-mer that is NOT in
        the document gets.  ``consistent`` is False when the member scores like a stranger: wrong hash seeds, modulus,
        bit order or row stride for this file."""
        fill = self.doc_fill()
        rep = {"header_layout": self.header_layout, "kind": "classic" if self.info.kind == _abi.XS_COBS_CLASSIC else "compact",
               "n_docs": self.n_docs, "num_hashes": self.num_hashes, "term_size": self.k,
               "fill_min": float(fill.min()), "fill_max": float(fill.max()),
               "design_fill_fpr_0.01": 0.01 ** (1.0 / self.num_hashes),
               "implausible_fill": bool(fill.max() > 0.75 or fill.max() < 0.02)}
        if member_sequence is not None:
            a = _as_bases(member_sequence)
            d = self.names.index(member_doc) if isinstance(member_doc, str) else int(member_doc or 0)
            n_win = max(0, (a.size - self.k) // step + 1)
            score = float(self.counts(a, step)[d]) / max(n_win, 1)
            stranger = float(fill[d]) ** self.num_hashes
            rep.update({"member_doc": self.names[d], "member_score": score, "stranger_score": stranger,
                        "consistent": bool(score > 0.5 and score > 4 * stranger) or bool(score > 0.9)})
        return rep

    def set_policy(self, policy: int) -> None:
        check(lib().xs_cobs_set_policy(self._h, int(policy)))
        self.info.policy = int(policy)

    def set_bucketed(self, enabled: bool = True, min_windows: int = 0, scratch_bytes: int = 0, bucket_shift: int = 0) -> None:
        """Tune the bucketed path large batches take on a large narrow-row index (see ``xs_cobs_set_bucketed``)."""
        check(lib().xs_cobs_set_bucketed(self._h, 1 if enabled else 0, int(min_windows), int(scratch_bytes), int(bucket_shift)))

    @property
    def bucketed_queries(self) -> int:
        n = C.c_uint64()
        check(lib().xs_cobs_bucketed_queries(self._h, C.byref(n)))
        return int(n.value)

    def close(self) -> None:
        h, self._h = self._h, C.c_void_p()
        if h:
            lib().xs_cobs_close(h)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def query(self, bases, seq_begin, seq_end, step: int = 1, dtype: int | None = None, out: np.ndarray | None = None) -> np.ndarray:
        """Per-sequence, per-document hit counts ``[n_seq, n_docs]`` (host buffers in, host array out)."""
        bases = _as_bases(bases)
        b, e = _as_u64(seq_begin), _as_u64(seq_end)
        n = b.size
        if e.size != n:
            raise ValueError("seq_begin and seq_end differ in length")
        if dtype is None:
            mx = int((e - b).max()) if n else 0
            dtype = pick_dtype(max(0, (mx - self.k) // step + 1))
        if out is None:
            small = n * self.n_docs * int(dtype) < (1 << 20)      # small results come back through the library's own staging
            out = np.empty((n, self.n_docs), _DT[dtype]) if small else pinned_empty((n, self.n_docs), _DT[dtype])
        elif out.dtype != _DT[dtype] or out.shape != (n, self.n_docs) or not out.flags.c_contiguous:
            raise ValueError("out has the wrong dtype/shape")
        check(lib().xs_cobs_query(self._h, _ptr(bases), bases.size, _ptr(b), _ptr(e), n, int(step), int(dtype), _ptr(out)))
        return out

    def classify(self, bases, seq_begin, seq_end, step: int = 1):
        """Read-level calls without the count matrix: ``(best, best_count, n_best, totals)`` — first document with
        the maximum count per record, that count, how many documents share it (> 1 = ambiguous), and per-document
        totals over all records.  The epilogue runs on the device; 12 bytes per record come back."""
        bases = _as_bases(bases)
        b, e = _as_u64(seq_begin), _as_u64(seq_end)
        n = b.size
        best, cnt, nb = (pinned_empty((n,), np.uint32) for _ in range(3))
        totals = np.zeros(self.n_docs, np.uint64)
        check(lib().xs_cobs_classify(self._h, _ptr(bases), bases.size, _ptr(b), _ptr(e), n, int(step), _ptr(best), _ptr(cnt),
                                     _ptr(nb), _ptr(totals)))
        return best, cnt, nb, totals

    def classify_file(self, path, fmt: int, step: int = 1, block_bytes: int = 0) -> dict:
        """Read-level calls of a FASTA (``fmt`` 1) / 4-line FASTQ (``fmt`` 2) file, streamed (``xs_cobs_classify_file``):
        parsing on all host threads, host-to-device copies, scoring and the argmax epilogue overlap block by block.
        Returns ``best, best_hits, n_best, seq_len, totals`` plus the raw record-id buffer (``id_buf``, ``id_end``)."""
        h = C.c_void_p()
        check(lib().xs_cobs_classify_file(self._h, str(path).encode(), int(fmt), int(step), int(block_bytes), C.byref(h)))
        owner = _FileCalls(h)
        n, nb, nid, nshort, nd = (C.c_uint64() for _ in range(5))
        ps, ts = C.c_double(), C.c_double()
        check(lib().xs_file_calls_info(h, C.byref(n), C.byref(nb), C.byref(nid), C.byref(nshort), C.byref(nd), C.byref(ps), C.byref(ts)))
        ptr = [C.c_void_p() for _ in range(7)]
        check(lib().xs_file_calls_view(h, *[C.byref(p) for p in ptr]))

        def view(p, count, dtype):     # zero-copy view of the library's result arrays; `owner` frees them with the last view
            if not count or not p.value:
                return np.empty(0, dtype)
            a = np.frombuffer((C.c_char * (count * np.dtype(dtype).itemsize)).from_address(p.value), dtype=dtype, count=count)
            owner.keep(a)
            return a
        best, cnt, nbst = (view(ptr[i], n.value, np.uint32) for i in range(3))
        seq_len, ids, id_end = view(ptr[3], n.value, np.uint64), view(ptr[4], nid.value, np.uint8), view(ptr[5], n.value, np.uint64)
        totals = view(ptr[6], nd.value, np.uint64).copy()
        return {"best": best, "best_hits": cnt, "n_best": nbst, "seq_len": seq_len, "totals": totals, "id_buf": ids, "id_end": id_end,
                "n_bases": int(nb.value), "n_short": int(nshort.value), "parse_s": ps.value, "total_s": ts.value, "_owner": owner}

    def query_device(self, d_bases: int, n_bases: int, d_begin: int, d_end: int, n_seq: int, step: int, dtype: int,
                     d_out: int, stream: int = 0, ld: int = 0) -> None:
        """Same with raw device pointers on this index's GPU; asynchronous on ``stream``.  ``ld`` > 0: output rows of
        that many elements (zero padded) instead of ``n_docs``."""
        if ld:
            check(lib().xs_cobs_query_device_ld(self._h, d_bases, n_bases, d_begin, d_end, n_seq, int(step), int(dtype), int(ld),
                                                d_out, stream))
        else:
            check(lib().xs_cobs_query_device(self._h, d_bases, n_bases, d_begin, d_end, n_seq, int(step), int(dtype), d_out, stream))

    def counts(self, sequence, step: int = 1) -> np.ndarray:
        """uint32 hit counts of one sequence."""
        a = _as_bases(sequence)
        return self.query(a, np.array([0], np.uint64), np.array([a.size], np.uint64), step, XS_U32)[0]

    def rows(self, sequence, step: int = 1):
        """Row ids ``[n_windows, num_hashes, n_pages]`` and the per-window valid flags (parity tests)."""
        a = _as_bases(sequence)
        n_w = (a.size - self.k) // step + 1 if a.size >= self.k else 0
        rows = np.zeros((n_w, self.num_hashes, self.info.n_pages), np.uint64)
        valid = np.zeros(n_w, np.uint8)
        if n_w:
            check(lib().xs_cobs_rows(self._h, _ptr(a), a.size, int(step), _ptr(rows), _ptr(valid)))
        return rows, valid

    @staticmethod
    def result_order(scores) -> np.ndarray:
        s = np.ascontiguousarray(scores, dtype=np.uint32)
        order = np.zeros(s.size, np.uint32)
        check(lib().xs_cobs_result_order(_ptr(s), s.size, _ptr(order)))
        return order


def result_order_batch(scores: np.ndarray) -> np.ndarray:
    """Per-row cobs result order of a [n_seq, n_docs] count matrix."""
    s = np.ascontiguousarray(scores, dtype=np.uint32)
    order = np.zeros(s.shape, np.uint32)
    if s.size:
        check(lib().xs_cobs_result_order_batch(_ptr(s), s.shape[0], s.shape[1], _ptr(order)))
    return order


class _FileCalls:
    """Owner of one ``xs_file_calls`` result: freed when the last numpy view of its arrays is gone."""

    def __init__(self, handle):
        self._h = handle
        self._views = []

    def keep(self, arr) -> None:
        import weakref
        self._views.append(weakref.ref(arr))
        # the array's base chain ends in a ctypes object created from an address; tie the owner to it
        base = arr
        while getattr(base, "base", None) is not None:
            base = base.base
        try:
            base._xs_owner = self
        except AttributeError:
            pass

    def __del__(self):
        h, self._h = self._h, None
        if h:
            try:
                lib().xs_file_calls_free(h)
            except Exception:
                pass


class SearchResult:
    """``doc_name`` / ``score`` pair, the shape probabilistic_filter_model.py:406-409 reads."""

    __slots__ = ("doc_name", "score")

    def __init__(self, doc_name: str, score: int):
        self.doc_name = doc_name
        self.score = score

    def __repr__(self):
        return f"SearchResult(doc_name={self.doc_name!r}, score={self.score})"


class Search:
    """Drop-in for ``cobs_index.Search(path, load_complete)``: the index lives in HBM either way."""

    def __init__(self, path, load_complete: bool = True, device: int = 0):
        self.index = CobsIndex(path, device=device)

    def search(self, query: str, step: int = 1):
        if len(query) < self.index.k:
            raise RuntimeError("query too short for the index term size")
        c = self.index.counts(query, step)
        return [SearchResult(self.index.names[i], int(c[i])) for i in CobsIndex.result_order(c)]


class Comm:
    """One rank's NCCL communicator for the document-column sharded exchange (``xs_comm_*``).  ``unique_id()`` on
    rank 0, ship the 128 bytes to the other ranks, ``Comm(id, rank, world, device)`` everywhere."""

    def __init__(self, uid: bytes, rank: int, world: int, device: int):
        self._h = C.c_void_p()
        buf = (C.c_uint8 * 128).from_buffer_copy(uid)
        check(lib().xs_comm_init(buf, int(rank), int(world), int(device), C.byref(self._h)))
        self.rank, self.world, self.device = rank, world, device

    @staticmethod
    def unique_id() -> bytes:
        buf = (C.c_uint8 * 128)()
        check(lib().xs_comm_unique_id(buf))
        return bytes(buf)

    @property
    def nccl_version(self) -> int:
        v = C.c_int()
        check(lib().xs_comm_info(self._h, None, None, C.byref(v)))
        return int(v.value)

    def allgather_scores(self, d_local: int, n_seq: int, row_bytes: int, d_all: int, stream: int = 0) -> None:
        check(lib().xs_allgather_scores(self._h, d_local, int(n_seq), int(row_bytes), d_all, stream))

    def allreduce_totals(self, d_totals: int, n: int, stream: int = 0) -> None:
        check(lib().xs_allreduce_totals(self._h, d_totals, int(n), stream))

    def close(self) -> None:
        h, self._h = self._h, C.c_void_p()
        if h:
            lib().xs_comm_destroy(h)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def sharded_reduce_device(d_all: int, n_seq: int, dtype: int, device: int, w: int, widths, d_best: int, d_best_count: int,
                          d_n_best: int, d_totals: int = 0, stream: int = 0) -> None:
    """Per-record first best document / count / tie multiplicity over ``[world][n_seq][w]`` all-gathered score blocks,
    in place (``xs_sharded_reduce_device``)."""
    wd = np.ascontiguousarray(widths, dtype=np.uint32)
    check(lib().xs_sharded_reduce_device(d_all, int(n_seq), int(dtype), int(device), wd.size, int(w), _ptr(wd), d_best or None,
                                         d_best_count or None, d_n_best or None, d_totals or None, stream))


def mlst_query(indices: list[CobsIndex], allele_len, bases, seq_begin, seq_end, step: int = 1,
               min_chunk_score: int = 50, chunk_from_len: int = 10000):
    """All loci of an MLST scheme over all records in one call (``xs_mlst_query``): chunking, per-chunk threshold and
    per-allele sums on the device.  Returns ``out[locus][record] = (doc_indices, scores)`` in the reference's
    result-dict order (probabilistic_filter_mlst_model.py:236-286)."""
    bases = _as_bases(bases)
    b, e = _as_u64(seq_begin), _as_u64(seq_end)
    n, nl = b.size, len(indices)
    handles = (C.c_void_p * nl)(*[ix._h for ix in indices])
    alen = np.ascontiguousarray(allele_len, dtype=np.uint32)
    docs = [ix.n_docs for ix in indices]
    total = n * int(sum(docs))
    out_n = np.zeros(nl * n, np.uint32)
    out_doc = np.empty(max(total, 1), np.uint32)
    out_score = np.empty(max(total, 1), np.uint32)
    check(lib().xs_mlst_query(handles, nl, _ptr(alen), _ptr(bases), bases.size, _ptr(b), _ptr(e), n, int(step),
                              int(min_chunk_score), int(chunk_from_len), _ptr(out_n), _ptr(out_doc), _ptr(out_score)))
    res, off = [], 0
    for li, d in enumerate(docs):
        per = []
        for i in range(n):
            m = int(out_n[li * n + i])
            o = off + i * d
            per.append((out_doc[o:o + m], out_score[o:o + m]))
        res.append(per)
        off += n * d
    return res


# --------------------------------------------------------------------------------------
# Bloom
# --------------------------------------------------------------------------------------
class BloomFilter:
    """An rbloom ``.bloom`` file resident in one GPU's HBM."""

    def __init__(self, path, k: int, device: int = 0):
        self._h = C.c_void_p()
        self.path = str(path)
        if not os.path.isfile(self.path):
            raise FileNotFoundError(f"Bloom filter file not found at {self.path}")
        check(lib().xs_bloom_open(self.path.encode(), int(k), int(device), C.byref(self._h)))
        info = _abi.BloomInfo()
        check(lib().xs_bloom_info(self._h, C.byref(info)))
        self.info = info

    k = property(lambda self: self.info.term_size)
    device = property(lambda self: self.info.device)

    def close(self) -> None:
        h, self._h = self._h, C.c_void_p()
        if h:
            lib().xs_bloom_close(h)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_bucketed(self, enabled: bool = True, min_windows: int = 0, scratch_bytes: int = 0, bucket_shift: int = 0,
                     member_pct: int = -1) -> None:
        """Tune the bucketed path large batches take (see ``xs_bloom_set_bucketed``); ``member_pct=0`` forces it
        whatever the sampled member fraction."""
        check(lib().xs_bloom_set_bucketed(self._h, 1 if enabled else 0, int(min_windows), int(scratch_bytes), int(bucket_shift),
                                          int(member_pct)))

    @property
    def bucketed_queries(self) -> int:
        n = C.c_uint64()
        check(lib().xs_bloom_bucketed_queries(self._h, C.byref(n)))
        return int(n.value)

    def query(self, bases, seq_begin, seq_end, step: int = 1, out: np.ndarray | None = None) -> np.ndarray:
        """Hits per sequence ``[n_seq]`` uint32."""
        bases = _as_bases(bases)
        b, e = _as_u64(seq_begin), _as_u64(seq_end)
        if out is None:
            out = np.empty(b.size, np.uint32) if b.size < (1 << 18) else pinned_empty((b.size,), np.uint32)
        check(lib().xs_bloom_query(self._h, _ptr(bases), bases.size, _ptr(b), _ptr(e), b.size, int(step), _ptr(out)))
        return out

    def query_device(self, d_bases: int, n_bases: int, d_begin: int, d_end: int, n_seq: int, step: int, d_out: int,
                     stream: int = 0) -> None:
        check(lib().xs_bloom_query_device(self._h, d_bases, n_bases, d_begin, d_end, n_seq, int(step), d_out, stream))

    def hits(self, sequence, step: int = 1) -> int:
        a = _as_bases(sequence)
        return int(self.query(a, np.array([0], np.uint64), np.array([a.size], np.uint64), step)[0])

    def contains(self, terms) -> np.ndarray:
        """Membership of k-mers exactly as given (no canonicalisation): ``terms`` is a list of str/bytes of length k
        or a uint8 array [n, k]; returns bool [n]."""
        if isinstance(terms, np.ndarray):
            arr = np.ascontiguousarray(terms, dtype=np.uint8).reshape(-1, self.k)
        else:
            enc = [t.encode("utf-8") if isinstance(t, str) else bytes(t) for t in terms]
            if any(len(t) != self.k for t in enc):
                raise ValueError("k-mer length differs from the model's k")
            arr = np.frombuffer(b"".join(enc), dtype=np.uint8).reshape(-1, self.k) if enc else np.zeros((0, self.k), np.uint8)
        out = np.zeros(arr.shape[0], np.uint8)
        check(lib().xs_bloom_contains(self._h, _ptr(arr), arr.shape[0], _ptr(out)))
        return out.astype(bool)

    def hashes(self, sequence, step: int = 1) -> np.ndarray:
        a = _as_bases(sequence)
        n_w = (a.size - self.k) // step + 1 if a.size >= self.k else 0
        out = np.zeros(n_w, np.uint64)
        if n_w:
            check(lib().xs_bloom_hashes(self._h, _ptr(a), a.size, int(step), _ptr(out)))
        return out


class Bloom:
    """Drop-in for the loaded ``rbloom.Bloom``: ``kmer in bf`` for a k-mer string of the model's k, hashed exactly
    as given (the caller canonicalises, like _generate_kmers does)."""

    def __init__(self, path, k: int, device: int = 0):
        self.filter = BloomFilter(path, k, device)

    @classmethod
    def load(cls, path, k: int, device: int = 0):
        return cls(path, k, device)

    def __contains__(self, kmer) -> bool:
        return bool(self.filter.contains([str(kmer)])[0])


# --------------------------------------------------------------------------------------
# stage helpers (parity tests)
# --------------------------------------------------------------------------------------
def pack_2bit(bases, device: int = 0):
    a = _as_bases(bases)
    n_words = a.size // 32 + 1
    packed = np.zeros(n_words, np.uint64)
    invalid = np.zeros(n_words, np.uint32)
    check(lib().xs_pack_2bit(_ptr(a), a.size, device, _ptr(packed), _ptr(invalid)))
    return packed, invalid


def canonical_kmers(bases, k: int, device: int = 0):
    a = _as_bases(bases)
    n_w = max(0, a.size - k + 1)
    codes = np.zeros(n_w, np.uint64)
    valid = np.zeros(n_w, np.uint8)
    check(lib().xs_canonical_kmers(_ptr(a), a.size, k, device, _ptr(codes), _ptr(valid)))
    return codes, valid


def kmer_rows(bases, k: int, num_hashes: int, sig_size: int, canonicalize: int = 1, step: int = 1, device: int = 0):
    """Row ids ``[n_windows, num_hashes]`` of a sequence for an explicit classic geometry (no index handle)."""
    a = _as_bases(bases)
    n_w = (a.size - k) // step + 1 if a.size >= k else 0
    rows = np.zeros((n_w, num_hashes), np.uint64)
    valid = np.zeros(n_w, np.uint8)
    if n_w:
        check(lib().xs_kmer_rows(_ptr(a), a.size, k, canonicalize, num_hashes, sig_size, step, device, _ptr(rows), _ptr(valid)))
    return rows, valid


def scores_reduce_device(d_counts: int, n_seq: int, n_docs: int, dtype: int, device: int, d_best: int = 0,
                         d_best_count: int = 0, d_n_best: int = 0, d_totals: int = 0, stream: int = 0) -> None:
    """Device epilogue over a [n_seq, n_docs] count matrix: best document / its count / tie multiplicity per record
    (uint32 each) and per-document totals (uint64, accumulated).  Raw device pointers; 0 = not wanted."""
    check(lib().xs_scores_reduce_device(d_counts, n_seq, n_docs, int(dtype), int(device), d_best or None, d_best_count or None,
                                        d_n_best or None, d_totals or None, stream or None))


def build_cobs(path, kind: int, k: int, num_hashes: int, fpr: float, names: list[str], bases, seq_begin, seq_end, seq_doc,
               canonicalize: int = 1, sig_size: int = 0, page_size: int = 0, device: int = 0) -> None:
    """Construct a COBS classic (kind 1) / compact (kind 2) index file on the GPU: sequence i belongs to document
    ``seq_doc[i]`` (index into ``names``)."""
    bases = _as_bases(bases)
    b, e = _as_u64(seq_begin), _as_u64(seq_end)
    doc = np.ascontiguousarray(seq_doc, dtype=np.uint32)
    if any("\n" in n for n in names):
        raise ValueError("document names must not contain line breaks")
    blob = ("\n".join(names) + "\n").encode("utf-8")
    check(lib().xs_cobs_build(str(path).encode(), int(device), int(kind), int(k), int(num_hashes), float(fpr), int(canonicalize),
                              int(sig_size), int(page_size), blob, len(names), _ptr(bases), bases.size, _ptr(b), _ptr(e), _ptr(doc), b.size))


def build_bloom(path, k: int, expected_items: int, fpr: float, bases, seq_begin, seq_end, device: int = 0) -> None:
    """Construct an rbloom filter file on the GPU from every k-mer of the given sequences."""
    bases = _as_bases(bases)
    b, e = _as_u64(seq_begin), _as_u64(seq_end)
    check(lib().xs_bloom_build(str(path).encode(), int(device), int(k), int(expected_items), float(fpr), _ptr(bases), bases.size,
                               _ptr(b), _ptr(e), b.size))


def device_count() -> int:
    n = C.c_int()
    rc = lib().xs_device_count(C.byref(n))
    return n.value if rc == 0 else 0


def launch_count() -> int:
    return int(lib().xs_launch_count())


def profile_enable(on: bool = True) -> None:
    check(lib().xs_profile_enable(1 if on else 0))


def profile_read() -> tuple[float, int]:
    """(summed dominant-kernel milliseconds, launches) since the last read; synchronises the events."""
    ms, n = C.c_double(), C.c_uint64()
    check(lib().xs_profile_read(C.byref(ms), C.byref(n)))
    return ms.value, n.value


def profile_read_phases() -> tuple[list[float], list[int]]:
    """Per-kernel split of :func:`profile_read`: [direct kernel, k_bucket_emit, k_bucket_fetch, k_bucket_reduce]."""
    ms, n = (C.c_double * 4)(), (C.c_uint64 * 4)()
    check(lib().xs_profile_read_phases(ms, n))
    return list(ms), [int(x) for x in n]
