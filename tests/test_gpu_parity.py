"""GPU parity: every count the CUDA path produces equals the CPU oracle's on the same seeded inputs.
All calls go through the C ABI (xspect2_b200.engine -> libxspect_b200.so).  Bit-exact (integer path)."""
import numpy as np
import pytest

from tests import synth

pytestmark = pytest.mark.gpu


def _mk_classic(oracle, tmp_path, rng, n_docs, k, h, length=3000, name="index.cobs_classic", fpr=0.01, **kw):
    docs = synth.make_genomes(rng, n_docs, length)
    p = tmp_path / name
    oracle.write_classic(p, docs, k=k, num_hashes=h, fpr=fpr, **kw)
    return p, docs


def _check_cobs(gpu, oracle, path, bases, b, e, step=1, dtype=None, policy=0, kernel=None, **open_kw):
    ix = gpu.CobsIndex(path, **open_kw)
    if kernel is not None:
        assert ix.kernel == kernel, (ix.kernel, ix.info.row_stride)
    ix.set_policy(policy)
    orc = oracle.CobsOracle(path, policy=policy)
    got = ix.query(bases, b, e, step=step, dtype=dtype)
    exp = orc.counts_batch(bases, b, e, step=step, threads=4)
    lo, hi = ix.info.doc_begin, ix.info.doc_end
    exp = exp[:, lo:hi]
    if dtype in (1, 2):
        exp = np.minimum(exp, 255 if dtype == 1 else 65535)
    assert got.shape == exp.shape
    bad = np.argwhere(got.astype(np.uint32) != exp)
    assert bad.size == 0, f"{bad.shape[0]} mismatches, first {bad[:5].tolist()}: got {got[tuple(bad[0])]} exp {exp[tuple(bad[0])]}"
    ix.close()
    return got


# ------------------------------------------------------------------------------------------ stages
def test_pack_2bit(gpu):
    rng = np.random.default_rng(1)
    for n in (0, 1, 31, 32, 33, 1000, 4097):
        a = synth.mutate(rng, synth.random_dna(rng, n), n_rate=0.05, lower=0.05, iupac=0.02)
        packed, invalid = gpu.pack_2bit(a)
        for i in range(n):
            ok = a[i] in b"ACGT"
            assert ((int(invalid[i // 32]) >> (i % 32)) & 1) == (0 if ok else 1)
            if ok:
                assert ((int(packed[i // 32]) >> (2 * (i % 32))) & 3) == b"ACGT".index(a[i])


@pytest.mark.parametrize("k", [1, 8, 21, 31, 32])
def test_canonical_kmers(gpu, oracle, k):
    rng = np.random.default_rng(2)
    a = synth.mutate(rng, synth.random_dna(rng, 700), n_rate=0.01)
    codes, valid = gpu.canonical_kmers(a, k)
    raw = a.tobytes()
    for p in range(len(raw) - k + 1):
        t = oracle.cobs_term(raw[p:p + k])
        assert valid[p] == (0 if t is None else 1)
        if t is not None:
            c = 0
            for ch in t:
                c = (c << 2) | b"ACGT".index(ch)
            assert int(codes[p]) == c


@pytest.mark.parametrize("k,h,step", [(21, 7, 1), (31, 1, 1), (21, 7, 3), (17, 3, 2), (32, 2, 1)])
def test_cobs_row_ids(gpu, oracle, tmp_path, k, h, step):
    rng = np.random.default_rng(3)
    p, _ = _mk_classic(oracle, tmp_path, rng, 9, k, h, length=500)
    a = synth.mutate(rng, synth.random_dna(rng, 900), n_rate=0.01)
    ix = gpu.CobsIndex(p)
    rows, valid = ix.rows(a, step)
    erows, evalid = oracle.CobsOracle(p).rows(a, step)
    assert np.array_equal(valid, evalid)
    assert np.array_equal(rows, erows)


# ------------------------------------------------------------------------------------------ bucketed probing
def _bucket_shift(path, gpu):
    """Rows per bucket (log2) that gives a small test index 100-250 buckets (the shipped geometry has 144)."""
    ix = gpu.CobsIndex(path)
    s = int(ix.info.sig_size_max)
    ix.close()
    shift = 1
    while ((s - 1) >> shift) + 1 > 250:
        shift += 1
    return shift


def _check_bucketed(gpu, oracle, path, bases, b, e, step=1, dtype=None, policy=0, scratch=0, shift=None):
    ix = gpu.CobsIndex(path)
    ix.set_policy(policy)
    ix.set_bucketed(True, min_windows=1, scratch_bytes=scratch, bucket_shift=shift or _bucket_shift(path, gpu))
    got = ix.query(bases, b, e, step=step, dtype=dtype)
    assert ix.bucketed_queries >= 1, "the bucketed kernels did not run"
    exp = oracle.CobsOracle(path, policy=policy).counts_batch(bases, b, e, step=step, threads=4)
    if dtype in (1, 2):
        exp = np.minimum(exp, 255 if dtype == 1 else 65535)
    bad = np.argwhere(got.astype(np.uint32) != exp)
    assert bad.size == 0, f"{bad.shape[0]} mismatches, first {bad[:5].tolist()}: got {got[tuple(bad[0])]} exp {exp[tuple(bad[0])]}"
    # and the direct-gather kernel agrees
    ix.set_bucketed(False)
    n0 = ix.bucketed_queries
    assert np.array_equal(ix.query(bases, b, e, step=step, dtype=dtype), got)
    assert ix.bucketed_queries == n0
    ix.close()
    return got


@pytest.mark.parametrize("n_docs", [8, 90, 97, 128])
@pytest.mark.parametrize("k,h", [(21, 7), (31, 1), (15, 3)])
def test_cobs_bucketed_reads(gpu, oracle, tmp_path, n_docs, k, h):
    """Ragged reads with N / substitutions through k_bucket_emit / fetch / reduce (both row packings: <= 96 and
    > 96 documents), several chunks, records straddling chunk and warp boundaries."""
    rng = np.random.default_rng(300 + n_docs)
    p, docs = _mk_classic(oracle, tmp_path, rng, n_docs, k, h)
    genomes = [s for v in docs.values() for s in v]
    bases, b, e = synth.sample_reads(rng, genomes, 700, (k, 300), sub=0.01, n_rate=0.003)
    got = _check_bucketed(gpu, oracle, p, bases, b, e)
    assert got.sum() > 0


@pytest.mark.parametrize("step", [1, 3, 500])
@pytest.mark.parametrize("dtype", [None, 1])
def test_cobs_bucketed_long_and_low_complexity(gpu, oracle, tmp_path, step, dtype):
    """Contigs spanning many chunks, homopolymer / short-period runs that overflow single blocks (direct-gather
    windows), empty and short records in between, saturating uint8 counts on long records."""
    rng = np.random.default_rng(17)
    p, docs = _mk_classic(oracle, tmp_path, rng, 12, 21, 7, length=20000)
    g = docs["doc00003"][0]
    rep = np.tile(np.frombuffer(b"ACGTTGCA", np.uint8), 900)
    parts = [synth.mutate(rng, g, n_rate=0.0005), np.full(6000, ord("A"), np.uint8), g[:0], g[:20], g[:21], rep, g[100:250],
             np.full(300, ord("N"), np.uint8), g[:5000] | 0x20, synth.random_dna(rng, 2500), g[::-1].copy()]
    bases = np.concatenate(parts)
    lens = np.array([x.size for x in parts], np.uint64)
    e = np.cumsum(lens, dtype=np.uint64)
    b = e - lens
    _check_bucketed(gpu, oracle, p, bases, b, e, step=step, dtype=dtype)
    if step == 1:
        _check_bucketed(gpu, oracle, p, bases, b, e, step=step, dtype=dtype, policy=1)


def test_cobs_bucketed_sub_batches_and_overlaps(gpu, oracle, tmp_path):
    """A scratch budget of a few chunks (many sub-batches) and overlapping / unordered segments (MLST-style)."""
    rng = np.random.default_rng(19)
    p, docs = _mk_classic(oracle, tmp_path, rng, 90, 21, 7, length=4000)
    genomes = [s for v in docs.values() for s in v]
    bases, b, e = synth.sample_reads(rng, genomes, 900, 150, sub=0.002, n_rate=0.0005)
    _check_bucketed(gpu, oracle, p, bases, b, e, scratch=3 << 20)
    g = np.concatenate(genomes[:3])
    n = 400
    bb = rng.integers(0, g.size - 600, n).astype(np.uint64)
    ee = bb + rng.integers(0, 600, n).astype(np.uint64)
    _check_bucketed(gpu, oracle, p, g, bb, ee, scratch=3 << 20)


def test_cobs_bucketed_runs_of_empty_records(gpu, oracle, tmp_path):
    """Thousands of zero-window records inside and between chunks (the chunk-local sequence table skips them)."""
    rng = np.random.default_rng(29)
    p, docs = _mk_classic(oracle, tmp_path, rng, 10, 21, 7)
    g = docs["doc00002"][0]
    parts = [g[:300]] + [g[:5]] * 3000 + [g[300:2900]] + [g[:0]] * 2500 + [g[700:1000]] + [g[:20]] * 600 + [g[:2200]]
    bases = np.concatenate(parts)
    lens = np.array([x.size for x in parts], np.uint64)
    e = np.cumsum(lens, dtype=np.uint64)
    _check_bucketed(gpu, oracle, p, bases, e - lens, e)
    _check_bucketed(gpu, oracle, p, bases, e - lens, e, dtype=1, scratch=2 << 20)


def test_cobs_bucketed_not_canonical_index_and_column_shard(gpu, oracle, tmp_path):
    """A non-canonical index (literal k-mers, windows with N are hashed as they are) and a document-column shard of a
    wider index (rows of the shard <= 16 bytes) through the bucketed kernels."""
    rng = np.random.default_rng(37)
    p, docs = _mk_classic(oracle, tmp_path, rng, 6, 21, 3, canonicalize=0)
    genomes = [s for v in docs.values() for s in v]
    bases, b, e = synth.sample_reads(rng, genomes, 500, (21, 260), sub=0.01, n_rate=0.004)
    _check_bucketed(gpu, oracle, p, bases, b, e)
    p2, docs2 = _mk_classic(oracle, tmp_path, rng, 200, 21, 7, length=1500, name="wide.cobs_classic")
    genomes2 = [s for v in docs2.values() for s in v]
    bases, b, e = synth.sample_reads(rng, genomes2, 400, (21, 200), sub=0.01, n_rate=0.002)
    exp = oracle.CobsOracle(p2).counts_batch(bases, b, e, 1, threads=4)
    for lo, hi in ((0, 96), (96, 200)):
        ix = gpu.CobsIndex(p2, doc_begin=lo, doc_end=hi)
        ix.set_bucketed(True, min_windows=1, bucket_shift=_bucket_shift(p2, gpu))
        got = ix.query(bases, b, e)
        assert ix.bucketed_queries >= 1
        assert np.array_equal(np.asarray(got).astype(np.uint32), exp[:, lo:hi])
        ix.close()


def test_cobs_bucketed_shift_21_and_few_buckets(gpu, oracle, tmp_path):
    """Other geometries of the record word: 2^21 rows per bucket (one bucket here) and 2 rows per bucket."""
    rng = np.random.default_rng(23)
    p, docs = _mk_classic(oracle, tmp_path, rng, 30, 21, 7, length=40)     # tiny signature: a few hundred rows
    genomes = [s for v in docs.values() for s in v]
    bases, b, e = synth.sample_reads(rng, genomes, 300, (21, 40), frac_random=0.1)
    _check_bucketed(gpu, oracle, p, bases, b, e, shift=21)
    _check_bucketed(gpu, oracle, p, bases, b, e, shift=_bucket_shift(p, gpu))


# ------------------------------------------------------------------------------------------ narrow rows
@pytest.mark.parametrize("n_docs", [1, 7, 8, 9, 90, 128])
@pytest.mark.parametrize("k,h", [(21, 7), (31, 1), (15, 3)])
def test_cobs_narrow_reads(gpu, oracle, tmp_path, n_docs, k, h):
    rng = np.random.default_rng(100 + n_docs)
    p, docs = _mk_classic(oracle, tmp_path, rng, n_docs, k, h)
    genomes = [s for v in docs.values() for s in v]
    bases, b, e = synth.sample_reads(rng, genomes, 600, (k, 300), sub=0.01, n_rate=0.003)
    got = _check_cobs(gpu, oracle, p, bases, b, e)
    assert got.sum() > 0


@pytest.mark.parametrize("step", [1, 2, 3, 4, 500])
def test_cobs_steps_and_long_sequences(gpu, oracle, tmp_path, step):
    rng = np.random.default_rng(7)
    p, docs = _mk_classic(oracle, tmp_path, rng, 12, 21, 7, length=20000)
    g = docs["doc00003"][0]
    # one long contig (many tiles), a medium one, reads, and an exact-k sequence
    parts = [synth.mutate(rng, g, n_rate=0.0005), g[:5000], g[100:250], g[:21], synth.random_dna(rng, 2500)]
    bases = np.concatenate(parts)
    lens = np.array([x.size for x in parts], np.uint64)
    e = np.cumsum(lens, dtype=np.uint64)
    b = e - lens
    got = _check_cobs(gpu, oracle, p, bases, b, e, step=step)
    if step == 1:
        assert got.dtype == np.uint16 and got[1, 3] == 5000 - 20  # true positives: every k-mer of doc 3


def test_cobs_golden_g1_shape(gpu, oracle, tmp_path):
    """Shape of the reference's G1/G2 known answers (tests/test_probabilistic_filter_model.py:73-93,149-161):
    an 80-bp substring of one training genome scores 60/step on it."""
    rng = np.random.default_rng(9)
    p, docs = _mk_classic(oracle, tmp_path, rng, 3, 21, 7, length=30000)
    q = docs["doc00000"][0][1000:1080]
    ix = gpu.CobsIndex(p)
    for step in (1, 2, 3, 4):
        c = ix.counts(q, step)
        assert c[0] == -(-60 // step)
        assert np.array_equal(c, oracle.CobsOracle(p).counts(q, step))


def test_cobs_edge_sequences(gpu, oracle, tmp_path):
    rng = np.random.default_rng(11)
    p, docs = _mk_classic(oracle, tmp_path, rng, 20, 21, 7)
    g = docs["doc00001"][0]
    # empty, shorter than k, exactly k, all-N, lower-case, overlapping segments, unordered offsets
    parts = [g[:0], g[:20], g[:21], np.full(100, ord("N"), np.uint8), g[:200] | 0x20, g[:400]]
    bases = np.concatenate(parts)
    lens = np.array([x.size for x in parts], np.uint64)
    e = np.cumsum(lens, dtype=np.uint64)
    b = e - lens
    b = np.concatenate([b, b[-1:] + 50, b[-1:] + 10, b[-1:]]).astype(np.uint64)
    e = np.concatenate([e, e[-1:] - 50, e[-1:] - 300, e[-1:]]).astype(np.uint64)
    _check_cobs(gpu, oracle, p, bases, b, e)
    _check_cobs(gpu, oracle, p, bases, b, e, policy=1)   # LITERAL non-ACGT policy
    # no sequences at all
    ix = gpu.CobsIndex(p)
    assert ix.query(bases, np.zeros(0, np.uint64), np.zeros(0, np.uint64)).shape == (0, 20)


def test_cobs_many_empty_sequences_fallback_tile(gpu, oracle, tmp_path):
    """More than a tile's worth of zero-window sequences between real ones (staging fallback path)."""
    rng = np.random.default_rng(12)
    p, docs = _mk_classic(oracle, tmp_path, rng, 10, 21, 7)
    g = docs["doc00002"][0]
    parts = [g[:300]] + [g[:5]] * 3000 + [g[300:700]] + [g[:0]] * 2500 + [g[700:1000]]
    bases = np.concatenate(parts)
    lens = np.array([x.size for x in parts], np.uint64)
    e = np.cumsum(lens, dtype=np.uint64)
    _check_cobs(gpu, oracle, p, bases, e - lens, e)
    _check_cobs(gpu, oracle, p, bases, e - lens, e, dtype=1)


@pytest.mark.parametrize("dtype", [1, 2, 4])
def test_cobs_output_dtypes_saturate(gpu, oracle, tmp_path, dtype):
    rng = np.random.default_rng(13)
    p, docs = _mk_classic(oracle, tmp_path, rng, 5, 21, 7, length=4000)
    g = docs["doc00004"][0]
    parts = [g, g[:150], g[:1500], synth.random_dna(rng, 150)]   # 3980 windows saturate u8
    bases = np.concatenate(parts)
    lens = np.array([x.size for x in parts], np.uint64)
    e = np.cumsum(lens, dtype=np.uint64)
    _check_cobs(gpu, oracle, p, bases, e - lens, e, dtype=dtype)


def test_cobs_not_canonical_index(gpu, oracle, tmp_path):
    rng = np.random.default_rng(14)
    p, docs = _mk_classic(oracle, tmp_path, rng, 6, 21, 3, canonicalize=0)
    genomes = [s for v in docs.values() for s in v]
    bases, b, e = synth.sample_reads(rng, genomes, 200, 150, n_rate=0.01)
    _check_cobs(gpu, oracle, p, bases, b, e)


# ------------------------------------------------------------------------------------------ wide rows
@pytest.mark.parametrize("n_docs,k,h", [(129, 21, 7), (300, 21, 7), (1000, 31, 1), (2100, 21, 3), (5000, 21, 7)])
def test_cobs_wide_reads(gpu, oracle, tmp_path, n_docs, k, h):
    rng = np.random.default_rng(200 + n_docs)
    p, docs = _mk_classic(oracle, tmp_path, rng, n_docs, k, h, length=400 if n_docs > 1000 else 800)
    genomes = [s for v in docs.values() for s in v]
    bases, b, e = synth.sample_reads(rng, genomes, 120, (k, 260), sub=0.01, n_rate=0.003)
    long_seq = np.concatenate([genomes[1], genomes[5], genomes[7]])
    bases = np.concatenate([bases, long_seq])
    b = np.concatenate([b, [e[-1]]]).astype(np.uint64)
    e = np.concatenate([e, [e[-1] + long_seq.size]]).astype(np.uint64)
    got = _check_cobs(gpu, oracle, p, bases, b, e)
    assert got.sum() > 0
    _check_cobs(gpu, oracle, p, bases, b, e, step=3, dtype=1)


def test_cobs_wide_multi_column_block(gpu, oracle, tmp_path):
    """20 000 documents = 157 column chunks = two column blocks (grid.y)."""
    rng = np.random.default_rng(21)
    p, docs = _mk_classic(oracle, tmp_path, rng, 20000, 21, 7, length=150)
    genomes = [s for v in docs.values() for s in v]
    bases, b, e = synth.sample_reads(rng, genomes[:50], 60, (21, 140), n_rate=0.003)
    long_seq = np.concatenate(genomes[:6])          # 900 bp: several 248-window work items
    bases = np.concatenate([bases, long_seq])
    b = np.concatenate([b, [e[-1]]]).astype(np.uint64)
    e = np.concatenate([e, [e[-1] + long_seq.size]]).astype(np.uint64)
    got = _check_cobs(gpu, oracle, p, bases, b, e)
    assert got.sum() > 0
    _check_cobs(gpu, oracle, p, bases, b, e, step=2, dtype=1)
    _check_cobs(gpu, oracle, p, bases, b, e, doc_begin=4096, doc_end=12288)


@pytest.mark.parametrize("n_docs,k,h", [(129, 21, 7), (200, 21, 7), (256, 31, 1), (300, 21, 7), (300, 15, 3), (512, 21, 7),
                                        (1000, 21, 7), (1000, 32, 8), (1024, 13, 2)])
def test_cobs_mid_rows(gpu, oracle, tmp_path, monkeypatch, n_docs, k, h):
    """Rows at a 32 / 64 / 128-byte stride take k_cobs_mid (lane groups per row): ragged reads with N / lower case / IUPAC,
    empty and shorter-than-k records, a contig that spans many warp tiles (added into by several warps) and saturates
    uint8, sampling steps; against the oracle and against k_cobs_wide on the same handle geometry."""
    rng = np.random.default_rng(1000 + n_docs + h)
    p, docs = _mk_classic(oracle, tmp_path, rng, n_docs, k, h, length=400)
    genomes = [s for v in docs.values() for s in v]
    bases, b, e = synth.sample_reads(rng, genomes, 400, (1, 260), n_rate=0.004, lower=0.002, iupac=0.002)
    contig = np.concatenate(genomes[:40])                      # 16 000 bp of documents 0..39: counts beyond 255
    empties = 70                                               # a run of empty records inside the batch
    b = np.concatenate([b, np.full(empties, e[-1]), [e[-1]]]).astype(np.uint64)
    e = np.concatenate([e, np.full(empties, e[-1]), [e[-1] + contig.size]]).astype(np.uint64)
    bases = np.concatenate([bases, contig])
    got = _check_cobs(gpu, oracle, p, bases, b, e, kernel="k_cobs_mid")
    assert got.sum() > 0 and got[-1].max() > 255
    _check_cobs(gpu, oracle, p, bases, b, e, dtype=1, kernel="k_cobs_mid")
    _check_cobs(gpu, oracle, p, bases, b, e, step=3, dtype=2, kernel="k_cobs_mid")
    _check_cobs(gpu, oracle, p, bases, b, e, policy=1, kernel="k_cobs_mid")
    monkeypatch.setenv("XS_NO_MID_KERNEL", "1")
    wide = _check_cobs(gpu, oracle, p, bases, b, e, kernel="k_cobs_wide")
    assert np.array_equal(wide, got)


def test_cobs_mid_rows_kernel_choice(gpu, oracle, tmp_path):
    """Column shards choose by their own width; more than 8 hash functions stay on k_cobs_wide."""
    rng = np.random.default_rng(77)
    p, docs = _mk_classic(oracle, tmp_path, rng, 2000, 21, 7, length=150)
    genomes = [s for v in docs.values() for s in v]
    bases, b, e = synth.sample_reads(rng, genomes, 200, (21, 150), n_rate=0.002)
    _check_cobs(gpu, oracle, p, bases, b, e, kernel="k_cobs_wide")                                   # 250-byte rows
    _check_cobs(gpu, oracle, p, bases, b, e, doc_begin=256, doc_end=768, kernel="k_cobs_mid")        # 64 bytes
    _check_cobs(gpu, oracle, p, bases, b, e, doc_begin=1024, doc_end=2000, kernel="k_cobs_mid")      # 122 -> 128 bytes
    _check_cobs(gpu, oracle, p, bases, b, e, doc_begin=128, doc_end=256, kernel="k_cobs_narrow")     # 16 bytes
    p9, docs9 = _mk_classic(oracle, tmp_path, rng, 300, 21, 9, length=150, name="h9.cobs_classic")
    _check_cobs(gpu, oracle, p9, bases, b, e, kernel="k_cobs_wide")


def test_cobs_wide_kernel_on_narrow_index(gpu, oracle, tmp_path, monkeypatch):
    rng = np.random.default_rng(15)
    p, docs = _mk_classic(oracle, tmp_path, rng, 90, 21, 7)
    genomes = [s for v in docs.values() for s in v]
    bases, b, e = synth.sample_reads(rng, genomes, 300, (21, 400), n_rate=0.002)
    monkeypatch.setenv("XS_FORCE_WIDE", "1")
    _check_cobs(gpu, oracle, p, bases, b, e)


@pytest.mark.parametrize("n_docs,lo,hi", [(90, 0, 48), (90, 48, 90), (1000, 256, 768), (1000, 768, 1000)])
def test_cobs_document_column_shards(gpu, oracle, tmp_path, n_docs, lo, hi):
    rng = np.random.default_rng(16)
    p, docs = _mk_classic(oracle, tmp_path, rng, n_docs, 21, 7, length=600)
    genomes = [s for v in docs.values() for s in v]
    bases, b, e = synth.sample_reads(rng, genomes, 150, 150)
    _check_cobs(gpu, oracle, p, bases, b, e, doc_begin=lo, doc_end=hi)


@pytest.mark.parametrize("n_docs", [5, 90, 1000])
@pytest.mark.parametrize("dtype", [1, 2, 4])
def test_scores_reduce_epilogue(gpu, n_docs, dtype):
    """Device argmax / tie multiplicity / totals equal NumPy on random count matrices (incl. all-zero rows)."""
    import torch
    rng = np.random.default_rng(n_docs + dtype)
    npdt = {1: np.uint8, 2: np.uint16, 4: np.uint32}[dtype]
    n = 3001
    counts = rng.integers(0, 6, size=(n, n_docs)).astype(npdt)
    counts[::7] = 0
    counts[5, n_docs - 1] = np.iinfo(npdt).max
    dev = torch.device("cuda", 0)
    tdt = {1: torch.uint8, 2: torch.uint16, 4: torch.uint32}[dtype]
    d_counts = torch.from_numpy(counts).to(dev)
    assert d_counts.dtype == tdt
    best = torch.empty(n, dtype=torch.int32, device=dev)
    cnt = torch.empty(n, dtype=torch.int32, device=dev)
    nb = torch.empty(n, dtype=torch.int32, device=dev)
    tot = torch.zeros(n_docs, dtype=torch.int64, device=dev)
    gpu.scores_reduce_device(d_counts.data_ptr(), n, n_docs, dtype, 0, best.data_ptr(), cnt.data_ptr(), nb.data_ptr(),
                             tot.data_ptr(), torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    c64 = counts.astype(np.int64)
    mx = c64.max(axis=1)
    assert np.array_equal(best.cpu().numpy().view(np.uint32), c64.argmax(axis=1).astype(np.uint32))
    assert np.array_equal(cnt.cpu().numpy().view(np.uint32).astype(np.int64), mx)
    assert np.array_equal(nb.cpu().numpy(), (c64 == mx[:, None]).sum(axis=1))
    assert np.array_equal(tot.cpu().numpy(), c64.sum(axis=0))


# ------------------------------------------------------------------------------------------ compact (MLST)
@pytest.mark.parametrize("n_alleles,k,page_size", [(40, 21, None), (600, 31, None), (600, 21, 4), (3000, 31, 32)])
def test_cobs_compact(gpu, oracle, tmp_path, n_alleles, k, page_size):
    rng = np.random.default_rng(300 + n_alleles)
    cons = synth.random_dna(rng, 450)
    docs = {}
    for a in range(n_alleles):
        s = cons.copy()
        for pos in rng.integers(0, 450, size=int(rng.integers(1, 6))):
            s[pos] = synth.ACGT[rng.integers(0, 4)]
        L = int(rng.integers(400, 451))
        docs[f"Allele_ID_{a + 1}"] = [s[:L]]
    p = tmp_path / "locus.cobs_compact"
    meta = oracle.write_compact(p, docs, k=k, num_hashes=1, fpr=0.001, page_size=page_size)
    genome = np.concatenate([synth.random_dna(rng, 5000), docs["Allele_ID_4"][0], synth.random_dna(rng, 5000)])
    chunks_b = np.arange(0, genome.size - 450, 450 - k + 1, dtype=np.uint64)
    chunks_e = np.minimum(chunks_b + 450, genome.size).astype(np.uint64)
    got = _check_cobs(gpu, oracle, p, genome, chunks_b, chunks_e)
    ix = gpu.CobsIndex(p)
    assert ix.names == meta["names"]
    col = ix.names.index("Allele_ID_4")
    assert got[:, col].sum() >= 400 - k + 1   # the planted allele, possibly split over two chunks


# ------------------------------------------------------------------------------------------ Bloom
def _mk_bloom(oracle, tmp_path, rng, k, length=20000, **kw):
    g = synth.random_dna(rng, length)
    p = tmp_path / "filter.bloom"
    meta = oracle.write_bloom(p, [g], k=k, **kw)
    return p, g, meta


@pytest.mark.parametrize("k", [21, 31, 13])
@pytest.mark.parametrize("step", [1, 3])
def test_bloom_reads(gpu, oracle, tmp_path, k, step):
    rng = np.random.default_rng(400 + k)
    p, g, meta = _mk_bloom(oracle, tmp_path, rng, k)
    assert meta["k_hashes"] == 6
    bases, b, e = synth.sample_reads(rng, [g], 500, (k, 300), sub=0.01, n_rate=0.004, lower=0.002, iupac=0.002)
    bf = gpu.BloomFilter(p, k)
    got = bf.query(bases, b, e, step)
    exp = oracle.BloomOracle(p, k).hits_batch(bases, b, e, step, threads=4)
    assert np.array_equal(got, exp)
    assert got.sum() > 0


def test_bloom_hash_stage_and_long_sequence(gpu, oracle, tmp_path):
    rng = np.random.default_rng(17)
    p, g, _ = _mk_bloom(oracle, tmp_path, rng, 21)
    bf = gpu.BloomFilter(p, 21)
    a = synth.mutate(rng, g[:3000], n_rate=0.003, lower=0.01)
    hs = bf.hashes(a)
    raw = a.tobytes()
    for i in range(0, len(raw) - 20, 7):
        assert int(hs[i]) == oracle.xxh3_64(oracle.bloom_term(raw[i:i + 21]))
    assert bf.hits(g) == g.size - 20          # every training k-mer is a member
    orc = oracle.BloomOracle(p, 21)
    long_seq = np.concatenate([g[5000:15000], synth.random_dna(rng, 30000)])
    assert bf.hits(long_seq) == orc.hits(long_seq)
    # golden G5 shape (tests/test_probabilistic_single_filter_model.py:42-56): a 22-mer of the training
    # genome has 2 member k-mers
    assert bf.hits(g[777:799]) == 2
    # many tiny sequences incl. shorter than k
    parts = [g[i:i + 21 + (i % 3) - 1] for i in range(0, 4000, 2)]
    bases = np.concatenate(parts)
    lens = np.array([x.size for x in parts], np.uint64)
    e = np.cumsum(lens, dtype=np.uint64)
    assert np.array_equal(bf.query(bases, e - lens, e), orc.hits_batch(bases, e - lens, e))


def _check_bloom_bucketed(gpu, oracle, path, k, bases, b, e, step=1, scratch=0):
    bf = gpu.BloomFilter(path, k)
    shift = 3
    while ((int(bf.info.n_bits) - 1) >> shift) + 1 > 250:
        shift += 1
    exp = oracle.BloomOracle(path, k).hits_batch(bases, b, e, step, threads=4)
    # member_pct 0: always the bucketed kernels; 100 / 35: the device-side choice sends the batch (or not) through k_bloom
    for pct in (0, 100, 35):
        bf.set_bucketed(True, min_windows=1, scratch_bytes=scratch, bucket_shift=shift, member_pct=pct)
        n0 = bf.bucketed_queries
        got = np.asarray(bf.query(bases, b, e, step)).copy()
        assert bf.bucketed_queries > n0, "the bucketed Bloom path was not taken"
        bad = np.flatnonzero(got != exp)
        assert bad.size == 0, f"member_pct {pct}: {bad.size} mismatches, first {bad[:5].tolist()}: got {got[bad[:5]]} exp {exp[bad[:5]]}"
    bf.set_bucketed(False)
    n0 = bf.bucketed_queries
    assert np.array_equal(bf.query(bases, b, e, step), got) and bf.bucketed_queries == n0
    bf.close()
    return got


@pytest.mark.parametrize("k,step", [(21, 1), (31, 1), (17, 2), (21, 500)])
def test_bloom_bucketed_reads(gpu, oracle, tmp_path, k, step):
    """Ragged reads with N / lower case / IUPAC (literal-byte terms) through k_bbucket_emit / fetch / reduce."""
    rng = np.random.default_rng(500 + k)
    p, g, meta = _mk_bloom(oracle, tmp_path, rng, k)
    bases, b, e = synth.sample_reads(rng, [g], 800, (k, 300), sub=0.01, n_rate=0.004, lower=0.002, iupac=0.002)
    got = _check_bloom_bucketed(gpu, oracle, p, k, bases, b, e, step)
    assert got.sum() > 0


def test_bloom_bucketed_long_low_complexity_and_sub_batches(gpu, oracle, tmp_path):
    """Sequences spanning many chunks, homopolymer runs that overflow blocks (direct-probe windows), runs of empty and
    short records, overlapping segments, and a scratch budget of a few chunks (many sub-batches)."""
    rng = np.random.default_rng(31)
    p, g, _ = _mk_bloom(oracle, tmp_path, rng, 21)
    rep = np.tile(np.frombuffer(b"ACGTTGCA", np.uint8), 900)
    parts = [g, np.full(6000, ord("A"), np.uint8), g[:0], g[:20], g[:21], rep] + [g[:5]] * 2500 + [g[100:250], np.full(300, ord("N"), np.uint8),
             g[:5000] | 0x20, synth.random_dna(rng, 30000), g[::-1].copy()]
    bases = np.concatenate(parts)
    lens = np.array([x.size for x in parts], np.uint64)
    e = np.cumsum(lens, dtype=np.uint64)
    got = _check_bloom_bucketed(gpu, oracle, p, 21, bases, e - lens, e)
    assert got[0] == g.size - 20
    _check_bloom_bucketed(gpu, oracle, p, 21, bases, e - lens, e, scratch=1 << 20)
    n = 400
    bb = rng.integers(0, g.size - 600, n).astype(np.uint64)
    ee = bb + rng.integers(0, 600, n).astype(np.uint64)
    _check_bloom_bucketed(gpu, oracle, p, 21, g, bb, ee, scratch=1 << 20)     # windows beyond the bound: tail launch of k_bloom


def test_single_record_shims(gpu, oracle, tmp_path):
    rng = np.random.default_rng(18)
    p, docs = _mk_classic(oracle, tmp_path, rng, 5, 21, 7)
    s = gpu.Search(str(p), True)
    q = docs["doc00002"][0][100:400].tobytes().decode()
    res = s.search(q, step=2)
    exp = oracle.CobsOracle(p).search(q, step=2)
    assert [(r.doc_name, r.score) for r in res] == [(r.doc_name, r.score) for r in exp]
    pb, g, _ = _mk_bloom(oracle, tmp_path, rng, 21)
    bf = gpu.Bloom.load(str(pb), 21)
    km = oracle.bloom_term(g[50:71].tobytes()).decode()
    assert km in bf
    # `in` hashes the k-mer as given, canonical or not (rbloom does not canonicalise)
    orc_bf = oracle.BloomOracle(pb, 21)
    probes = [g[i:i + 21].tobytes().decode() for i in range(0, 2000, 37)] + ["TTGCAACCAGACGTAATCTCT", "acgtnACGTNacgtnACGTNa"]
    assert [p in bf for p in probes] == [p in orc_bf for p in probes]
    assert np.array_equal(bf.filter.contains(probes), np.array([p in orc_bf for p in probes]))
    assert (oracle.bloom_term(b"ACGTTGCATGCATGCATGCAA").decode() in bf) == (
        oracle.bloom_term(b"ACGTTGCATGCATGCATGCAA").decode() in oracle.BloomOracle(pb, 21))


def test_concurrent_queries_on_one_handle(gpu, oracle, tmp_path):
    """Handles are immutable after open: several host threads may query the same index / filter at once
    (include/xspect_b200.h conventions; the reference's web app runs background tasks in threads, web.py:73-90)."""
    import threading
    rng = np.random.default_rng(31)
    p, docs = _mk_classic(oracle, tmp_path, rng, 90, 21, 7, length=2000)
    genomes = [s for v in docs.values() for s in v]
    pb, g, _ = _mk_bloom(oracle, tmp_path, rng, 21)
    ix = gpu.CobsIndex(p)
    bf = gpu.BloomFilter(pb, 21)
    jobs = []
    for t in range(6):
        bases, b, e = synth.sample_reads(np.random.default_rng(100 + t), genomes + [g], 3000, (21, 300), n_rate=0.002)
        jobs.append((bases, b, e))
    expected = [(oracle.CobsOracle(p).counts_batch(*j, 1, threads=4), oracle.BloomOracle(pb, 21).hits_batch(*j, 1, threads=4)) for j in jobs]
    got = [None] * len(jobs)
    errors = []

    def work(i):
        try:
            for _ in range(3):
                got[i] = (np.asarray(ix.query(*jobs[i], 1)).astype(np.uint32), np.asarray(bf.query(*jobs[i], 1)).copy())
        except Exception as exc:   # noqa: BLE001
            errors.append(exc)

    threads = [threading.Thread(target=work, args=(i,)) for i in range(len(jobs))]
    [t.start() for t in threads]
    [t.join() for t in threads]
    assert not errors, errors
    for (gc, gb), (ec, eb) in zip(got, expected):
        assert np.array_equal(gc, ec) and np.array_equal(gb, eb)


# ------------------------------------------------------------------------------------------ config 5 building blocks
def test_synthetic_index_ld_rows_and_sharded_epilogue(gpu, oracle):
    """BASELINE config 5 on one GPU: the synthetic index (rows from the counter-based generator, regenerated by the
    oracle), column shards opened side by side, score tiles written at a common padded row length
    (xs_cobs_query_device_ld), laid out [shard][record][w] as ncclAllGather delivers them, and the in-place epilogue
    (xs_sharded_reduce_device): first best document, its count, tie multiplicity, totals."""
    import torch
    from xspect2_b200 import distributed as xd
    from xspect2_b200._abi import XS_U8, XS_U16
    D, S, K, H, seed = 1000, 20011, 21, 3, 6
    orc = oracle.SynthCobsOracle(D, S, K, H, seed)
    rng = np.random.default_rng(12)
    genome = synth.random_dna(rng, 5000)
    bases, b, e = synth.sample_reads(rng, [genome], 700, (21, 400), n_rate=0.003)
    exp = orc.counts_batch(bases, b, e, 1, threads=4)
    assert exp.max() > 3                       # fill 0.25 ** 3 over 1000 documents: real ties and real maxima
    whole = gpu.CobsIndex.synthetic(D, S, K, H, seed)
    assert whole.n_docs == D and whole.header_layout.startswith("synthetic")
    assert np.array_equal(whole.query(bases, b, e, 1, dtype=4), exp)
    assert abs(whole.doc_fill().mean() - 0.25) < 0.01
    dev = torch.device("cuda", 0)
    d_bases = torch.from_numpy(bases).to(dev)
    d_b = torch.from_numpy(b.view(np.int64)).to(dev)
    d_e = torch.from_numpy(e.view(np.int64)).to(dev)
    n = b.size
    for world in (1, 3, 8):
        shards = xd.column_shards(D, world)
        widths = [hi - lo for lo, hi in shards]
        for dt, tdt, mx in ((XS_U8, torch.uint8, 255), (XS_U16, torch.uint16, 65535)):
            unit = 16 // dt
            w = -(-max(widths) // unit) * unit
            al = torch.full((world, n, w), 7, dtype=tdt, device=dev)          # stale values must be overwritten
            for g, (lo, hi) in enumerate(shards):
                sh = gpu.CobsIndex.synthetic(D, S, K, H, seed, doc_begin=lo, doc_end=hi)
                sh.query_device(d_bases.data_ptr(), bases.size, d_b.data_ptr(), d_e.data_ptr(), n, 1, dt, al[g].data_ptr(), 0, ld=w)
                torch.cuda.synchronize()
                sh.close()
            got = torch.cat([al[g, :, : widths[g]] for g in range(world)], dim=1).cpu().numpy()
            ref = np.minimum(exp, mx)
            assert np.array_equal(got, ref), (world, dt)
            assert int(al[0, :, widths[0]:].to(torch.int64).sum()) == 0       # padding columns are zero
            best = torch.empty(n, dtype=torch.int32, device=dev)
            cnt = torch.empty(n, dtype=torch.int32, device=dev)
            nb = torch.empty(n, dtype=torch.int32, device=dev)
            tot = torch.zeros(D, dtype=torch.int64, device=dev)
            gpu.sharded_reduce_device(al.data_ptr(), n, dt, 0, w, widths, best.data_ptr(), cnt.data_ptr(), nb.data_ptr(), tot.data_ptr(), 0)
            torch.cuda.synchronize()
            assert np.array_equal(best.cpu().numpy(), ref.argmax(axis=1)), (world, dt)
            assert np.array_equal(cnt.cpu().numpy(), ref.max(axis=1))
            assert np.array_equal(nb.cpu().numpy(), (ref == ref.max(axis=1)[:, None]).sum(axis=1))
            assert np.array_equal(tot.cpu().numpy(), ref.sum(axis=0, dtype=np.int64))
    whole.close()
