"""Size-independent properties of the CUDA path at sizes the oracle would take minutes for (1 M reads against a
20 M-row index; the index is synthesised like bench.py's).  No oracle involved: the properties follow from the
definition of the score (a sum over sampled windows of the AND of h rows selected by the *canonical* k-mer)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

K, H, D, S = 21, 7, 90, 20_000_003
N_READS, L = 1_000_000, 150


@pytest.fixture(scope="module")
def big(tmp_path_factory, gpu):
    from xspect2_b200 import synth
    td = tmp_path_factory.mktemp("big")
    dev = torch.device("cuda", 0)
    genome = synth.synth_genome(1_000_000, seed=1)
    rows, valid = gpu.kmer_rows(genome, K, H, S)
    path = td / "index.cobs_classic"
    synth.write_classic_index(path, n_docs=D, k=K, num_hashes=H, sig_size=S, seed=2, plant={0: rows[valid.astype(bool)].reshape(-1)}, device=dev)
    reads = synth.synth_reads(genome, N_READS, L, seed=3, n_rate=0.0, device=dev).cpu().numpy()
    b, e = synth.fixed_offsets(N_READS, L)
    ix = gpu.CobsIndex(path)
    return dict(ix=ix, path=path, reads=reads, b=b, e=e, genome=genome)


def _revcomp_reads(reads: np.ndarray) -> np.ndarray:
    lut = np.arange(256, dtype=np.uint8)
    for x, y in zip(b"ACGT", b"TGCA"):
        lut[x] = y
    return np.ascontiguousarray(lut[reads.reshape(N_READS, L)[:, ::-1]]).reshape(-1)


def test_reverse_complement_invariance(big):
    """A read and its reverse complement have the same canonical k-mers, hence identical counts."""
    ix = big["ix"]
    fwd = np.asarray(ix.query(big["reads"], big["b"], big["e"], 1)).copy()
    rc = np.asarray(ix.query(_revcomp_reads(big["reads"]), big["b"], big["e"], 1))
    assert fwd.shape == (N_READS, D) and np.array_equal(fwd, rc)
    assert int(fwd[:, 0].max()) == L - K + 1          # reads drawn from the planted genome hit document 0 on every window
    big["fwd"] = fwd


def test_bucketed_and_direct_kernels_agree(big):
    """The 1 M-read batch is large enough for the bucketed kernels (probe records grouped by L2-sized row ranges);
    the direct-gather kernel must give the same matrix, and so must a run with a scratch budget of a few hundred
    chunks (many sub-batches) and one with uint32 counts."""
    ix = big["ix"]
    fwd = big.get("fwd")
    n0 = ix.bucketed_queries
    if fwd is None:
        fwd = np.asarray(ix.query(big["reads"], big["b"], big["e"], 1)).copy()
        big["fwd"] = fwd
    assert ix.bucketed_queries > 0, "large batches should take the bucketed path by default"
    ix.set_bucketed(False)
    n0 = ix.bucketed_queries
    direct = np.asarray(ix.query(big["reads"], big["b"], big["e"], 1))
    assert ix.bucketed_queries == n0
    assert np.array_equal(direct, fwd)
    ix.set_bucketed(True, min_windows=1 << 20, scratch_bytes=256 << 20)
    n = 400_000
    small = np.asarray(ix.query(big["reads"][: n * L], big["b"][:n], big["e"][:n], 1, dtype=4))
    assert ix.bucketed_queries > n0
    assert np.array_equal(small, fwd[:n])
    ix.set_bucketed(True, min_windows=32 << 20, scratch_bytes=24 << 30)


def test_additivity_over_overlapping_chunks_and_steps(big):
    """Counts are additive over windows: a long sequence equals the sum of its chunks overlapping by k-1 bases
    (what the MLST splitter relies on), and step-s sampling at the s offsets partitions the step-1 windows."""
    ix = big["ix"]
    g = big["genome"][:600_000]
    whole = ix.counts(g, 1).astype(np.int64)
    chunk = 4999
    starts = np.arange(0, g.size - K + 1, chunk - K + 1, dtype=np.uint64)
    ends = np.minimum(starts + chunk, g.size).astype(np.uint64)
    parts = np.asarray(ix.query(g, starts, ends, 1)).astype(np.int64)
    assert np.array_equal(parts.sum(axis=0), whole)
    assert whole[0] == g.size - K + 1
    for s in (2, 7):
        by_offset = sum(ix.counts(g[o:], s).astype(np.int64) for o in range(s))
        assert np.array_equal(by_offset, whole)


def test_totals_epilogue_shards_and_dtypes_agree(big, gpu):
    ix = big["ix"]
    fwd = big.get("fwd")
    if fwd is None:
        fwd = np.asarray(ix.query(big["reads"], big["b"], big["e"], 1)).copy()
    best, cnt, nb, totals = ix.classify(big["reads"], big["b"], big["e"], 1)
    f = fwd.astype(np.int64)
    assert np.array_equal(totals.astype(np.int64), f.sum(axis=0))
    assert np.array_equal(np.asarray(cnt).astype(np.int64), f.max(axis=1)) and np.array_equal(np.asarray(best).astype(np.int64), f.argmax(axis=1))
    u32 = np.asarray(ix.query(big["reads"][: 200_000 * L], big["b"][:200_000], big["e"][:200_000], 1, dtype=4))
    assert np.array_equal(u32, fwd[:200_000])
    lo = gpu.CobsIndex(big["path"], doc_begin=0, doc_end=48)
    hi = gpu.CobsIndex(big["path"], doc_begin=48, doc_end=90)
    n = 300_000
    joined = np.concatenate([np.asarray(lo.query(big["reads"][: n * L], big["b"][:n], big["e"][:n], 1)),
                             np.asarray(hi.query(big["reads"][: n * L], big["b"][:n], big["e"][:n], 1))], axis=1)
    assert np.array_equal(joined, fwd[:n])


def test_one_very_long_sequence_through_the_host_pipeline(big):
    """A single 70 Mbp record (larger than one pipeline batch) equals the sum of its two overlapping halves, and a
    mixed batch with records before and after it keeps every row where it belongs."""
    ix = big["ix"]
    rng = np.random.default_rng(5)
    n = 70_000_000
    seq = np.frombuffer(b"ACGT", np.uint8)[rng.integers(0, 4, size=n)]
    seq[10_000_000:10_600_000] = big["genome"][:600_000]              # a stretch of the planted genome
    whole = ix.counts(seq, 1).astype(np.int64)
    half = n // 2
    a = ix.counts(seq[: half + K - 1], 1).astype(np.int64)
    b = ix.counts(seq[half:], 1).astype(np.int64)
    assert np.array_equal(a + b, whole) and whole[0] >= 600_000 - K + 1
    small = big["reads"][: 1000 * L]
    bases = np.concatenate([small, seq, small])
    begin = np.concatenate([big["b"][:1000], [1000 * L], big["b"][:1000] + np.uint64(1000 * L + n)]).astype(np.uint64)
    end = np.concatenate([big["e"][:1000], [1000 * L + n], big["e"][:1000] + np.uint64(1000 * L + n)]).astype(np.uint64)
    mixed = np.asarray(ix.query(bases, begin, end, 1, dtype=4)).astype(np.int64)
    assert np.array_equal(mixed[1000], whole)
    assert np.array_equal(mixed[:1000], mixed[1001:])
    ref = np.asarray(ix.query(small, big["b"][:1000], big["e"][:1000], 1, dtype=4)).astype(np.int64)
    assert np.array_equal(mixed[:1000], ref)
