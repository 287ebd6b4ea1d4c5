"""Randomised cross-check of the two kernel paths: for random geometries (documents, k, hashes, step, bucket size,
scratch budget) and ragged inputs (reads, contigs, empty / short records, N, lower case, overlapping segments) the
bucketed kernels must return exactly what the direct-gather kernels return; a subset is also checked against the
oracle.  Both sides run on the GPU, so many cases fit in seconds (XS_FUZZ_CASES raises the number of cases)."""
import os

import numpy as np
import pytest

from tests import synth

pytestmark = pytest.mark.gpu


def _random_batch(rng, genomes, k):
    parts = []
    for _ in range(int(rng.integers(1, 6))):
        kind = int(rng.integers(0, 6))
        if kind == 0:      # ragged reads
            n = int(rng.integers(1, 400))
            b, bb, ee = synth.sample_reads(rng, genomes, n, (max(1, k - 3), int(rng.integers(k + 1, 400))),
                                           sub=0.01, n_rate=float(rng.choice([0.0, 0.002, 0.02])), lower=float(rng.choice([0.0, 0.01])))
            parts += [b[int(x):int(y)] for x, y in zip(bb, ee)]
        elif kind == 1:    # a contig spanning several chunks
            g = genomes[int(rng.integers(0, len(genomes)))]
            reps = int(rng.integers(1, 4))
            parts.append(synth.mutate(rng, np.tile(g, reps), n_rate=0.0005))
        elif kind == 2:    # runs of empty / too short records
            parts += [genomes[0][: int(rng.integers(0, k))]] * int(rng.integers(1, 3000))
        elif kind == 3:    # low complexity
            parts.append(np.tile(np.frombuffer(b"ACGTTG"[: int(rng.integers(1, 7))], np.uint8), int(rng.integers(50, 1500))))
        elif kind == 4:    # exactly k, k + 1
            g = genomes[0]
            parts += [g[:k], g[5:5 + k + 1]]
        else:
            parts.append(synth.random_dna(rng, int(rng.integers(k, 5000))))
    bases = np.concatenate(parts) if parts else np.zeros(0, np.uint8)
    lens = np.array([p.size for p in parts], np.uint64)
    e = np.cumsum(lens, dtype=np.uint64)
    b = e - lens
    if rng.random() < 0.3 and bases.size > 700:      # overlapping, unordered segments on top
        n = int(rng.integers(1, 200))
        ob = rng.integers(0, bases.size - 600, n).astype(np.uint64)
        oe = ob + rng.integers(0, 600, n).astype(np.uint64)
        b, e = np.concatenate([b, ob]), np.concatenate([e, oe])
    return bases, b.astype(np.uint64), e.astype(np.uint64)


@pytest.mark.parametrize("seed", range(int(os.environ.get("XS_FUZZ_CASES", 24))))
def test_cobs_bucketed_equals_direct(gpu, oracle, tmp_path, seed):
    rng = np.random.default_rng(9000 + seed)
    n_docs = int(rng.choice([1, 8, 33, 90, 96, 97, 128]))
    k = int(rng.choice([13, 21, 31, 32]))
    h = int(rng.choice([1, 3, 7]))
    docs = synth.make_genomes(rng, n_docs, int(rng.integers(200, 3000)))
    p = tmp_path / "index.cobs_classic"
    oracle.write_classic(p, docs, k=k, num_hashes=h, fpr=float(rng.choice([0.01, 0.1])))
    genomes = [s for v in docs.values() for s in v]
    bases, b, e = _random_batch(rng, genomes, k)
    step = int(rng.choice([1, 1, 2, 5]))
    dtype = [None, 1, 2, 4][int(rng.integers(0, 4))]
    ix = gpu.CobsIndex(p)
    ix.set_policy(int(rng.integers(0, 2)))
    ix.set_bucketed(False)
    direct = np.asarray(ix.query(bases, b, e, step=step, dtype=dtype)).copy()
    sig = int(ix.info.sig_size_max)
    shift = int(rng.integers(1, 22))
    while ((sig - 1) >> shift) + 1 > 256:
        shift += 1
    scratch = int(rng.choice([0, 1 << 20, 8 << 20]))
    ix.set_bucketed(True, min_windows=1, scratch_bytes=scratch, bucket_shift=shift)
    n0 = ix.bucketed_queries
    got = np.asarray(ix.query(bases, b, e, step=step, dtype=dtype))
    if bases.size // step >= 1 and b.size:
        assert ix.bucketed_queries > n0
    assert np.array_equal(got, direct), f"seed {seed}: n_docs {n_docs} k {k} h {h} step {step} shift {shift} scratch {scratch}"
    if seed % 4 == 0:
        exp = oracle.CobsOracle(p, policy=int(ix.info.policy)).counts_batch(bases, b, e, step=step, threads=4)
        if dtype in (1, 2):
            exp = np.minimum(exp, 255 if dtype == 1 else 65535)
        assert np.array_equal(got.astype(np.uint32), exp)


@pytest.mark.parametrize("seed", range(int(os.environ.get("XS_FUZZ_CASES", 24)) // 2))
def test_bloom_bucketed_equals_direct(gpu, oracle, tmp_path, seed):
    rng = np.random.default_rng(9500 + seed)
    k = int(rng.choice([13, 21, 31]))
    g = synth.random_dna(rng, int(rng.integers(500, 30000)))
    p = tmp_path / "filter.bloom"
    oracle.write_bloom(p, [g], k=k, fpr=float(rng.choice([0.01, 0.05, 0.3])))
    bases, b, e = _random_batch(rng, [g], k)
    step = int(rng.choice([1, 1, 3]))
    bf = gpu.BloomFilter(p, k)
    bf.set_bucketed(False)
    direct = np.asarray(bf.query(bases, b, e, step)).copy()
    shift = int(rng.integers(3, 20))
    while ((int(bf.info.n_bits) - 1) >> shift) + 1 > 256:
        shift += 1
    bf.set_bucketed(True, min_windows=1, scratch_bytes=int(rng.choice([0, 1 << 20])), bucket_shift=shift,
                    member_pct=int(rng.choice([0, 0, 35, 100])))
    got = np.asarray(bf.query(bases, b, e, step))
    assert np.array_equal(got, direct), f"seed {seed}: k {k} step {step} shift {shift}"
    if seed % 4 == 0:
        assert np.array_equal(got, oracle.BloomOracle(p, k).hits_batch(bases, b, e, step, threads=4))
