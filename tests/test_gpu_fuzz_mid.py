"""Randomised cross-check of k_cobs_mid (rows at a 32 / 64 / 128-byte stride) against k_cobs_wide on the same file and,
for a subset, against the oracle: random document counts in 129 .. 1024, k, hash counts up to 8, steps, count types,
non-ACGT policies and the ragged inputs of the bucketed fuzz test (reads, contigs over many warp tiles, runs of empty
records, low complexity, overlapping unordered segments).  XS_FUZZ_CASES raises the number of cases."""
import os

import numpy as np
import pytest

from tests import synth
from tests.test_gpu_fuzz_bucketed import _random_batch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("seed", range(int(os.environ.get("XS_FUZZ_CASES", 24)) // 2))
def test_cobs_mid_equals_wide(gpu, oracle, tmp_path, monkeypatch, seed):
    rng = np.random.default_rng(7000 + seed)
    n_docs = int(rng.choice([129, 130, 200, 255, 256, 257, 300, 511, 512, 513, 777, 1000, 1023, 1024]))
    k = int(rng.choice([13, 21, 31, 32]))
    h = int(rng.choice([1, 2, 3, 5, 6, 7, 8]))
    docs = synth.make_genomes(rng, n_docs, int(rng.integers(100, 500)))
    p = tmp_path / "index.cobs_classic"
    oracle.write_classic(p, docs, k=k, num_hashes=h, fpr=float(rng.choice([0.01, 0.1, 0.3])))
    genomes = [s for v in docs.values() for s in v]
    bases, b, e = _random_batch(rng, genomes, k)
    step = int(rng.choice([1, 1, 2, 5]))
    dtype = [None, 1, 2, 4][int(rng.integers(0, 4))]
    policy = int(rng.integers(0, 2))
    ix = gpu.CobsIndex(p)
    assert ix.kernel == "k_cobs_mid"
    ix.set_policy(policy)
    got = np.asarray(ix.query(bases, b, e, step=step, dtype=dtype)).copy()
    ix.close()
    monkeypatch.setenv("XS_NO_MID_KERNEL", "1")
    wx = gpu.CobsIndex(p)
    assert wx.kernel == "k_cobs_wide"
    wx.set_policy(policy)
    wide = np.asarray(wx.query(bases, b, e, step=step, dtype=dtype))
    assert np.array_equal(got, wide), f"seed {seed}: n_docs {n_docs} k {k} h {h} step {step} dtype {dtype} policy {policy}"
    wx.close()
    if seed % 3 == 0:
        exp = oracle.CobsOracle(p, policy=policy).counts_batch(bases, b, e, step=step, threads=4)
        if dtype in (1, 2):
            exp = np.minimum(exp, 255 if dtype == 1 else 65535)
        assert np.array_equal(got.astype(np.uint32), exp)
