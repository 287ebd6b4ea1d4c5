"""Seeded synthetic genomes, reads and model files for the tests (test infrastructure; uses the oracle's
format-faithful writers because the reference's own fixtures are NCBI downloads, tests/conftest.py:12-48)."""
from __future__ import annotations

import numpy as np

ACGT = np.frombuffer(b"ACGT", dtype=np.uint8)


def random_dna(rng: np.random.Generator, n: int) -> np.ndarray:
    return ACGT[rng.integers(0, 4, size=n)]


def revcomp(a: np.ndarray) -> np.ndarray:
    lut = np.arange(256, dtype=np.uint8)
    for x, y in zip(b"ACGTacgt", b"TGCAtgca"):
        lut[x] = y
    return lut[a[::-1]]


def mutate(rng, a: np.ndarray, sub: float = 0.0, n_rate: float = 0.0, lower: float = 0.0, iupac: float = 0.0) -> np.ndarray:
    a = a.copy()
    n = a.size
    if sub:
        m = rng.random(n) < sub
        a[m] = ACGT[rng.integers(0, 4, size=int(m.sum()))]
    if n_rate:
        a[rng.random(n) < n_rate] = ord("N")
    if iupac:
        m = rng.random(n) < iupac
        a[m] = np.frombuffer(b"RYKMSWBDHVnUu-*", dtype=np.uint8)[rng.integers(0, 15, size=int(m.sum()))]
    if lower:
        m = rng.random(n) < lower
        a[m] = a[m] | 0x20
    return a


def make_genomes(rng, n_docs: int, length: int, shared: float = 0.3) -> dict[str, list[np.ndarray]]:
    """n_docs genomes sharing a fraction of an ancestral sequence, so documents overlap in k-mers."""
    anc = random_dna(rng, length)
    docs = {}
    for d in range(n_docs):
        g = random_dna(rng, length)
        m = rng.random(length // 100 + 1) < shared           # share in blocks of 100 bases
        mask = np.repeat(m, 100)[:length]
        g[mask] = anc[mask]
        docs[f"doc{d:05d}"] = [g]
    return docs


def sample_reads(rng, genomes: list[np.ndarray], n_reads: int, read_len, frac_random: float = 0.3, **mut):
    """Concatenated reads + offsets. read_len: int or (lo, hi) for ragged reads."""
    parts, lens = [], []
    for _ in range(n_reads):
        L = int(read_len) if np.isscalar(read_len) else int(rng.integers(read_len[0], read_len[1] + 1))
        if rng.random() < frac_random or not genomes:
            r = random_dna(rng, L)
        else:
            g = genomes[int(rng.integers(0, len(genomes)))]
            if g.size <= L:
                r = random_dna(rng, L)
            else:
                s = int(rng.integers(0, g.size - L))
                r = g[s:s + L]
                if rng.random() < 0.5:
                    r = revcomp(r)
        parts.append(mutate(rng, r, **mut) if mut else r)
        lens.append(L)
    bases = np.concatenate(parts) if parts else np.zeros(0, np.uint8)
    ends = np.cumsum(np.array(lens, dtype=np.uint64), dtype=np.uint64)
    begins = ends - np.array(lens, dtype=np.uint64)
    return bases, begins, ends
