"""Synthetic workloads of BASELINE.json's shapes (bench.py, smoke): model files in the reference's
own on-disk formats and Illumina-style read batches.  There is no network for real models or datasets, so
`index.cobs_classic` / `filter.bloom` files are synthesised with the geometry of the trained models
 and then loaded through the normal file path, unchanged.

torch is used as a random-number and array engine only (on the GPU when there is one).
"""

from __future__ import annotations

import struct
from pathlib import Path

import numpy as np
import torch

ACGT = np.frombuffer(b"ACGT", dtype=np.uint8)


def classic_header(k: int, canonicalize: int, names: list[str], sig_size: int, num_hashes: int) -> bytes:
    """COBS classic index header: magic, version 1, term size, canonicalize, document count,
    signature size, hash count, one document name per line, magic."""
    h = b"COBS:CLASSIC_INDEX" + struct.pack("<II", 1, k) + struct.pack("<B", canonicalize)
    h += struct.pack("<I", len(names)) + struct.pack("<QQ", sig_size, num_hashes)
    h += b"".join(n.encode() + b"\n" for n in names)
    return h + b"CLASSIC_INDEX"


def synth_genome(length: int, seed: int, n_rate: float = 0.0) -> np.ndarray:
    rng = np.random.default_rng(seed)
    g = ACGT[rng.integers(0, 4, size=length)]
    if n_rate:
        n_runs = max(1, int(length * n_rate / 50))
        for s in rng.integers(0, length - 100, size=n_runs):
            g[s : s + int(rng.integers(1, 101))] = ord("N")
    return g


def _dev(device) -> torch.device:
    if device is None:
        return torch.device("cuda", torch.cuda.current_device()) if torch.cuda.is_available() else torch.device("cpu")
    return torch.device(device)


def write_classic_index(path, *, n_docs: int, k: int, num_hashes: int, sig_size: int, seed: int,
                        plant: dict[int, np.ndarray] | None = None, fill=(0.25, 0.52), device=None,
                        names: list[str] | None = None, rows_chunk: int = 1 << 22) -> list[str]:
    """Write ``index.cobs_classic`` with iid Bernoulli(p_d) bits per document d (p_d spread over ``fill``)
    and, for ``plant = {doc: row ids}``, the bits of those rows set in document ``doc`` (true positives)."""
    dev = _dev(device)
    names = names or [f"{1000 + d}" for d in range(n_docs)]
    row = (n_docs + 7) // 8
    gen = torch.Generator(device=dev).manual_seed(seed)
    p = torch.zeros(row * 8, device=dev)
    p[:n_docs] = torch.linspace(fill[0], fill[1], n_docs, device=dev)[torch.randperm(n_docs, generator=gen, device=dev)]
    weights = (2 ** torch.arange(8, device=dev)).to(torch.int32)
    plant_t = {d: torch.from_numpy(np.unique(np.asarray(r, dtype=np.uint64)).astype(np.int64)).to(dev)
               for d, r in (plant or {}).items()}
    with open(path, "wb") as f:
        f.write(classic_header(k, 1, names, sig_size, num_hashes))
        for r0 in range(0, sig_size, rows_chunk):
            n = min(rows_chunk, sig_size - r0)
            bits = (torch.rand((n, row * 8), generator=gen, device=dev) < p).view(n, row, 8).to(torch.int32)
            data = (bits * weights).sum(dim=2).to(torch.uint8)
            for d, rws in plant_t.items():
                sel = rws[(rws >= r0) & (rws < r0 + n)] - r0
                data[sel, d // 8] |= 1 << (d % 8)
            f.write(data.cpu().numpy().tobytes())
    return names


def synth_reads(genome: np.ndarray, n_reads: int, read_len: int, seed: int, *, frac_genome: float = 0.6,
                sub_rate: float = 0.001, n_rate: float = 0.0005, device=None, chunk: int = 1 << 20) -> torch.Tensor:
    """``n_reads`` reads of ``read_len`` bases, concatenated (uint8 ASCII) on ``device``: ``frac_genome`` sampled
    from ``genome`` (either strand), the rest uniform random; substitution errors and stray N."""
    dev = _dev(device)
    gen = torch.Generator(device=dev).manual_seed(seed)
    g = torch.from_numpy(np.ascontiguousarray(genome)).to(dev)
    lut = torch.arange(256, dtype=torch.uint8, device=dev)
    for a, b in zip(b"ACGT", b"TGCA"):
        lut[a] = b
    acgt = torch.from_numpy(ACGT.copy()).to(dev)
    out = torch.empty(n_reads * read_len, dtype=torch.uint8, device=dev)
    ar = torch.arange(read_len, device=dev)
    for r0 in range(0, n_reads, chunk):
        c = min(chunk, n_reads - r0)
        pos = torch.randint(0, max(1, g.numel() - read_len), (c,), generator=gen, device=dev)
        r = g[pos[:, None] + ar[None, :]]
        strand = torch.rand(c, generator=gen, device=dev) < 0.5
        r = torch.where(strand[:, None], lut[r.flip(1).long()], r)
        rnd = acgt[torch.randint(0, 4, (c, read_len), generator=gen, device=dev)]
        from_g = torch.rand(c, generator=gen, device=dev) < frac_genome
        r = torch.where(from_g[:, None], r, rnd)
        sub = torch.rand((c, read_len), generator=gen, device=dev) < sub_rate
        r = torch.where(sub, acgt[torch.randint(0, 4, (c, read_len), generator=gen, device=dev)], r)
        nm = torch.rand((c, read_len), generator=gen, device=dev) < n_rate
        r = torch.where(nm, torch.full_like(r, ord("N")), r)
        out[r0 * read_len : (r0 + c) * read_len] = r.reshape(-1)
    return out


def fixed_offsets(n_reads: int, read_len: int) -> tuple[np.ndarray, np.ndarray]:
    b = np.arange(n_reads, dtype=np.uint64) * np.uint64(read_len)
    return b, b + np.uint64(read_len)
