"""Latency of the single-record shims (one C-ABI call per record, like the reference's loop)."""
import sys, time, tempfile
from pathlib import Path
import numpy as np
sys.path.insert(0, str(Path(__file__).resolve().parents[2]))
from oracle import oracle
from tests import model_fixtures as mf, synth
from xspect2_b200 import engine
rng = np.random.default_rng(1)
td = Path(tempfile.mkdtemp())
sp_json, genomes, _ = mf.species_model(oracle, td, rng, n_species=90, genome_len=20000, svm=False)
s = engine.Search(str(td / "testgenus-species" / "index.cobs_classic"))
g = next(iter(genomes.values()))
reads = [g[i:i + 150].tobytes().decode() for i in range(0, 15000, 15)]
for r in reads[:50]:
    s.search(r)
t = time.perf_counter()
for r in reads:
    s.search(r)
dt = (time.perf_counter() - t) / len(reads)
print(f"Search.search(150 bp): {dt * 1e6:.1f} us per call")
t = time.perf_counter()
for r in reads:
    s.index.counts(r)
print(f"CobsIndex.counts(150 bp): {(time.perf_counter() - t) / len(reads) * 1e6:.1f} us per call")
orc = oracle.CobsOracle(td / "testgenus-species" / "index.cobs_classic")
t = time.perf_counter()
for r in reads:
    orc.search(r)
print(f"oracle search (CPU, 1 thread): {(time.perf_counter() - t) / len(reads) * 1e6:.1f} us per call")
contig = g.tobytes().decode()
t = time.perf_counter()
for _ in range(20):
    s.search(contig)
print(f"Search.search(20 kbp): {(time.perf_counter() - t) / 20 * 1e6:.1f} us per call")
