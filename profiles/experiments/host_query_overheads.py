import sys, time, numpy as np, torch, tempfile
sys.path.insert(0, '/root/repo')
from pathlib import Path
from oracle import oracle
from tests import model_fixtures as mf
from xspect2_b200 import engine, synth
from xspect2_b200._abi import XS_U8
rng = np.random.default_rng(9)
td = Path(tempfile.mkdtemp())
sp_json, genomes, _ = mf.species_model(oracle, td, rng, n_species=90, genome_len=20000, svm=False)
ix = engine.CobsIndex(td / "testgenus-species" / "index.cobs_classic")
n, L = 2_000_000, 150
g = np.concatenate(list(genomes.values()))
reads = synth.synth_reads(g, n, L, seed=10, device="cuda:0").cpu().numpy()
hb, he = synth.fixed_offsets(n, L)
pb = engine.pinned_empty(n * L, np.uint8); pb[:] = reads
out = engine.pinned_empty((n, 90), np.uint8)
for name, bases, o in (("pageable in, pinned out", reads, out), ("pinned in/out", pb, out), ("pinned in, alloc out", pb, None)):
    for rep in range(3):
        t = time.perf_counter(); r = ix.query(bases, hb, he, 1, XS_U8, out=o); dt = time.perf_counter() - t
        print(name, rep, round(dt * 1e3, 1), "ms", flush=True)
t = time.perf_counter(); mx = int((he - hb).max()); print("max", (time.perf_counter() - t) * 1e3)
