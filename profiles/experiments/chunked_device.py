"""Experiment: device-resident query of the bench workload issued as C chunks round-robin over S streams."""
import os, sys, time, tempfile
from pathlib import Path
import numpy as np, torch
sys.path.insert(0, str(Path(__file__).resolve().parents[2]))
import bench
from xspect2_b200 import engine
from xspect2_b200._abi import XS_U8
from xspect2_b200.synth import fixed_offsets

dev = torch.device("cuda", 0)
wd = Path(tempfile.mkdtemp())
path, d_bases = bench.build_workload(wd, dev, lambda g: engine.kmer_rows(g, bench.K, bench.H, bench.SIG_SIZE))
ix = engine.CobsIndex(path)
N, L, D = bench.N_READS, bench.READ_LEN, bench.D
hb, he = fixed_offsets(N, L)
d_out = torch.empty((N, D), dtype=torch.uint8, device=dev)
for chunks, nstream in [(1, 1), (4, 1), (4, 2), (16, 1), (16, 3), (64, 3), (64, 8), (256, 4)]:
    per = N // chunks
    streams = [torch.cuda.Stream() for _ in range(nstream)]
    lb = torch.from_numpy((np.arange(per, dtype=np.uint64) * L).view(np.int64)).to(dev)
    le = lb + L
    def run():
        for c in range(chunks):
            s = streams[c % nstream]
            ix.query_device(d_bases.data_ptr() + c * per * L, per * L, lb.data_ptr(), le.data_ptr(), per, 1, XS_U8,
                            d_out.data_ptr() + c * per * D, s.cuda_stream)
    run(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(3):
        run()
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / 3
    print(f"chunks={chunks:4d} streams={nstream}: {dt*1e3:8.2f} ms  {N*130/dt/1e9:6.2f} G lookups/s", flush=True)
