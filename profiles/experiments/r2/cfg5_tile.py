"""One GPU, config-5 geometry at a reduced row count: a few read tiles through k_cobs_wide + k_sharded_reduce, for an
ncu capture of those two kernels (python profiles/experiments/r2/cfg5_tile.py [rows] [docs_lo docs_hi])."""
import sys
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[3]))
from xspect2_b200 import engine, synth  # noqa: E402
from xspect2_b200._abi import XS_U8  # noqa: E402

S = int(sys.argv[1]) if len(sys.argv) > 1 else 8_000_000
lo, hi = (int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (0, 10_000)
D, TILE, L = 10_000, 100_000, 150
dev = torch.device("cuda", 0)
ix = engine.CobsIndex.synthetic(D, S, 21, 7, 6, doc_begin=lo, doc_end=hi)
genome = synth.synth_genome(1_000_000, seed=7)
reads = synth.synth_reads(genome, 4 * TILE, L, seed=7, device=dev)
hb, he = synth.fixed_offsets(TILE, L)
d_b = torch.from_numpy(hb.view(np.int64)).to(dev)
d_e = torch.from_numpy(he.view(np.int64)).to(dev)
w = -(-(hi - lo) // 16) * 16
local = torch.empty((TILE, w), dtype=torch.uint8, device=dev)
best = torch.empty(TILE, dtype=torch.int32, device=dev)
cnt = torch.empty(TILE, dtype=torch.int32, device=dev)
nb = torch.empty(TILE, dtype=torch.int32, device=dev)
s = torch.cuda.current_stream().cuda_stream
for t in range(4):
    ix.query_device(reads.data_ptr() + t * TILE * L, TILE * L, d_b.data_ptr(), d_e.data_ptr(), TILE, 1, XS_U8, local.data_ptr(), s, ld=w)
    engine.sharded_reduce_device(local.data_ptr(), TILE, XS_U8, 0, w, [hi - lo], best.data_ptr(), cnt.data_ptr(), nb.data_ptr(), 0, s)
torch.cuda.synchronize()
print("ok", int(best.sum()), ix.info.row_stride)
