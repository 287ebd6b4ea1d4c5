"""Direct-gather kernel vs the bucketed kernels on the bench workload (device-resident reads), per-kernel times,
and the sensitivity to the scratch budget (= windows per sub-batch = how often every index row is re-used from L2).

    python profiles/experiments/bucketed_phases.py [scratch_GiB[:prefetch] ...]

prefetch = 0 switches the next-slice L2 prefetch of k_bucket_fetch off.  (Earlier revisions of this script also drove
multi-stream variants — emit / fetch / reduce of neighbouring sub-batches on three streams, and fetch(i) next to
emit(i+1) — which were measured and removed: profiles/r1_bucketed_notes.md.)
"""
import os
import json
import sys
import tempfile
import time
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
import bench  # noqa: E402
from xspect2_b200 import engine  # noqa: E402
from xspect2_b200._abi import XS_U8  # noqa: E402
from xspect2_b200.synth import fixed_offsets  # noqa: E402


def main():
    configs = sys.argv[1:] or ["24"]
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    wd = Path(tempfile.mkdtemp(prefix="xs_bk_"))
    path, d_bases = bench.build_workload(wd, dev, lambda g: engine.kmer_rows(g, bench.K, bench.H, bench.SIG_SIZE, device=0))
    ix = engine.CobsIndex(path)
    N, L, D = bench.N_READS, bench.READ_LEN, bench.D
    hb, he = fixed_offsets(N, L)
    d_b = torch.from_numpy(hb.view(np.int64)).to(dev)
    d_e = torch.from_numpy(he.view(np.int64)).to(dev)
    out_a = torch.empty((N, D), dtype=torch.uint8, device=dev)
    out_b = torch.empty((N, D), dtype=torch.uint8, device=dev)
    stream = torch.cuda.current_stream()
    lookups = N * (L - bench.K + 1)

    def run(out, steps=3):
        for _ in range(2):
            ix.query_device(d_bases.data_ptr(), N * L, d_b.data_ptr(), d_e.data_ptr(), N, 1, XS_U8, out.data_ptr(), stream.cuda_stream)
        torch.cuda.synchronize()
        engine.profile_enable(True)
        engine.profile_read()
        t0 = time.perf_counter()
        for _ in range(steps):
            ix.query_device(d_bases.data_ptr(), N * L, d_b.data_ptr(), d_e.data_ptr(), N, 1, XS_U8, out.data_ptr(), stream.cuda_stream)
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / steps
        ms, n = engine.profile_read_phases()
        engine.profile_enable(False)
        return dt, [m / steps for m in ms], n

    ix.set_bucketed(False)
    dt, ms, n = run(out_a)
    print(json.dumps({"path": "direct", "ms_per_step": dt * 1e3, "G_lookups_s": lookups / dt / 1e9, "kernel_ms": ms, "launches": n}), flush=True)
    names = ["XS_BK_PREFETCH"]
    for cfg in configs:
        f = cfg.split(":")
        gib = float(f[0])
        for name, v in zip(names, f[1:] + [""]):
            if v:
                os.environ[name] = v
            else:
                os.environ.pop(name, None)
        ix.set_bucketed(True, min_windows=1 << 20, scratch_bytes=int(gib * (1 << 30)))
        out_b.zero_()
        dt, ms, n = run(out_b)
        same = bool(torch.equal(out_a, out_b))
        print(json.dumps({"path": "bucketed", "config": cfg, "ms_per_step": dt * 1e3, "G_lookups_s": lookups / dt / 1e9,
                          "kernel_ms[direct,emit,fetch,reduce]": ms, "launches": n, "identical_to_direct": same,
                          "mem_GB": torch.cuda.mem_get_info()[0] / 1e9}), flush=True)


if __name__ == "__main__":
    main()
