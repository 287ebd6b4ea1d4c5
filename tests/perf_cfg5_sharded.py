"""Config 5 geometry, document-column sharded over the GPUs of one box (run under torchrun):
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tests/perf_cfg5_sharded.py
D = 10 000 documents, h = 7, k = 21; every rank holds its column range of every row, scores every read tile,
and the tiles are combined (a) by the NCCL all-gather of the score rows, (b) by the reduced exchange of
per-record maxima.  Rank 0 prints one JSON line.  Test infrastructure (parity sample against the oracle)."""
import json
import os
import sys
import time
from pathlib import Path

import numpy as np
import torch
import torch.distributed as dist

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from xspect2_b200 import distributed as xd, engine, synth  # noqa: E402
from xspect2_b200._abi import XS_U8  # noqa: E402

D, H, K = 10_000, 7, 21
S = int(os.environ.get("XS_CFG5_ROWS", 4_000_000))
N_READS = int(os.environ.get("XS_CFG5_READS", 1_000_000))
TILE = 100_000
L = 150


def main():
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    path = Path(os.environ.get("XS_CFG5_DIR", "/tmp")) / f"xs_cfg5_{S}.cobs_classic"
    if rank == 0 and not path.exists():
        gen = torch.Generator(device=dev).manual_seed(6)
        row = D // 8
        with open(path, "wb") as f:
            f.write(synth.classic_header(K, 1, [f"d{i}" for i in range(D)], S, H))
            for r0 in range(0, S, 1 << 19):
                n = min(1 << 19, S - r0)
                a = torch.randint(0, 256, (n, row), generator=gen, device=dev, dtype=torch.uint8)
                b = torch.randint(0, 256, (n, row), generator=gen, device=dev, dtype=torch.uint8)
                f.write((a & b).cpu().numpy().tobytes())
    if world > 1:
        dist.barrier()
    sh = xd.ColumnShardedIndex(path, rank=rank, world=world, device=local) if world > 1 else None
    ix = sh.index if sh else engine.CobsIndex(path, device=local)
    genome = synth.synth_genome(1_000_000, seed=7)
    reads = synth.synth_reads(genome, N_READS, L, seed=8, device=dev)
    hb, he = synth.fixed_offsets(TILE, L)
    d_b = torch.from_numpy(hb.view(np.int64)).to(dev)
    d_e = torch.from_numpy(he.view(np.int64)).to(dev)
    tiles = [(reads.data_ptr() + t * TILE * L, TILE * L, d_b.data_ptr(), d_e.data_ptr(), TILE) for t in range(N_READS // TILE)]
    calls = {}

    def consume_rows(t, scores):           # the consumer of the full score rows: per-read argmax on the device
        calls[t] = scores.argmax(dim=1)

    def run_gather():
        if sh:
            sh.query_tiles(iter(tiles), 1, XS_U8, consume_rows)
        else:
            s = torch.cuda.current_stream().cuda_stream
            for t, (b0, nb, pb, pe, n) in enumerate(tiles):
                out = torch.empty((n, D), dtype=torch.uint8, device=dev)
                ix.query_device(b0, nb, pb, pe, n, 1, XS_U8, out.data_ptr(), s)
                consume_rows(t, out)
        torch.cuda.synchronize()

    best = {}

    def run_reduced():
        if sh:
            sh.classify_tiles(iter(tiles), 1, XS_U8, lambda t, b, c, tie: best.__setitem__(t, b))
        torch.cuda.synchronize()

    def timed(fn):
        fn()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        fn()
        dt = time.perf_counter() - t0
        if world > 1:
            t = torch.tensor([dt], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
        return dt

    t_gather = timed(run_gather)
    t_reduced = timed(run_reduced) if sh else None
    if sh:
        assert all(torch.equal(calls[t], best[t]) for t in calls)
    if rank == 0:
        # parity of the combined argmax on a sample
        from oracle import oracle
        sample = 200
        exp = oracle.CobsOracle(path, load_complete=False).counts_batch(reads[: sample * L].cpu().numpy(), hb[:sample], he[:sample], 1, threads=8)
        assert np.array_equal(calls[0][:sample].cpu().numpy(), np.minimum(exp, 255).argmax(axis=1))
        lookups = N_READS * (L - K + 1)
        print(json.dumps({"config": f"cfg5 geometry D={D} h={H} S={S}, {N_READS} x {L}bp reads, column-sharded x{world}",
                          "n_gpus": world, "allgather_s": t_gather, "lookups_per_sec_allgather": lookups / t_gather,
                          "reduced_exchange_s": t_reduced, "lookups_per_sec_reduced": (lookups / t_reduced) if t_reduced else None,
                          "row_bytes_per_gpu": int(ix.info.row_stride), "parity_sample_reads": sample}), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
