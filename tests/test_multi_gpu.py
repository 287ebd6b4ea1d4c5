"""Two-GPU tests (NCCL): document-column sharding with the score all-gather, and read sharding, against the
oracle.  Skipped on boxes with fewer than two GPUs (the single-GPU column-shard parity is in test_gpu_parity)."""
import os
import sys
from pathlib import Path

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
pytestmark = pytest.mark.gpu


def _worker(rank: int, world: int, port: int, tmp: str):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        from oracle import oracle
        from tests import synth
        from xspect2_b200 import distributed as xd
        from xspect2_b200 import engine
        from xspect2_b200._abi import XS_U8, XS_U16
        path = Path(tmp) / "index.cobs_classic"
        orc = oracle.CobsOracle(path)
        rng = np.random.default_rng(123)
        genomes = [synth.random_dna(rng, 2000) for _ in range(3)]
        bases, b, e = synth.sample_reads(rng, genomes, 4000, (21, 220), n_rate=0.002)
        full = orc.counts_batch(bases, b, e, 1, threads=4)
        dev = torch.device("cuda", rank)

        # ---- document-column sharding + all-gather, tiles of 1000 records
        sh = xd.ColumnShardedIndex(path, rank=rank, world=world, device=rank)
        d_bases = torch.from_numpy(bases).to(dev)
        d_b = torch.from_numpy(b.view(np.int64)).to(dev)
        d_e = torch.from_numpy(e.view(np.int64)).to(dev)
        tiles = [(d_bases.data_ptr(), bases.size, d_b[i:i + 1000].data_ptr(), d_e[i:i + 1000].data_ptr(), 1000) for i in range(0, 4000, 1000)]
        for dt, npdt in ((XS_U8, np.uint8), (XS_U16, np.uint16)):
            got = {}
            sh.query_tiles(iter(tiles), 1, dt, lambda t, scores: got.__setitem__(t, scores.cpu().numpy()))
            assert sorted(got) == [0, 1, 2, 3]
            allrows = np.concatenate([got[t] for t in range(4)])
            assert np.array_equal(allrows, np.minimum(full, np.iinfo(npdt).max).astype(npdt)), f"rank {rank} dtype {dt}"

        # ---- reduced exchange: local argmax per rank, 3 integers per record travel
        calls = {}
        totals = sh.classify_tiles(iter(tiles), 1, XS_U8, lambda t, bst, c, tie: calls.__setitem__(t, (bst.cpu().numpy(), c.cpu().numpy(), tie.cpu().numpy())))
        f8 = np.minimum(full, 255).astype(np.int64)
        gb = np.concatenate([calls[t][0] for t in range(4)])
        gc = np.concatenate([calls[t][1] for t in range(4)])
        gt = np.concatenate([calls[t][2] for t in range(4)])
        assert np.array_equal(gb, f8.argmax(axis=1)) and np.array_equal(gc, f8.max(axis=1))
        assert np.array_equal(gt, (f8 == f8.max(axis=1)[:, None]).sum(axis=1) > 1)
        lo_d, hi_d = sh.shards[rank]
        assert np.array_equal(totals.cpu().numpy(), f8[:, lo_d:hi_d].sum(axis=0))

        # ---- the exchange behind the C ABI: xs_comm_init + xs_cobs_query_device_ld + xs_allgather_scores (ncclAllGather)
        # + xs_sharded_reduce_device consuming [world][n][w] in place, double-buffered over the tiles
        comm = xd.make_comm(rank, world, rank)
        assert comm.nccl_version > 20000
        scorer = xd.ShardedScorer(sh.index, sh.shards, comm, XS_U8, 1000)
        res = {}
        scorer.run(iter(tiles), 1, lambda t, bst, c, nb_: res.__setitem__(t, (bst.cpu().numpy(), c.cpu().numpy(), nb_.cpu().numpy())),
                   time_exchange=True)
        assert sorted(res) == [0, 1, 2, 3] and scorer.exchange_ms > 0
        assert np.array_equal(np.concatenate([res[t][0] for t in range(4)]), f8.argmax(axis=1))
        assert np.array_equal(np.concatenate([res[t][1] for t in range(4)]), f8.max(axis=1))
        assert np.array_equal(np.concatenate([res[t][2] for t in range(4)]), (f8 == f8.max(axis=1)[:, None]).sum(axis=1))
        tot = torch.from_numpy(f8[:, lo_d:hi_d].sum(axis=0)).to(dev)
        full_tot = torch.zeros(f8.shape[1], dtype=torch.int64, device=dev)
        full_tot[lo_d:hi_d] = tot
        comm.allreduce_totals(full_tot.data_ptr(), full_tot.numel(), torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
        assert np.array_equal(full_tot.cpu().numpy(), f8.sum(axis=0))
        comm.close()

        # ---- read sharding: whole index per GPU, a slice of the records per rank, totals all-reduced
        ix = engine.CobsIndex(path, device=rank)
        lo, hi = xd.read_shard(b.size, rank, world)
        local = ix.query(bases, b[lo:hi], e[lo:hi], 1)
        assert np.array_equal(local.astype(np.uint32), full[lo:hi])
        totals = xd.allreduce_totals(torch.from_numpy(local.sum(axis=0, dtype=np.int64)).to(dev))
        assert np.array_equal(totals.cpu().numpy(), full.sum(axis=0, dtype=np.int64))
    finally:
        dist.destroy_process_group()


def test_two_gpu_column_and_read_sharding(tmp_path, gpu, oracle):
    if gpu.device_count() < 2:
        pytest.skip("needs two GPUs (gpurun --gpus 2)")
    from tests import synth
    rng = np.random.default_rng(6)
    docs = synth.make_genomes(rng, 1000, 600)
    oracle.write_classic(tmp_path / "index.cobs_classic", docs, k=21, num_hashes=7, fpr=0.01)
    mp.spawn(_worker, args=(2, 29400 + os.getpid() % 500, str(tmp_path)), nprocs=2, join=True)
