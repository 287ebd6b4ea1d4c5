"""world_size-2 (and 3) gloo tests of the multi-GPU host logic: read sharding, document-column sharding and
the score all-gather.  The per-rank 'query' is the oracle on the rank's column slice, so the test runs without
a GPU; the GPU kernels' column-shard parity is covered by tests/test_gpu_parity.py."""
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


def _worker(rank: int, world: int, port: int, tmp: str, mode: str):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import oracle
        from tests import synth
        from xspect2_b200 import distributed as xd
        path = Path(tmp) / "index.cobs_classic"
        orc = oracle.CobsOracle(path)
        rng = np.random.default_rng(99)          # same reads on every rank
        genome = synth.random_dna(rng, 3000)
        bases, b, e = synth.sample_reads(rng, [genome], 101, (21, 200), n_rate=0.002)
        full = orc.counts_batch(bases, b, e)
        if mode == "reads":
            lo, hi = xd.read_shard(b.size, rank, world)
            local = orc.counts_batch(bases, b[lo:hi], e[lo:hi])
            assert np.array_equal(local, full[lo:hi])
            totals = xd.allreduce_totals(local.sum(axis=0))
            assert np.array_equal(totals.numpy(), full.sum(axis=0, dtype=np.int64))
            # every record lands on exactly one rank
            cover = torch.zeros(b.size, dtype=torch.int64)
            cover[lo:hi] = 1
            dist.all_reduce(cover)
            assert bool((cover == 1).all())
        elif mode == "grid":
            # 2 column groups x world / 2 read groups: columns all-gathered inside the group, records dealt to the groups
            c = 2
            rg, cr, members = xd.grid_layout(rank, world, c)
            groups = [dist.new_group(list(range(g * c, (g + 1) * c))) for g in range(world // c)]   # every rank creates all
            shards = xd.column_shards(orc.n_docs, c, align=8)
            lo, hi = xd.read_shard(b.size, rg, world // c)
            dlo, dhi = shards[cr]
            local = torch.from_numpy(np.ascontiguousarray(orc.counts_batch(bases, b[lo:hi], e[lo:hi])[:, dlo:dhi]))
            got = xd.allgather_columns(local, shards, group=groups[rg])
            assert np.array_equal(got.numpy(), full[lo:hi])
            cover = torch.zeros(b.size, dtype=torch.int64)
            if cr == 0:
                cover[lo:hi] = 1
            dist.all_reduce(cover)
            assert bool((cover == 1).all())
        else:
            shards = xd.column_shards(orc.n_docs, world, align=8 if mode == "cols8" else 128)
            assert shards[0][0] == 0 and shards[-1][1] == orc.n_docs
            assert all(a[1] == bb[0] for a, bb in zip(shards, shards[1:]))
            lo, hi = shards[rank]
            # reduced exchange: each rank's local (best, count, multiplicity) -> global best / tie flag
            loc = full[:, lo:hi].astype(np.int64)
            if hi > lo:
                lb, lc = loc.argmax(axis=1), loc.max(axis=1)
                ln = (loc == lc[:, None]).sum(axis=1)
            else:
                lb = lc = ln = np.zeros(full.shape[0], np.int64)
            gb, gc, gt = xd.merge_local_best(torch.from_numpy(lb), torch.from_numpy(lc), torch.from_numpy(ln), shards, rank)
            f = full.astype(np.int64)
            assert np.array_equal(gc.numpy(), f.max(axis=1))
            assert np.array_equal(gb.numpy(), f.argmax(axis=1))
            assert np.array_equal(gt.numpy(), (f == f.max(axis=1)[:, None]).sum(axis=1) > 1)
            for dt in (np.uint8, np.uint16, np.uint32):
                local = torch.from_numpy(np.ascontiguousarray(np.minimum(full[:, lo:hi], np.iinfo(dt).max).astype(dt)))
                got = xd.allgather_columns(local, shards)
                assert got.dtype == local.dtype
                assert np.array_equal(got.numpy(), np.minimum(full, np.iinfo(dt).max).astype(dt))
    finally:
        dist.destroy_process_group()


def _run(world: int, mode: str, tmp_path, n_docs: int):
    from oracle import oracle
    from tests import synth
    rng = np.random.default_rng(5)
    docs = synth.make_genomes(rng, n_docs, 500)
    oracle.write_classic(tmp_path / "index.cobs_classic", docs, k=21, num_hashes=3, fpr=0.05)
    port = 29500 + (os.getpid() + world * 7 + len(mode)) % 2000
    mp.spawn(_worker, args=(world, port, str(tmp_path), mode), nprocs=world, join=True)


def test_read_sharding_world2(tmp_path):
    _run(2, "reads", tmp_path, 20)


def test_column_sharding_allgather_world2(tmp_path):
    _run(2, "cols", tmp_path, 300)


def test_column_sharding_uneven_world3(tmp_path):
    _run(3, "cols8", tmp_path, 50)


def test_grid_world4(tmp_path):
    _run(4, "grid", tmp_path, 50)


def test_grid_layout_and_column_group_policy():
    from xspect2_b200 import distributed as xd
    assert [xd.grid_layout(r, 8, 2)[:2] for r in range(8)] == [(0, 0), (0, 1), (1, 0), (1, 1), (2, 0), (2, 1), (3, 0), (3, 1)]
    assert xd.grid_layout(5, 8, 2)[2] == [4, 5] and xd.grid_layout(3, 4, 4) == (0, 3, [0, 1, 2, 3])
    with pytest.raises(ValueError):
        xd.grid_layout(0, 8, 3)
    hbm = 180 * 10**9
    # BASELINE config 5: 96 M rows at a 1280-byte stride = 123 GB -> two column groups on 2, 4 and 8 B200s
    assert [xd.choose_column_groups(96_000_000 * 1280, hbm, w) for w in (1, 2, 4, 8)] == [1, 2, 2, 2]
    assert xd.choose_column_groups(2_400_000_000, hbm, 8) == 1            # the config-2 index is replicated
    assert xd.choose_column_groups(600 * 10**9, hbm, 8) == 8 and xd.choose_column_groups(300 * 10**9, hbm, 8) == 4


def test_shard_boundaries():
    from xspect2_b200 import distributed as xd
    assert [xd.read_shard(10, r, 4) for r in range(4)] == [(0, 2), (2, 5), (5, 7), (7, 10)]
    for n_docs, world in [(90, 2), (10000, 8), (10000, 3), (129, 2), (128, 2)]:
        sh = xd.column_shards(n_docs, world)
        assert sh[0][0] == 0 and sh[-1][1] == n_docs
        assert all(lo % 128 == 0 for lo, _ in sh)
        assert sum(hi - lo for lo, hi in sh) == n_docs
