"""bench.py's config-5 leg alone (no config-2 setup), for one or more layouts:

    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tests/perf_cfg5_grid.py [C ...]

C = column groups (divisor of N); default: the policy layout (bench.cfg5_column_groups) and pure column sharding (C = N).
One JSON line per layout on rank 0.  Not collected by pytest."""
import argparse
import json
import os
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))


def main() -> None:
    import torch
    import torch.distributed as dist
    import bench

    ap = argparse.ArgumentParser()
    ap.add_argument("col_groups", nargs="*", type=int)
    ap.add_argument("--steps", type=int, default=3)
    args = ap.parse_args()
    rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("WORLD_SIZE", 1), ("LOCAL_RANK", 0)))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    try:
        layouts = args.col_groups or sorted({bench.cfg5_column_groups(world, local), world})
        for c in layouts:
            out = bench.run_cfg5(args, rank, world, local, col_groups=c)
            if rank == 0:
                print(json.dumps(out), flush=True)
    finally:
        if world > 1:
            dist.destroy_process_group()


if __name__ == "__main__":
    main()
