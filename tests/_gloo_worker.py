"""worker of tests/test_ranks_gloo.py: run under torchrun with the gloo backend (CPU, world_size 2+).
Exercises the host-side logic of the one-process-per-GPU path: partition, unique-id distribution, the in-place
slice exchange the engine performs with NCCL, and the max/sum reductions bench.py reports with."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from steps_b200 import ranks  # noqa: E402


def main():
    rank, world, _ = ranks.env_rank()
    dist.init_process_group("gloo")
    try:
        for n in (10, 1001, 4096 + rank * 0):
            parts = ranks.all_partitions(n, world)
            # contiguous, ordered, covering [0, n), sizes differing by at most one
            assert parts[0][0] == 0 and parts[-1][1] == n
            assert all(parts[r][1] == parts[r + 1][0] for r in range(world - 1))
            sizes = [hi - lo for lo, hi in parts]
            assert max(sizes) - min(sizes) <= 1
            # every rank agrees on every other rank's range
            got = [None] * world
            dist.all_gather_object(got, ranks.partition(n, world, rank))
            assert [tuple(g) for g in got] == parts

            # position exchange: each rank only has valid data in its own slice; after the exchange all replicas agree
            rng = np.random.default_rng(1234)
            truth = rng.standard_normal(3 * n)
            lo, hi = parts[rank]
            mine = np.full(3 * n, np.nan)
            mine[3 * lo:3 * hi] = truth[3 * lo:3 * hi]
            x = torch.from_numpy(mine)
            ranks.gather_owned_slices(dist, rank, world, x, n)
            assert np.array_equal(x.numpy(), truth), "replica differs after the slice exchange"

        uid = ranks.share_unique_id(dist, rank, world, lambda: bytes((7 * i + 3) % 256 for i in range(128)))
        assert uid == bytes((7 * i + 3) % 256 for i in range(128))
        try:
            ranks.share_unique_id(dist, rank, world, lambda: b"short")
            raise AssertionError("short id accepted")
        except ranks._lib.StepsError:
            pass

        assert ranks.reduce_scalar(dist, world, 1.0 + rank, "max") == float(world)
        assert ranks.reduce_scalar(dist, world, 1.0 + rank, "sum") == world * (world + 1) / 2.0
        dist.barrier()
        if rank == 0:
            print("GLOO_WORKER_OK", world)
    finally:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
