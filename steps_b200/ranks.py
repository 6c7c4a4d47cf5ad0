"""One-process-per-GPU plumbing over ``torch.distributed`` (any backend): what a launcher needs around the
C ABI's ``steps_b200_engine_comm_init`` -- the analogue of the reference's MPI calls on the force path
(StePS/src/main.cc:1198-1239 Bcast of N/x/M, :1581-1607 i-partition, step.cc:183-228 Bcast x / gather F).

Nothing here computes forces; it only moves small host values between ranks, so it runs with the ``gloo``
backend on CPU (tests/test_ranks_gloo.py) exactly as it does with ``nccl`` under torchrun (bench.py).
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Callable, List, Tuple

from . import _lib


def env_rank() -> Tuple[int, int, int]:
    """(rank, world, local_rank) from the torchrun environment; (0, 1, 0) when launched plainly"""
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))


def partition(n: int, nranks: int, rank: int) -> Tuple[int, int]:
    """[i_lo, i_hi) owned by `rank`: steps_b200_partition (contiguous, remainder spread one-each; replaces
    main.cc:1581-1607 / forces_cuda.cu:942-951 which hand the whole remainder to rank/GPU 0)"""
    lo, hi = C.c_int(), C.c_int()
    _lib.load().steps_b200_partition(n, nranks, rank, C.byref(lo), C.byref(hi))
    return lo.value, hi.value


def all_partitions(n: int, nranks: int) -> List[Tuple[int, int]]:
    return [partition(n, nranks, r) for r in range(nranks)]


def share_unique_id(dist, rank: int, world: int, make_id: Callable[[], bytes]) -> bytes:
    """rank 0 creates the 128-byte NCCL unique id (steps_b200_nccl_unique_id) and every rank receives it"""
    if world == 1:
        return make_id()
    ids = [make_id() if rank == 0 else None]
    dist.broadcast_object_list(ids, src=0)
    uid = ids[0]
    if not isinstance(uid, (bytes, bytearray)) or len(uid) != 128:
        raise _lib.StepsError("NCCL unique id must be 128 bytes")
    return bytes(uid)


def reduce_scalar(dist, world: int, v: float, op: str, device: str = "cpu") -> float:
    """max / sum of one float over ranks (timings are reported as the max over ranks, counts as the sum)"""
    if world == 1:
        return float(v)
    import torch

    t = torch.tensor([v], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX if op == "max" else dist.ReduceOp.SUM)
    return float(t.item())


def gather_owned_slices(dist, rank: int, world: int, x_replica, n: int):
    """Host emulation of the engine's per-step position exchange (engine.cu gather_positions): every rank
    broadcasts the AoS slice x[3*lo:3*hi] it owns, in place, into every other rank's replica.  `x_replica`
    is a 1-D torch tensor of length 3n.  Used by the gloo tests to pin the slice arithmetic the NCCL path uses."""
    if world == 1:
        return x_replica
    for r in range(world):
        lo, hi = partition(n, world, r)
        dist.broadcast(x_replica[3 * lo:3 * hi], src=r)
    return x_replica
