"""Sharding of independent bodies across ranks (one process per GPU) and the max-over-ranks
reduction of device timings.  The path has no data-path collective: bodies never interact
(the reference runs one body per PBDServer process, CProgram/src/main.cpp:69-98), so a rank only
steps its own slice; the only communication is the timing reduction bench.py prints.

Pure host logic (torch.distributed with any backend): covered on CPU with gloo, world size 2
(tests/test_shard_cpu.py); bench.py uses the same functions over NCCL.
"""
from __future__ import annotations


def body_slice(n_bodies: int, world: int, rank: int) -> list[int]:
    """Indices of the bodies rank `rank` owns: round-robin (b % world == rank), so every rank gets
    the same mix of bodies whatever their order, and the slices differ in size by at most one."""
    if world < 1 or not 0 <= rank < world:
        raise ValueError("need 0 <= rank < world")
    return list(range(rank, n_bodies, world))


def reduce_max(values, dist=None, device=None):
    """Element-wise max over all ranks of a short list of floats (device ms, wall seconds)."""
    vals = [float(v) for v in values]
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return vals
    import torch
    t = torch.tensor(vals, dtype=torch.float64, device=device if device is not None else "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return [float(x) for x in t.tolist()]


def gather_counts(n_local: int, dist=None, device=None) -> list[int]:
    """How many bodies every rank holds (rank order); used to check that the slices cover the batch."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return [int(n_local)]
    import torch
    t = torch.zeros(dist.get_world_size(), dtype=torch.int64, device=device if device is not None else "cpu")
    t[dist.get_rank()] = int(n_local)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return [int(x) for x in t.tolist()]
