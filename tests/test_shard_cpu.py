"""The N > 1 path on CPU: two processes over gloo (127.0.0.1).  Every rank takes its slice of a
batch of bodies, the slices must partition the batch, and timings reduce to the max over ranks --
exactly what bench.py does over NCCL with one rank per GPU (no data-path collective exists)."""
import importlib
import os
import socket
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n_bodies, out):
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    shard = importlib.import_module("cs121-softbodysim_b200.shard")
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        mine = shard.body_slice(n_bodies, world, rank)
        counts = shard.gather_counts(len(mine), dist)
        # pretend device times: rank r took (r + 1) * 10 ms for its slice, wall (r + 1) s
        ms, wall = shard.reduce_max([10.0 * (rank + 1), 1.0 * (rank + 1)], dist)
        out.put((rank, mine, counts, ms, wall))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_bodies", [4096, 7, 1])
def test_two_ranks_partition_the_batch_and_reduce_to_max(n_bodies):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_bodies, out)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(out.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    owned = sorted(b for _, mine, _, _, _ in res for b in mine)
    assert owned == list(range(n_bodies))                       # disjoint and complete
    for rank, mine, counts, ms, wall in res:
        assert counts == [len(res[0][1]), len(res[1][1])] and sum(counts) == n_bodies
        assert abs(len(res[0][1]) - len(res[1][1])) <= 1
        assert ms == 20.0 and wall == 2.0                       # max over ranks, identical on every rank


def test_single_process_helpers():
    shard = importlib.import_module("cs121-softbodysim_b200.shard")
    assert shard.body_slice(10, 1, 0) == list(range(10))
    assert shard.body_slice(10, 4, 3) == [3, 7]
    assert shard.reduce_max([1.5, 2.5]) == [1.5, 2.5]
    assert shard.gather_counts(5) == [5]
    with pytest.raises(ValueError):
        shard.body_slice(4, 2, 2)
