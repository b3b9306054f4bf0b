"""Multi-GPU host logic on CPU: frame sharding + the max-over-ranks reduction (gloo, world size 2)."""
import os
import socket

import pytest

from vkresample_b200 import sharding as sh


def test_striding_matches_reference_formula():
    for num_files in (1, 2, 5, 8, 9, 256, 257):
        for workers in (1, 2, 3, 8):
            seen = []
            for t in range(workers):
                mine = sh.frames_for_worker(num_files, workers, t)
                assert mine == [f for f in range(1, num_files + 1) if (f - 1) % workers == t]
                seen += mine
            assert sorted(seen) == list(range(1, num_files + 1))


def test_device_mapping():
    assert [sh.device_for_worker(t, 8) for t in range(10)] == [0, 1, 2, 3, 4, 5, 6, 7, 0, 1]
    assert [sh.device_for_worker(t, 1, 3) for t in range(3)] == [0, 0, 0]
    assert sh.aggregate_frames_per_s(32, 8, 0.5) == 512.0


def _worker(rank, world, port, num_files, q):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mine = sh.frames_for_worker(num_files, world, rank)
    # every rank "processes" its frames: result = frame number squared; gather and restore order
    gathered = [None] * world
    dist.all_gather_object(gathered, [(f, f * f) for f in mine])
    # timing reduction exactly as bench.py does it: MAX over ranks
    t = torch.tensor([0.1 * (rank + 1)], dtype=torch.float64)
    dist.barrier()
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        flat = sorted(x for part in gathered for x in part)
        q.put((flat, float(t.item())))
    dist.destroy_process_group()


def test_two_rank_gloo_sharding():
    import torch.multiprocessing as mp
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    num_files = 9
    procs = [ctx.Process(target=_worker, args=(r, 2, port, num_files, q)) for r in range(2)]
    [p.start() for p in procs]
    flat, tmax = q.get(timeout=120)
    [p.join(timeout=60) for p in procs]
    assert all(p.exitcode == 0 for p in procs)
    assert flat == [(f, f * f) for f in range(1, num_files + 1)]   # every frame exactly once, order restored
    assert abs(tmax - 0.2) < 1e-12                                  # slowest rank defines the job time
    assert sh.aggregate_frames_per_s(5, 2, tmax) == pytest.approx(50.0)
