"""CPU test of the N>1 path of bench.py with the gloo backend (world_size 2): configuration
broadcast, max-over-ranks timing, stream sharding.  The data path itself has no collective."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import bench
    cfg = bench.broadcast_config([640.0, 20, 5, 16] if rank == 0 else [0, 0, 0, 0], "cpu")
    worst = bench.max_over_ranks(10.0 + 5 * rank, "cpu")
    mine = bench.streams_of_rank(11, world, rank)
    gathered = [None] * world
    dist.all_gather_object(gathered, mine)
    q.put((rank, cfg, worst, gathered, bench.seed_of_rank(rank)))
    dist.barrier()
    dist.destroy_process_group()


def test_two_ranks_gloo():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, cfg, worst, gathered, seed in res:
        assert cfg == [640.0, 20.0, 5.0, 16.0]          # every rank has rank 0's configuration
        assert worst == 15.0                             # max over ranks
        allstreams = sorted(s for part in gathered for s in part)
        assert allstreams == list(range(11))             # every stream decoded exactly once
        assert seed == 1 + rank
