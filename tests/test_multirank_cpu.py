"""CPU test of the N>1 path of bench.py with the gloo backend (world_size 2): configuration
broadcast, max-over-ranks timing, parity flags over all ranks, stream sharding.  The data path itself has no collective."""
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
    every = (bench.all_ranks_ok(True, "cpu"), bench.all_ranks_ok(rank == 0, "cpu"))   # parity flags are ANDed over the ranks
    mine = bench.streams_of_rank(11, world, rank)
    gathered = [None] * world
    dist.all_gather_object(gathered, mine)
    q.put((rank, cfg, worst, gathered, bench.seed_of_rank(rank), every))
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
    for rank, cfg, worst, gathered, seed, every in res:
        assert every == (True, False)                    # one rank's failed check fails the job's parity flag
        assert cfg == [640.0, 20.0, 5.0, 16.0]          # every rank has rank 0's configuration
        assert worst == 15.0                             # max over ranks
        allstreams = sorted(s for part in gathered for s in part)
        assert allstreams == list(range(11))             # every stream decoded exactly once
        assert seed == 1 + rank


def test_clock_sampler_reports_the_samples_of_the_timed_region_only():
    """bench.ClockSampler: nvidia-smi runs from before the warm-up; mark() brackets the timed region and only its samples
    (or, for a region shorter than the sampling interval, the nearest ones) are reported"""
    sys.path.insert(0, ROOT)
    import bench

    class FakeProc:
        def terminate(self): pass
        def wait(self, timeout=None): return 0
        def kill(self): pass

    s = bench.ClockSampler(0)
    s.proc = FakeProc()
    row = lambda sm, hw="Not Active", pw="Not Active": "%d, 1965, 300.0, %s, Not Active, Not Active, %s" % (sm, hw, pw)
    # start-up and warm-up samples (a throttled one among them), then the timed region, then one long after it
    s.rows = [(0.0, row(600, hw="Active")), (0.5, row(1200)), (1.00, row(1950)), (1.20, row(1965, pw="Active")), (1.40, row(1965)), (3.0, row(500))]
    s.marks = [0.95, 1.30]
    r = s.stop()
    assert r["sm_mhz"] == 1965.0 and r["sm_max_mhz"] == 1965.0 and r["samples"] == 3 and r["reasons"] == ["sw_power_cap"]
    # a region that no sample fell into: the two nearest
    s2 = bench.ClockSampler(0)
    s2.proc = FakeProc()
    s2.rows = [(0.0, row(1000)), (1.0, row(1900)), (2.0, row(1960)), (9.0, row(700))]
    s2.marks = [1.40, 1.45]
    r2 = s2.stop()
    assert r2["samples"] == 2 and r2["sm_mhz"] == 1930.0
    # all GPUs of a box from one process: every row counts
    s3 = bench.ClockSampler(None)
    assert s3.index is None
