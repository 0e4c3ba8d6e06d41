"""CPU, world_size 2 over gloo: the N>1 plumbing of the picture-/GOP-parallel path
(sharding without overlap or gaps, max-over-ranks timing, whole-job throughput)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from schroedinger_b200 import sharding
    mine = sharding.shard(13, rank, world)
    gops = sharding.gop_shard(37, 8, rank, world)
    # every rank "takes" a different time; the job takes the slowest rank's
    elapsed = sharding.max_over_ranks(1.0 + rank)
    # gather the shards to check the partition
    got = [None] * world
    dist.all_gather_object(got, (mine, gops))
    dist.barrier()
    q.put((rank, elapsed, got, sharding.aggregate_throughput(8, elapsed, world)))
    dist.destroy_process_group()


def test_two_rank_sharding_and_timing():
    world = 2
    port = _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, elapsed, got, thr in results:
        assert elapsed == 2.0                       # max over ranks, on every rank
        units = sorted(sum((g[0] for g in got), []))
        assert units == list(range(13))             # no gaps, no overlap
        pics = sorted(sum((g[1] for g in got), []))
        assert pics == list(range(37))
        for g in got:                               # GOPs stay whole
            for p0 in g[1]:
                assert (p0 // 8) % world == got.index(g)
        assert thr == 8 * world / 2.0


def test_single_process_defaults():
    from schroedinger_b200 import sharding
    assert sharding.shard(5, 0, 1) == [0, 1, 2, 3, 4]
    assert sharding.max_over_ranks(3.5) == 3.5
