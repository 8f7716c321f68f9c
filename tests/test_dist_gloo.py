"""world_size-2 gloo test of the multi-GPU plumbing (stream pinning + max-over-ranks timing) on CPU."""
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
    from videosd_b200.parallel import aggregate_fps, max_over_ranks, shard_streams

    mine = shard_streams(8, world, rank)
    seconds = 1.0 + rank            # rank 1 is the slow one
    slowest = max_over_ranks(seconds)
    fps = aggregate_fps(len(mine) * 10, seconds)
    gathered = [None] * world
    dist.all_gather_object(gathered, mine)
    q.put((rank, mine, slowest, fps, gathered))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharding_and_timing():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (r0, m0, s0, f0, g0), (r1, m1, s1, f1, g1) = res
    assert m0 == [0, 2, 4, 6] and m1 == [1, 3, 5, 7]
    assert sorted(g0[0] + g0[1]) == list(range(8))
    assert s0 == s1 == 2.0                        # the max over ranks, on every rank
    assert abs(f0 - 80 / 2.0) < 1e-9 and f0 == f1  # all frames / slowest rank's time
