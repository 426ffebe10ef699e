"""world_size-2 gloo test (CPU) of the multi-process host logic: sequence sharding and max-over-ranks timing."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from fseend_b200.parallel import max_over_ranks, shard_range, shard_sequences


def test_shard_range_covers_everything_once():
    for n in (0, 1, 7, 64, 65):
        for world in (1, 2, 3, 8):
            got = []
            for r in range(world):
                a, b = shard_range(n, r, world)
                assert 0 <= a <= b <= n and (b - a) - n // world in (0, 1)
                got += list(range(a, b))
            assert got == list(range(n))


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    items = list(range(11))
    mine = shard_sequences(items, rank, world)
    # each rank "times" its shard; the job time is the slowest rank's
    t = max_over_ranks([float(len(mine)), 10.0 - rank], torch.device("cpu"))
    gathered = [None] * world
    dist.all_gather_object(gathered, mine)
    if rank == 0:
        out.put((t, gathered))
    dist.barrier()
    dist.destroy_process_group()


def test_two_process_gloo_sharding_and_max_timing():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    t, gathered = q.get(timeout=120)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert t == [6.0, 10.0]
    assert gathered[0] + gathered[1] == list(range(11))
