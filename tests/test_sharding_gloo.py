"""world_size-2 gloo test (CPU) of the N>1 path: frame sharding is a partition, per-rank seeds are
distinct, and the whole-job throughput bookkeeping is max-time / sum-units."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from link_b200.sharding import frame_seed, reduce_throughput, shard_frames


def _free_port():
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    mine = shard_frames(11, rank, world)
    gathered = [None] * world
    dist.all_gather_object(gathered, mine)
    ms, units = reduce_throughput(10.0 + 5.0 * rank, 1000.0 * (rank + 1))
    seeds = [None] * world
    dist.all_gather_object(seeds, [frame_seed(rank, s) for s in range(3)])
    if rank == 0:
        out.put((gathered, ms, units, seeds))
    dist.destroy_process_group()


def test_two_rank_sharding_and_reduction():
    ctx = mp.get_context('spawn')
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    gathered, ms, units, seeds = out.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    flat = sorted(i for part in gathered for i in part)
    assert flat == list(range(11))                      # partition: disjoint and complete
    assert gathered[0] == [0, 2, 4, 6, 8, 10] and gathered[1] == [1, 3, 5, 7, 9]
    assert ms == 15.0 and units == 3000.0               # max time, summed units
    assert len({s for part in seeds for s in part}) == 6


def test_single_process_is_identity():
    assert reduce_throughput(3.5, 7.0) == (3.5, 7.0)
    assert shard_frames(5, 0, 1) == [0, 1, 2, 3, 4]
