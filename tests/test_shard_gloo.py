"""CPU test of the N > 1 path (world_size 2, gloo): the frame/QP partition is disjoint and complete, the timing reduction
is the maximum over ranks, and the final gather restores the single-process order."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from vvc_intra_b200 import shard


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    units = shard.work_units(8, (32, 27, 37, 22))
    mine = shard.shard_units(units, rank, world)
    dist.barrier()
    slowest = shard.max_over_ranks(10.0 + 5.0 * rank, dist)
    stats = shard.gather_stats([dict(unit=u, rank=rank) for u in mine], dist)
    q.put((rank, mine, slowest, stats))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_partition_and_reductions():
    world = 2
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    units = shard.work_units(8, (32, 27, 37, 22))
    assert len(units) == 32 and len(set(units)) == 32
    a, b = got[0][1], got[1][1]
    assert not set(a) & set(b) and sorted(a + b) == sorted(units) and abs(len(a) - len(b)) <= 1
    assert set(units) == {(f, q) for f in range(8) for q in (32, 27, 37, 22)}
    assert got[0][2] == got[1][2] == 15.0                       # max over ranks on every rank
    for _, _, _, stats in got:
        assert [s['unit'] for s in stats] == units              # the gather restores the sequence order
        assert [s['rank'] for s in stats] == [i % 2 for i in range(32)]


def test_single_process_degenerates():
    units = shard.work_units(3, (22, 37))
    assert shard.shard_units(units, 0, 1) == units
    assert shard.max_over_ranks(3.5) == 3.5 and shard.gather_stats([1, 2]) == [1, 2]
    with pytest.raises(ValueError):
        shard.shard_units(units, 2, 2)
