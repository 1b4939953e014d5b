"""CPU tier: the multi-GPU host logic (static split + verdict gather to rank 0, as
bench.py uses it) with world_size 2 on the gloo backend."""
import ctypes as C
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import affaction_b200 as ab
from affaction_b200 import sharding


def test_shard_bounds_cover_everything():
    for total in (0, 1, 7, 1000, 28416):
        for world in (1, 2, 4, 8):
            blocks = [sharding.shard_bounds(total, r, world) for r in range(world)]
            assert blocks[0][0] == 0 and blocks[-1][1] == total
            assert all(blocks[i][1] == blocks[i + 1][0] for i in range(world - 1))
            sizes = [hi - lo for lo, hi in blocks]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        sharding.shard_bounds(10, 2, 2)


def _worker(rank, world, port, total, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = sharding.shard_bounds(total, rank, world)
    recs = (ab.Result * (hi - lo))()
    for k, i in enumerate(range(lo, hi)):
        recs[k].idx = i
        recs[k].success = i % 3 == 0
        recs[k].jl_cost = 0.001 * i
    local = torch.frombuffer(bytearray(bytes(recs)), dtype=torch.uint8)
    out = sharding.gather_verdicts(local, world, dist, rank)
    if rank == 0:
        got = sharding.records_from_bytes(out.numpy().tobytes(), ab.Result)
        q.put((rank, [(r.idx, r.success, r.jl_cost) for r in got]))
    else:
        q.put((rank, out))
    dist.destroy_process_group()


def test_gather_verdicts_world2():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    total, world = 64, 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, total, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = dict(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    expect = [(i, int(i % 3 == 0), 0.001 * i) for i in range(total)]
    assert results[0] == expect and results[1] is None     # the records of every rank, in rank order, on rank 0 only
    order = ab.rank([_mk(*e) for e in expect])
    assert order[0] == 0 and expect[order[0]][1] == 1


def _mk(idx, success, jl):
    r = ab.Result()
    r.idx, r.success, r.jl_cost = idx, success, jl
    return r
