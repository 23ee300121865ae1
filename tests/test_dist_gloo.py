"""N>1 host logic on CPU: world_size-2 gloo run of the stream sharding + end-of-run reduction
that bench.py uses under torchrun (the data path itself has no collective)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from backscrub_b200 import sharding


def test_streams_partition():
    for world in (1, 2, 4, 8):
        seen = []
        for r in range(world):
            seen += sharding.streams_for_rank(8, r, world)
        assert sorted(seen) == list(range(8))
    assert sharding.streams_for_rank(8, 1, 8) == [1]
    assert sharding.streams_for_rank(3, 1, 2) == [1]
    with pytest.raises(ValueError):
        sharding.streams_for_rank(8, 2, 2)


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mine = sharding.streams_for_rank(5, rank, world)
    frames_local = 32 * len(mine)                  # every stream contributes one 32-frame step
    secs_local = 0.010 * (rank + 1)                # rank 1 is the slow one
    frames, secs, fps = sharding.reduce_throughput(frames_local, secs_local, dist)
    dist.barrier()
    q.put((rank, mine, frames, secs, fps))
    dist.destroy_process_group()


def test_world2_gloo_reduction():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs: p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs: p.join(timeout=60)
    assert res[0][1] == [0, 2, 4] and res[1][1] == [1, 3]
    for _, _, frames, secs, fps in res:
        assert frames == 160 and abs(secs - 0.020) < 1e-9 and abs(fps - 8000.0) < 1e-3


def test_single_process_passthrough():
    assert sharding.reduce_throughput(64, 0.5) == (64, 0.5, 128.0)
