"""World-size-2 gloo test of the multi-GPU host logic (view sharding + composite gather), CPU only."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from glimpsw_b200 import sharding


def test_view_deal():
    assert sharding.views_for_rank(64, 3, 8) == list(range(3, 64, 8))
    assert sharding.views_for_rank(5, 1, 2) == [1, 3]
    assert sharding.rounds(5, 2) == 3 and sharding.rounds(64, 8) == 8
    got = sorted(v for r in range(4) for v in sharding.views_for_rank(10, r, 4))
    assert got == list(range(10))


def _worker(rank, world, port, num_views, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)

    def render_view(v):   # stand-in for the GPU frame: a deterministic "image" per view
        return torch.full((8, 12), v * 7 + 1, dtype=torch.int32)

    out = sharding.render_views_sharded(num_views, render_view, rank, world)
    if rank == 0:
        q.put([int(o[0, 0]) for o in out])
    else:
        assert out is None
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("num_views", [4, 5])
def test_gather_world2_gloo(num_views):
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, num_views, q)) for r in range(2)]
    for p in procs:
        p.start()
    result = q.get(timeout=120)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert result == [v * 7 + 1 for v in range(num_views)]
