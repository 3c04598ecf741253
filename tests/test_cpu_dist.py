"""CPU suite, part 3: the N>1 sharding / all-gather plumbing with world_size 2 on the gloo backend."""
import os
import socket

import pytest

from directxtex_b200 import dist as D


def test_shard_ranges_cover_everything():
    for n in (1, 7, 8, 1024, 1000):
        for world in (1, 2, 3, 4, 8):
            seen = []
            for r in range(world):
                lo, hi = D.shard_range(n, world, r)
                assert 0 <= lo <= hi <= n
                seen += list(range(lo, hi))
            assert seen == list(range(n))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_items, item_bytes, q):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = D.shard_range(n_items, world, rank)
    # "packed blocks" of item i = bytes derived from i
    local = torch.cat([torch.full((item_bytes,), i % 251, dtype=torch.uint8) for i in range(lo, hi)]) if hi > lo else torch.zeros(0, dtype=torch.uint8)
    out = D.all_gather_blocks(local, n_items, item_bytes, world, rank, dist, torch)
    want = torch.cat([torch.full((item_bytes,), i % 251, dtype=torch.uint8) for i in range(n_items)])
    q.put((rank, bool(torch.equal(out[: n_items * item_bytes], want))))
    dist.destroy_process_group()


@pytest.mark.parametrize("n_items", [8, 7])
def test_all_gather_blocks_world2_gloo(n_items):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_items, 48, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, True), (1, True)]
