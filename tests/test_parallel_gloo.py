"""world_size-2 gloo test (CPU) of the frame-sharding host logic: contiguous ranges cover every frame
exactly once and the padded all_gather returns frames in order (speech-to-video-mpp_b200/parallel.py)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from s2v_b200 import parallel


@pytest.mark.parametrize("n,world", [(0, 2), (1, 2), (122, 1), (122, 2), (1497, 8), (14997, 8), (7, 8)])
def test_shard_ranges_cover(n, world):
    seen = []
    for r in range(world):
        lo, hi = parallel.shard_range(n, r, world)
        assert 0 <= lo <= hi <= n
        seen += list(range(lo, hi))
    assert seen == list(range(n))
    sizes = [parallel.shard_range(n, r, world)[1] - parallel.shard_range(n, r, world)[0] for r in range(world)]
    assert max(sizes) == (-(-n // world) if n else 0)


def _worker(rank, world, port, n, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        lo, hi = parallel.shard_range(n, rank, world)
        # "generated frames": frame i is filled with the value i (stands in for the per-rank pipeline output)
        local = torch.arange(lo, hi, dtype=torch.float32).reshape(-1, 1, 1, 1).expand(hi - lo, 3, 4, 4).contiguous()
        full = parallel.gather_frames(local, n)
        ok = full.shape == (n, 3, 4, 4) and bool((full[:, 0, 0, 0] == torch.arange(n, dtype=torch.float32)).all())
        q.put((rank, ok))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n", [122, 5, 2])
def test_gather_frames_gloo_world2(n):
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, True), (1, True)]
