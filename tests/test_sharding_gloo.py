"""N > 1 host logic on CPU: world_size-2 (and 3) gloo groups exercise the block sharding and the
compact-result gather that the multi-GPU bench uses with NCCL."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from pyperiod_b200 import sharding


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, total, k, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        lo, hi = sharding.shard_bounds(total, world, rank)
        # deterministic fake compact results: a function of the global window index only
        idx = torch.arange(lo, hi, dtype=torch.int64)
        periods = ((idx[:, None] * 7 + torch.arange(k)[None, :]) % 1024 + 2).to(torch.int32)
        powers = (idx[:, None].double() + 1.0) / (torch.arange(k)[None, :].double() + 1.0)
        status = (idx % 5 == 0).to(torch.int32)
        out = sharding.gather_compact(periods, powers, status, total, dst=0)
        if rank == 0:
            ret["periods"], ret["powers"], ret["status"] = (t.numpy() for t in out)
        else:
            assert out is None
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,total", [(2, 10), (2, 7), (3, 8)])
def test_gather_compact_matches_unsharded(world, total):
    k = 4
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), total, k, ret), nprocs=world, join=True)
    idx = np.arange(total)
    want_p = ((idx[:, None] * 7 + np.arange(k)[None, :]) % 1024 + 2).astype(np.int32)
    want_w = (idx[:, None] + 1.0) / (np.arange(k)[None, :] + 1.0)
    assert np.array_equal(ret["periods"], want_p)
    assert np.array_equal(ret["powers"], want_w)
    assert np.array_equal(ret["status"], (idx % 5 == 0).astype(np.int32))


def test_shard_bounds_cover_everything_once():
    for total in (0, 1, 7, 8, 1_048_576):
        for world in (1, 2, 3, 8):
            seen = []
            for r in range(world):
                lo, hi = sharding.shard_bounds(total, world, r)
                assert 0 <= lo <= hi <= total
                seen += list(range(lo, hi)) if total < 100 else [(lo, hi)]
            if total < 100:
                assert seen == list(range(total))
            else:
                assert seen[0][0] == 0 and seen[-1][1] == total
                assert all(a[1] == b[0] for a, b in zip(seen, seen[1:]))


def test_stream_shard_halo():
    # 1M hop-512 windows of 4096 samples over 8 ranks: each rank needs its block plus a 3584-sample halo
    first, n_local, lo, hi = sharding.stream_shard(1_048_576, 4096, 512, 8, 3)
    assert (first, n_local) == (3 * 131_072, 131_072)
    assert lo == first * 512 and hi - lo == (n_local - 1) * 512 + 4096
    assert sharding.stream_shard(4, 4096, 512, 8, 7)[1] == 0
