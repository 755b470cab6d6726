"""pp_comm_* / pp_gather (include/pyperiod_b200.h): the NCCL gather of compact results owned by the C library.
World size 1 here (one GPU per test box); tools/check_gather.py is the same check under torchrun on N GPUs."""
import os
import socket

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def test_pp_gather_world_one():
    import torch
    import torch.distributed as dist
    from pyperiod_b200 import _lib, sharding
    lib = _lib.load()
    assert lib.pp_comm_version() >= 20000
    torch.cuda.set_device(0)
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ["MASTER_PORT"] = str(_free_port())
    dist.init_process_group("nccl", rank=0, world_size=1, device_id=torch.device("cuda", 0))
    try:
        rng = np.random.default_rng(3)
        per = torch.from_numpy(rng.integers(2, 1000, (37, 10)).astype(np.int32)).cuda()
        pw = torch.from_numpy(rng.random((37, 10))).cuda()
        st = torch.from_numpy(rng.integers(0, 3, 37).astype(np.int32)).cuda()
        p2, w2, s2 = sharding.gather_compact(per, pw, st, 37, dst=0)
        torch.cuda.synchronize()
        assert torch.equal(p2, per) and torch.equal(w2, pw) and torch.equal(s2, st)
        # raw entry point: bytes in, bytes out
        comm = sharding.Comm.for_group(None)
        send = torch.arange(1000, dtype=torch.uint8, device="cuda")
        recv = torch.zeros(1000, dtype=torch.uint8, device="cuda")
        comm.gather(send, recv, 0)
        torch.cuda.synchronize()
        assert torch.equal(send, recv)
    finally:
        sharding.Comm.destroy_all()
        dist.destroy_process_group()
