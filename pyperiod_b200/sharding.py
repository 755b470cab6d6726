"""Multi-GPU: windows are independent, so the batch is split into contiguous blocks, one per rank
(one process per GPU), with no data-path collective.  The only exchange is a gather of the compact
results -- periods u32[B_g, K], powers f64[B_g, K], status i32[B_g] (~124 B per window at K = 10) --
to one rank.  On GPUs it is ONE pp_gather call (include/pyperiod_b200.h: ncclGather over NVLink, enqueued on
the compute stream behind the last kernel, no host synchronisation) of one packed buffer per rank; the
communicator is bootstrapped through torch.distributed (rank 0's NCCL unique id is broadcast).  CPU tensors
(the gloo tests) go through torch.distributed.gather.  Bases stay resident on the device that produced them.
"""
from __future__ import annotations

import ctypes as C

import torch
import torch.distributed as dist


def shard_bounds(total: int, world: int, rank: int) -> tuple[int, int]:
    """Contiguous block [lo, hi) of `total` windows owned by `rank`: blocks of ceil(total / world)."""
    per = -(-total // world)
    lo = min(rank * per, total)
    return lo, min(lo + per, total)


def stream_shard(n_windows: int, n: int, hop: int, world: int, rank: int) -> tuple[int, int, int, int]:
    """Sample range of a hop-framed stream needed by `rank`: (first_window, n_local, sample_lo, sample_hi).

    The range includes the (n - hop)-sample halo the rank's last window reads past its block.
    """
    lo, hi = shard_bounds(n_windows, world, rank)
    if hi <= lo:
        return lo, 0, lo * hop, lo * hop
    return lo, hi - lo, lo * hop, (hi - 1) * hop + n


class Comm:
    """NCCL communicator owned by the C library (pp_comm_*), one per process group."""
    _by_group: dict = {}

    def __init__(self, group=None):
        from . import _lib
        self.lib = _lib.load()
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        dev = torch.device("cuda", torch.cuda.current_device())
        uid = torch.zeros(128, dtype=torch.uint8)
        if self.rank == 0:
            buf = (C.c_char * 128)()
            _lib.check(self.lib.pp_comm_unique_id(buf), "pp_comm_unique_id")
            uid = torch.frombuffer(bytearray(buf.raw), dtype=torch.uint8).clone()
        uid = uid.to(dev)
        dist.broadcast(uid, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
        raw = bytes(uid.cpu().numpy().tobytes())
        self.handle = C.c_void_p()
        _lib.check(self.lib.pp_comm_init(C.byref(self.handle), self.world, self.rank, raw), "pp_comm_init")

    @classmethod
    def for_group(cls, group=None) -> "Comm":
        key = id(group) if group is not None else 0
        if key not in cls._by_group:
            cls._by_group[key] = cls(group)
        return cls._by_group[key]

    def gather(self, send: torch.Tensor, recv, root: int):
        """Every rank sends its contiguous `send` bytes; `recv` (root: world * send.nbytes) gets them in rank order.
        Enqueued on the current CUDA stream."""
        from . import _lib
        stream = C.c_void_p(torch.cuda.current_stream(send.device).cuda_stream)
        _lib.check(self.lib.pp_gather(self.handle, C.c_void_p(send.data_ptr()),
                                      C.c_void_p(0 if recv is None else recv.data_ptr()),
                                      send.numel() * send.element_size(), int(root), stream), "pp_gather")

    def destroy(self):
        if self.handle:
            self.lib.pp_comm_destroy(self.handle)
            self.handle = C.c_void_p()

    @classmethod
    def destroy_all(cls):
        for c in cls._by_group.values():
            c.destroy()
        cls._by_group.clear()


def gather_compact(periods: torch.Tensor, powers: torch.Tensor, status: torch.Tensor, total: int,
                   dst: int = 0, group=None):
    """Gather per-rank compact results (block-sharded by shard_bounds) to `dst`.

    Every rank passes its local (B_g, K) periods (int32 view of the uint32 values), (B_g, K) powers and
    (B_g,) status.  Returns (periods[total, K], powers[total, K], status[total]) on `dst`, None elsewhere.
    Ranks pad to the common block size, so one fixed-size exchange suffices: on CUDA tensors the three arrays are
    packed into one buffer per rank and travel in ONE pp_gather (NCCL) on the current stream.
    """
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    per = -(-total // world)
    k = periods.shape[1]
    dev = periods.device
    p32 = periods.view(torch.int32) if periods.dtype != torch.int32 else periods

    if dev.type == "cuda":
        # packed layout per rank: powers f64[per, k] | periods i32[per, k] | status i32[per]   (8-byte aligned pieces)
        nb_w, nb_p, nb_s = per * k * 8, per * k * 4, per * 4
        nb = nb_w + ((nb_p + nb_s + 7) // 8) * 8
        send = torch.zeros(nb, dtype=torch.uint8, device=dev)
        b = p32.shape[0]
        send[: nb_w].view(torch.float64).view(per, k)[:b] = powers
        send[nb_w: nb_w + nb_p].view(torch.int32).view(per, k)[:b] = p32
        send[nb_w + nb_p: nb_w + nb_p + nb_s].view(torch.int32)[:b] = status
        recv = torch.empty(world * nb, dtype=torch.uint8, device=dev) if rank == dst else None
        Comm.for_group(group).gather(send, recv, dst)
        if rank != dst:
            return None
        r = recv.view(world, nb)
        w_all = r[:, :nb_w].contiguous().view(torch.float64).view(world * per, k)
        p_all = r[:, nb_w: nb_w + nb_p].contiguous().view(torch.int32).view(world * per, k)
        s_all = r[:, nb_w + nb_p: nb_w + nb_p + nb_s].contiguous().view(torch.int32).view(world * per)
        return p_all[:total], w_all[:total], s_all[:total]

    def pad(t, shape, dtype):
        out = torch.zeros(shape, dtype=dtype, device=dev)
        out[: t.shape[0]] = t
        return out

    p_loc = pad(p32, (per, k), torch.int32)
    w_loc = pad(powers, (per, k), torch.float64)
    s_loc = pad(status, (per,), torch.int32)
    if rank == dst:
        p_all = [torch.empty_like(p_loc) for _ in range(world)]
        w_all = [torch.empty_like(w_loc) for _ in range(world)]
        s_all = [torch.empty_like(s_loc) for _ in range(world)]
    else:
        p_all = w_all = s_all = None
    dist.gather(p_loc, p_all, dst=dst, group=group)
    dist.gather(w_loc, w_all, dst=dst, group=group)
    dist.gather(s_loc, s_all, dst=dst, group=group)
    if rank != dst:
        return None
    return (torch.cat(p_all)[:total], torch.cat(w_all)[:total], torch.cat(s_all)[:total])
