"""Multi-GPU: windows are independent, so the batch is split into contiguous blocks, one per rank
(one process per GPU), with no data-path collective.  The only exchange is a gather of the compact
results -- periods u32[B_g, K], powers f64[B_g, K], status i32[B_g] (~124 B per window at K = 10) --
to one rank, done with torch.distributed (NCCL over NVLink on GPUs, gloo in the CPU tests).
Bases stay resident on the device that produced them.
"""
from __future__ import annotations

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


def gather_compact(periods: torch.Tensor, powers: torch.Tensor, status: torch.Tensor, total: int,
                   dst: int = 0, group=None):
    """Gather per-rank compact results (block-sharded by shard_bounds) to `dst`.

    Every rank passes its local (B_g, K) periods (int32 view of the uint32 values), (B_g, K) powers and
    (B_g,) status.  Returns (periods[total, K], powers[total, K], status[total]) on `dst`, None elsewhere.
    Ranks pad to the common block size so a single fixed-size gather per array suffices.
    """
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    per = -(-total // world)
    k = periods.shape[1]
    dev = periods.device

    def pad(t, shape, dtype):
        out = torch.zeros(shape, dtype=dtype, device=dev)
        out[: t.shape[0]] = t
        return out

    p_loc = pad(periods.view(torch.int32) if periods.dtype != torch.int32 else periods, (per, k), torch.int32)
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
