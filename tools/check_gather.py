"""pp_gather on N GPUs: torchrun --nproc-per-node N tools/check_gather.py  (each rank contributes a distinct block;
rank 0 checks the packed NCCL gather against torch.distributed.all_gather)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
from pyperiod_b200 import _lib, sharding
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
total = 1000 * world - 7
lo, hi = sharding.shard_bounds(total, world, rank)
rng = np.random.default_rng(100 + rank)
per = torch.from_numpy(rng.integers(2, 1000, (hi - lo, 10)).astype(np.int32)).cuda()
pw = torch.from_numpy(rng.random((hi - lo, 10))).cuda()
st = torch.full((hi - lo,), rank, dtype=torch.int32, device="cuda")
for rep in range(3):
    out = sharding.gather_compact(per, pw, st, total, dst=0)
torch.cuda.synchronize()
blk = -(-total // world)
pads = [torch.zeros((blk, 10), dtype=torch.float64, device="cuda") for _ in range(world)]
mine = torch.zeros((blk, 10), dtype=torch.float64, device="cuda"); mine[: hi - lo] = pw
dist.all_gather(pads, mine)
if rank == 0:
    p_all, w_all, s_all = out
    ref = torch.cat(pads)[:total]
    ok = torch.equal(w_all, ref) and all(int(s_all[min(r * blk, total - 1)]) == r for r in range(world))
    print("pp_gather", "OK" if ok else "MISMATCH", "world", world, "nccl", _lib.load().pp_comm_version())
sharding.Comm.destroy_all()
dist.destroy_process_group()
