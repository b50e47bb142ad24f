"""torchrun check: point-sharded rendering + ONE NCCL min-reduce == single-GPU rendering, key for key, on every rank."""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from read_b200 import synth, ops, dist as rdist          # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
N, W, H, L = 3_000_000, 1920, 1088, 4
xyz = synth.street_scene(N)
proj, view = synth.camera_batch(W, H, list(range(world)))
m = torch.from_numpy(synth.total_matrix(proj, view)).to(dev)
full = ops.Pyramid(world, W, H, L, dev)
full.clear()
ops.raster_project(full, torch.from_numpy(xyz).to(dev), m)
start, count = rdist.shard_range(N, rank, world)
pyr = ops.Pyramid(world, W, H, L, dev)
rdist.render_sharded(pyr, torch.from_numpy(xyz[start:start + count]).to(dev), start, m)
torch.cuda.synchronize()
ok = torch.equal(pyr.buf, full.buf)
# the same with spatial tiles: rank r holds the r-th contiguous range of the Morton-sorted store
store = ops.SortedPoints(torch.from_numpy(xyz).to(dev)).shard(start, count)
pyr2 = ops.Pyramid(world, W, H, L, dev)
rdist.render_sharded(pyr2, store, 0, m)
torch.cuda.synchronize()
ok = ok and torch.equal(pyr2.buf, full.buf)
flag = torch.tensor([1 if ok else 0], device=dev)
dist.all_reduce(flag, op=dist.ReduceOp.MIN)
if rank == 0:
    print(f"sharded render over {world} ranks identical to single-GPU render on every rank: {bool(flag.item())}")
dist.destroy_process_group()
sys.exit(0 if flag.item() else 1)
