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
# throughput mode with look-ahead (read_b200.dist.ShardedFrameStream): 4 consecutive steps, each rank's frame == the single-GPU frame
# of its view, although the rasterizer + reduce-scatter of step s+1 run under the net of step s
from read_b200 import _lib as RL                           # noqa: E402
from read_b200.engine import UNetEngine                   # noqa: E402
sd = synth.synth_state_dict(synth.SEED)
tex_nd = torch.rand((N, 8), device=dev, generator=torch.Generator(device=dev).manual_seed(5))
eng = UNetEngine(sd, 1, H, W, dev, precision="bf16")
sfs = rdist.ShardedFrameStream(store, tex_nd, eng, W, H, L, RL.FEAT_NHWC_BF16)
steps = []
for s_ in range(5):
    p_, v_ = synth.camera_batch(W, H, [(3 * s_ + r_) % 64 for r_ in range(world)])
    steps.append(torch.from_numpy(synth.total_matrix(p_, v_)).to(dev))
full_store = ops.SortedPoints(torch.from_numpy(xyz).to(dev))
eng1 = UNetEngine(sd, 1, H, W, dev, precision="bf16")
one = ops.Pyramid(1, W, H, L, dev)
ok_stream = True
for s_ in range(4):
    out = sfs.step(steps[s_], steps[s_ + 1]).clone()
    one.clear()
    ops.raster_project_sorted(one, full_store, steps[s_][rank:rank + 1].contiguous())
    ops.pyramid_resolve_gather(tex_nd, one, eng1.inputs, RL.FEAT_NHWC_BF16)
    ref = eng1.run()
    torch.cuda.synchronize()
    ok_stream = ok_stream and bool(torch.equal(out, ref))
torch.cuda.current_stream().wait_stream(sfs.side)
torch.cuda.synchronize()
ok = ok and ok_stream
flag = torch.tensor([1 if ok else 0], device=dev)
dist.all_reduce(flag, op=dist.ReduceOp.MIN)
if rank == 0:
    print(f"sharded render over {world} ranks identical to single-GPU render on every rank (pyramids key for key, and 4 look-ahead "
          f"steps of ShardedFrameStream frame for frame): {bool(flag.item())}")
dist.destroy_process_group()
sys.exit(0 if flag.item() else 1)
