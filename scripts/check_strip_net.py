"""torchrun check + timing of the strip-parallel refinement net: the frame stitched from `world` strips (halo exchange over
NVLink peer memory) must equal the single-GPU frame bit for bit; then latency per frame vs the single-GPU engine.
   python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 scripts/check_strip_net.py [H W N]"""
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from read_b200 import synth, ops, _lib as L, dist as rdist          # noqa: E402
from read_b200.engine import UNetEngine                              # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
H = int(sys.argv[1]) if len(sys.argv) > 1 else 1088
W = int(sys.argv[2]) if len(sys.argv) > 2 else 1920
N = int(sys.argv[3]) if len(sys.argv) > 3 else 10_000_000
sd = synth.synth_state_dict(synth.SEED)
xyz = torch.from_numpy(synth.street_scene(N)).to(dev)
full_store = ops.SortedPoints(xyz)
start, count = rdist.shard_range(N, rank, world)
shard = full_store.shard(start, count)
tex = torch.rand((N, 8), generator=torch.Generator().manual_seed(1)).to(dev)
fr = rdist.StripFrameRenderer(shard, tex, sd, W, H, dev)
res = {"world": world, "H": H, "W": W, "exchanges_per_frame": fr.eng.n_exchanges(), "launches_per_frame": fr.eng.n_launches()}
# reference: the single-GPU path on every rank (same scene, same pose)
pyr = ops.Pyramid(1, W, H, 4, dev)
eng1 = UNetEngine(sd, 1, H, W, dev, precision="bf16", use_graph=True)
ok = True
for pose in (7, 23):
    proj, view = synth.camera_batch(W, H, [pose])
    m = torch.from_numpy(synth.total_matrix(proj, view)).to(dev)
    got = fr.render(m).clone()
    pyr.clear()
    ops.raster_project_sorted(pyr, full_store, m)
    ops.pyramid_resolve_gather(tex, pyr, eng1.inputs, L.FEAT_NHWC_BF16)
    want = eng1.run()[0]
    torch.cuda.synchronize()
    same = bool(torch.equal(got, want))
    err = float((got - want).abs().max())
    ok = ok and same
    res[f"pose{pose}"] = {"equal": same, "max_abs_diff": err}
flag = torch.tensor([1 if ok else 0], device=dev)
dist.all_reduce(flag, op=dist.ReduceOp.MIN)
res["identical_on_every_rank"] = bool(flag.item())


def timed(fn, reps=20):
    dist.barrier(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record(); torch.cuda.synchronize()
    t = torch.tensor([a.elapsed_time(b) / reps], device=dev, dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


proj, view = synth.camera_batch(W, H, [9])
m = torch.from_numpy(synth.total_matrix(proj, view)).to(dev)


def single():
    ops.raster_project_sorted(pyr, full_store, m)
    ops.pyramid_resolve_gather(tex, pyr, eng1.inputs, L.FEAT_NHWC_BF16, reset_level0=True)
    eng1.run()


pyr.clear()
for _ in range(3):
    single(); fr.render(m)
res["ms_per_frame_single_gpu"] = timed(single)
res["ms_per_frame_strip_parallel"] = timed(lambda: fr.render(m))
res["ms_net_only_single"] = timed(lambda: eng1.run())
res["ms_net_only_strip"] = timed(lambda: fr.eng.run())
res["latency_speedup"] = res["ms_per_frame_single_gpu"] / res["ms_per_frame_strip_parallel"]
if rank == 0:
    print(json.dumps(res), flush=True)
dist.destroy_process_group()
sys.exit(0 if flag.item() else 1)
