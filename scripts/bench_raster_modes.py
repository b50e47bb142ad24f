"""Time the rasterizer variants on the C3 frame (10M points, 1920x1088, level 0) and check that every variant produces
the bit-identical packed z-buffer.  python scripts/bench_raster_modes.py [sorted]"""
import os, sys, json
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from read_b200 import synth, ops, _lib as L
lib = L.load()
dev = torch.device("cuda", 0)
N, W, H = 10_000_000, 1920, 1088
xyz = torch.from_numpy(synth.street_scene(N)).to(dev)
poses = [7, 23, 40]
mats = []
for t in poses:
    proj, view = synth.camera_batch(W, H, [t])
    mats.append(torch.from_numpy(synth.total_matrix(proj, view)).to(dev))
pyr = ops.Pyramid(1, W, H, 1, dev)

def setopt(**kw):
    for k, v in kw.items():
        L.check(lib.read_set_option(k.encode(), v))

def render(m):
    pyr.clear(); ops.raster_project(pyr, xyz, m); torch.cuda.synchronize()
    return pyr.buf.clone()

def timeit(m, reps=12):
    ts = []
    for _ in range(reps):
        pyr.clear(); torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); ops.raster_project(pyr, xyz, m); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1e3)
    ts = sorted(ts[2:])
    return ts[len(ts) // 2], ts[0]

setopt(raster_mode=0)
refs = [render(m) for m in mats]
vis = int((refs[0] != 0x7FFFFFFFFFFFFFFF).sum())
print(f"covered pixels (pose {poses[0]}): {vis} of {W*H}")
rows = []
for mode, occs in ((0, (0,)), (1, (0,)), (2, (0, 2)), (3, (0,)), (4, (0,)), (5, (0,))):
    for occ in occs:
        setopt(raster_mode=mode, raster_occupancy=occ)
        bad = sum(int((render(m) != r).sum()) for m, r in zip(mats, refs))
        med, best = timeit(mats[0])
        gbs = 12 * N / (med * 1e-6) / 1e9
        rows.append({"mode": mode, "occ": occ, "us_median": med, "us_best": best, "xyz_GBps": gbs, "mismatches": bad})
        print(f"mode {mode} occ {occ}: median {med:7.1f} us  best {best:7.1f} us  {gbs:7.1f} GB/s of xyz  mismatching keys {bad}")
setopt(raster_mode=0, raster_occupancy=0)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(rows, open("gpurun_out/raster_modes.json", "w"), indent=1)
