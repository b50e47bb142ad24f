"""Where the streaming rasterizer's time goes at C3 (READ_DIAG build): project only / + early-z reads / full, interleaved.
   READ_B200_LIB=read_b200/libread_b200_diag.so python scripts/raster_breakdown.py"""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from read_b200 import synth, ops, _lib as L
lib = L.load()
dev = torch.device("cuda", 0)
N, W, H = 10_000_000, 1920, 1088
store = ops.SortedPoints(torch.from_numpy(synth.street_scene(N)).to(dev))
proj, view = synth.camera_batch(W, H, [7])
m = torch.from_numpy(synth.total_matrix(proj, view)).to(dev)
pyr = ops.Pyramid(1, W, H, 1, dev)
ts = {4: [], 5: [], 2: []}
for rep in range(10):
    for mode in (4, 5, 2):
        L.check(lib.read_set_option(b"raster_mode", mode))
        pyr.clear(); torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); ops.raster_project_sorted(pyr, store, m); b.record(); torch.cuda.synchronize()
        ts[mode].append(a.elapsed_time(b) * 1e3)
L.check(lib.read_set_option(b"raster_mode", 2))
vis = int((pyr.buf != 0x7FFFFFFFFFFFFFFF).sum())
for mode, name in ((4, "project + cull only"), (5, "+ early-z reads"), (2, "full (reads + atomics)")):
    v = sorted(ts[mode][2:])
    print(f"{name:26s}: median {v[len(v)//2]:6.1f} us  best {v[0]:6.1f} us")
print("covered pixels", vis)
