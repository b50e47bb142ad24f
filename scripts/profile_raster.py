"""A few frames of the C3 rasterizer + resolve/gather only (for ncu).  python scripts/profile_raster.py"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from read_b200 import synth, ops, _lib as L
dev = torch.device("cuda", 0)
N, W, H = 10_000_000, 1920, 1088
store = ops.SortedPoints(torch.from_numpy(synth.street_scene(N)).to(dev))
tex = torch.rand((N, 8), device=dev)
proj, view = synth.camera_batch(W, H, [7])
m = torch.from_numpy(synth.total_matrix(proj, view)).to(dev)
pyr = ops.Pyramid(1, W, H, 4, dev)
outs = [torch.empty((1, H >> l, W >> l, 8), dtype=torch.bfloat16, device=dev) for l in range(4)]
pyr.clear()
for _ in range(3):
    ops.raster_project_sorted(pyr, store, m)
    ops.pyramid_resolve_gather(tex, pyr, outs, L.FEAT_NHWC_BF16, reset_level0=True)
torch.cuda.synchronize()
print("done")
