"""ABAB of a launch-time option on WHOLE frames (rasterizer + resolve/gather + net graph), C3: blocks of 5 frames per setting,
alternating, median frame time per setting.   python scripts/ab_frame_opt.py "raster_carveout=-1;raster_carveout=45,raster_stages=2" """
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from read_b200 import synth, ops, _lib as L
from read_b200.engine import UNetEngine

lib = L.load()
dev = torch.device("cuda", 0)
N, W, H = 10_000_000, 1920, 1088
store = ops.SortedPoints(torch.from_numpy(synth.street_scene(N)).to(dev))
tex = torch.rand((N, 8), device=dev)
eng = UNetEngine(synth.synth_state_dict(synth.SEED), 1, H, W, dev, precision="bf16")
pyr = ops.Pyramid(1, W, H, 4, dev)
pyr.clear()
mats = []
for t in range(8):
    proj, view = synth.camera_batch(W, H, [t * 5])
    mats.append(torch.from_numpy(synth.total_matrix(proj, view)).to(dev))
settings = [dict((kv.split("=")[0], int(kv.split("=")[1])) for kv in blk.split(",")) for blk in sys.argv[1].split(";")]


def frame(i):
    ops.raster_project_sorted(pyr, store, mats[i % 8])
    ops.pyramid_resolve_gather(tex, pyr, eng.inputs, L.FEAT_NHWC_BF16, reset_level0=True)
    eng.run()


for i in range(4):
    frame(i)
torch.cuda.synchronize()
ts = [[] for _ in settings]
rs = [[] for _ in settings]
for rep in range(6):
    for j, st in enumerate(settings):
        for k, v in st.items():
            L.check(lib.read_set_option(k.encode(), v))
        frame(0); torch.cuda.synchronize()
        a, b, c = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        a.record()
        for i in range(5):
            frame(i)
        b.record(); torch.cuda.synchronize()
        ts[j].append(a.elapsed_time(b) / 5)
        a.record(); ops.raster_project_sorted(pyr, store, mats[0]); c.record(); torch.cuda.synchronize()
        rs[j].append(a.elapsed_time(c) * 1e3)
        pyr.clear()
for st, t, r in zip(settings, ts, rs):
    print(st, "frame ms median", round(float(np.median(t[1:])), 4), " raster alone us (after a frame)", round(float(np.median(r[1:])), 1))
