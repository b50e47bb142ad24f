"""Time the rasterizer variants on the C3 frame (10M points, 1920x1088, level 0) and check that every variant produces
the bit-identical packed z-buffer.  python scripts/bench_raster_modes.py [sorted]"""
import os, sys, json
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from read_b200 import synth, ops, _lib as L
lib = L.load()
dev = torch.device("cuda", 0)
N, W, H = 10_000_000, 1920, 1088
xyz_np = synth.street_scene(N)
if len(sys.argv) > 1 and sys.argv[1].startswith("sort"):
    # EXPERIMENT: spatially sorted cloud (ids change, so only the timings are meaningful): how much do the scattered
    # z-buffer accesses gain from a warp's points landing in neighbouring pixels?
    def part1by2(v):
        v = v.astype(np.uint64) & np.uint64(0x1FFFFF)
        v = (v | (v << np.uint64(32))) & np.uint64(0x1F00000000FFFF)
        v = (v | (v << np.uint64(16))) & np.uint64(0x1F0000FF0000FF)
        v = (v | (v << np.uint64(8))) & np.uint64(0x100F00F00F00F00F)
        v = (v | (v << np.uint64(4))) & np.uint64(0x10C30C30C30C30C3)
        v = (v | (v << np.uint64(2))) & np.uint64(0x1249249249249249)
        return v
    lo, hi = xyz_np.min(0), xyz_np.max(0)
    cell = float(sys.argv[1][4:] or 0.25)                       # metres per grid cell
    q = np.floor((xyz_np - lo) / cell).astype(np.int64)
    code = part1by2(q[:, 0]) | (part1by2(q[:, 1]) << np.uint64(1)) | (part1by2(q[:, 2]) << np.uint64(2))
    order = np.argsort(code, kind="stable")
    xyz_np = np.ascontiguousarray(xyz_np[order])
    print(f"sorted by Morton code of {cell} m cells")
xyz = torch.from_numpy(xyz_np).to(dev)
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
setopt(raster_mode=2, raster_occupancy=0)
# the sorted store (real thing: original ids carried, bit-identical result)
store = ops.SortedPoints(torch.from_numpy(synth.street_scene(N)).to(dev))
setopt(raster_mode=0)
refs0 = []
x0 = torch.from_numpy(synth.street_scene(N)).to(dev)
for m in mats:
    pyr.clear(); ops.raster_project(pyr, x0, m); torch.cuda.synchronize(); refs0.append(pyr.buf.clone())
setopt(raster_mode=2)
for dedup, occ, run, nbr in ((0, 0, 0, 0), (0, 0, 0, 1), (0, 0, 16, 1), (0, 0, 32, 1), (0, 4, 0, 1), (1, 0, 0, 0)):
    if True:
        setopt(raster_dedup=dedup, raster_occupancy=occ, raster_run=run, raster_nbr_filter=nbr)
        bad = 0
        for m, r in zip(mats, refs0):
            pyr.clear(); ops.raster_project_sorted(pyr, store, m); torch.cuda.synchronize()
            bad += int((pyr.buf != r).sum())
        ts = []
        for _ in range(12):
            pyr.clear(); torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); ops.raster_project_sorted(pyr, store, mats[0]); b.record(); torch.cuda.synchronize()
            ts.append(a.elapsed_time(b) * 1e3)
        ts = sorted(ts[2:]); med = ts[len(ts) // 2]
        rows.append({"mode": "sorted", "dedup": dedup, "occ": occ, "run": run, "nbr_filter": nbr, "us_median": med, "us_best": ts[0],
                     "xyz_GBps": 12 * N / (med * 1e-6) / 1e9, "mismatches": bad})
        print(f"sorted store dedup {dedup} occ {occ} run {run} nbr {nbr}: median {med:7.1f} us  best {ts[0]:7.1f} us  "
              f"{12 * N / (med * 1e-6) / 1e9:7.1f} GB/s of xyz  mismatching keys vs unsorted render {bad}")
setopt(raster_dedup=0, raster_occupancy=0, raster_run=0, raster_nbr_filter=0)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(rows, open("gpurun_out/raster_modes.json", "w"), indent=1)
