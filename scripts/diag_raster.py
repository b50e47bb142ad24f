"""Diagnose raster variants at full size: compare {pipelined, plain} x {bulk TMA, LDG} renders of the same frame."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from read_b200 import synth, ops, _lib as L
lib = L.load()
dev = torch.device("cuda", 0)
N, W, H = 10_000_000, 1920, 1088
xyz = torch.from_numpy(synth.street_scene(N)).to(dev)
proj, view = synth.camera_batch(W, H, [7])
m = torch.from_numpy(synth.total_matrix(proj, view)).to(dev)
def render(pipe, bulk, w=W, h=H):
    L.check(lib.read_set_option(b"raster_pipelined", pipe)); L.check(lib.read_set_option(b"raster_bulk_tma", bulk))
    p = ops.Pyramid(1, w, h, 1, dev); p.clear(); ops.raster_project(p, xyz, m); torch.cuda.synchronize()
    return p.buf.clone()
ref = render(0, 0)
for pipe in (0, 1):
    for bulk in (0, 1):
        for rep in range(4):
            b = render(pipe, bulk)
            nd = int((b != ref).sum())
            worse = int((b > ref).sum())
            print(f"pipelined={pipe} bulk={bulk} rep={rep}: differing keys {nd} (of {ref.numel()}), of which larger-than-ref {worse}")
