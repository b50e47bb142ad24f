"""A/B the conv tuning options on the C3 net (1920x1088): whole-graph time + a few single layers, outputs compared bit for bit.
   python scripts/ab_conv.py [out.json]"""
import json
import sys
import os
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from read_b200 import synth, _lib as L
from read_b200.engine import UNetEngine

lib = L.load()
dev = torch.device("cuda", 0)
sd = synth.synth_state_dict(synth.SEED)
H, W = 1088, 1920
g = torch.Generator().manual_seed(3)
feats = [torch.rand((1, 8, H >> l, W >> l), generator=g) for l in range(4)]
PICK = ["Encoder.0.layers.0.main.0", "Encoder.0.layers.0.main.1", "Encoder.1.layers.0.main.0", "Encoder.1.layers.0.main.1",
        "Encoder.2.layers.0.main.0", "Encoder.3.layers.0.main.0", "Convs.2", "AFFs.0.conv.0", "feat_extract.5", "SCM2.conv"]


def setopt(**kw):
    for k, v in kw.items():
        L.check(lib.read_set_option(k.encode(), int(v)))


def tm(fn, reps=20):
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


rows, ref = [], None
combos = [dict(tc_mt=1, tc_role_rot=0, tc_pdl=0), dict(tc_mt=1, tc_role_rot=1, tc_pdl=0), dict(tc_mt=0, tc_role_rot=0, tc_pdl=0),
          dict(tc_mt=0, tc_role_rot=1, tc_pdl=0), dict(tc_mt=0, tc_role_rot=1, tc_pdl=1), dict(tc_mt=1, tc_role_rot=0, tc_pdl=1),
          dict(tc_mt=2, tc_role_rot=1, tc_pdl=1)]
for c in combos:
    setopt(**c)
    eng = UNetEngine(sd, 1, H, W, dev, precision="bf16", use_graph=True)
    eng.set_inputs_nchw([f.to(dev) for f in feats])
    for _ in range(3):
        eng.run()
    t = tm(eng.run)
    out = eng.output.clone()
    if ref is None:
        ref = out
    same = bool(torch.equal(out, ref))
    err = float((out - ref).abs().max())
    sp = L.stream_ptr()
    layers = {}
    for ly in eng.ops:
        if ly.name in PICK:
            for _ in range(2):
                eng.launch_op(ly, sp)
            layers[ly.name] = round(tm(lambda: eng.launch_op(ly, sp), 10) * 1e3, 1)
    row = dict(c, graph_ms=round(t, 4), equal_to_first=same, max_abs_diff=err, layers_us=layers)
    rows.append(row)
    print(json.dumps(row), flush=True)
    del eng
    torch.cuda.empty_cache()
setopt(tc_mt=0, tc_role_rot=1, tc_pdl=1)
if len(sys.argv) > 1:
    json.dump(rows, open(sys.argv[1], "w"), indent=1)
