"""Interleaved (ABAB) timing of conv tuning options on single layers of the C3 net and on the whole CUDA graph, so that clock /
power drift cancels.   python scripts/ab_layers.py [out.json]"""
import json, os, sys
import numpy as np, torch
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
        "Encoder.2.layers.0.main.0", "Encoder.3.layers.0.main.1", "Convs.2", "AFFs.0.conv.0", "AFFs.0.conv.1", "feat_extract.5",
        "SCM2.conv", "FAM2.merge", "Decoder.3.layers.0.main.0"]


def setopt(**kw):
    for k, v in kw.items():
        L.check(lib.read_set_option(k.encode(), int(v)))


def mk(mt, graph, rot=1, pdl=1):
    setopt(tc_mt=mt, tc_role_rot=rot, tc_pdl=pdl)
    e = UNetEngine(sd, 1, H, W, dev, precision="bf16", use_graph=graph)
    e.set_inputs_nchw([f.to(dev) for f in feats])
    for _ in range(2):
        e.run()
    torch.cuda.synchronize()
    return e


def t1(fn):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); fn(); b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b)


out = {}
# ---- single layers, eager, pdl off
e1, e4 = mk(1, False), mk(0, False)
sp = L.stream_ptr()
variants = [("mt1 rot0", e1, 0), ("mt1 rot1", e1, 1), ("auto rot0", e4, 0), ("auto rot1", e4, 1)]
setopt(tc_pdl=0)
for name in PICK:
    ts = {v[0]: [] for v in variants}
    for rep in range(9):
        for vn, e, rot in variants:
            ly = next(l for l in e.ops if l.name == name)
            setopt(tc_role_rot=rot)
            ts[vn].append(t1(lambda: e.launch_op(ly, sp)) * 1e3)
    row = {k: round(float(np.median(v[2:])), 1) for k, v in ts.items()}
    out[name] = row
    print(f"{name:30s}", row, flush=True)
assert torch.equal(e1.output, e4.output), "mt=1 and mt=auto outputs differ"
del e1, e4
torch.cuda.empty_cache()
# ---- whole graph
graphs = [("mt1 rot0 pdl0", mk(1, True, 0, 0)), ("auto rot1 pdl0", mk(0, True, 1, 0)), ("auto rot1 pdl1", mk(0, True, 1, 1)),
          ("auto rot0 pdl1", mk(0, True, 0, 1)), ("mt1 rot1 pdl1", mk(1, True, 1, 1))]
ts = {k: [] for k, _ in graphs}
for rep in range(8):
    for k, e in graphs:
        ts[k].append(t1(lambda: [e.run() for _ in range(5)]) / 5)
ref = graphs[0][1].output
for k, e in graphs:
    same = bool(torch.equal(e.output, ref))
    out["graph " + k] = {"ms": round(float(np.median(ts[k][2:])), 4), "equal": same}
    print("graph", k, out["graph " + k], flush=True)
setopt(tc_mt=0, tc_role_rot=1, tc_pdl=1)
if len(sys.argv) > 1:
    json.dump(out, open(sys.argv[1], "w"), indent=1)
