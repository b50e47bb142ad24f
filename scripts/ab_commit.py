"""ABAB timing of the tcgen05.commit experiments on single layers of the C3 net: does the number of commits
per unit of MMA work bound the kernels?   python scripts/ab_commit.py [out.json]"""
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
        "Decoder.3.layers.1.main.0", "Decoder.2.layers.1.main.1", "FAM2.merge", "AFFs.0.conv.1", "AFFs.1.conv.1", "Encoder.2.layers.0.main.0"]


def setopt(**kw):
    for k, v in kw.items():
        L.check(lib.read_set_option(k.encode(), int(v)))


def mk(**opts):
    base = dict(tc_mt=1, tc_commit_late=0, tc_merge_done=1, tc_bpair=0, tc_probe=0, tc_pair=0, tc_role_rot=1, tc_pdl=0)
    base.update(opts)
    setopt(**base)
    e = UNetEngine(sd, 1, H, W, dev, precision="bf16", use_graph=False)
    e.set_inputs_nchw([f.to(dev) for f in feats])
    e.run()
    torch.cuda.synchronize()
    return e


def t1(fn):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); fn(); b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b)


variants = [("no merge", mk(tc_merge_done=0)), ("merge", mk()), ("pair", mk(tc_pair=2))]
setopt(tc_pdl=0)
sp = L.stream_ptr()
ref = variants[0][1].output.clone()
out = {}
for k, e in variants:
    out["equal " + k] = bool(torch.equal(e.output, ref))
print({k: v for k, v in out.items()}, flush=True)
for name in PICK:
    ts = {k: [] for k, _ in variants}
    for rep in range(9):
        for k, e in variants:
            ly = next(l for l in e.ops if l.name == name)
            ts[k].append(t1(lambda: e.launch_op(ly, sp)) * 1e3)
    row = {k: round(float(np.median(v[2:])), 1) for k, v in ts.items()}
    out[name] = row
    print(f"{name:30s}", row, flush=True)
setopt(tc_mt=1, tc_commit_late=0, tc_merge_done=1, tc_bpair=0, tc_probe=0, tc_pair=1, tc_pdl=1)
if len(sys.argv) > 1:
    json.dump(out, open(sys.argv[1], "w"), indent=1)
