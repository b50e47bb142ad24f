"""Where does the CTA-pair conv kernel's epilogue time go?  ABAB timing of C=64 layers on the READ_DIAG build with the epilogue's
global stores (tc_debug 2) and / or its residual loads (tc_debug 32) switched off.
   READ_B200_LIB=read_b200/libread_b200_diag.so python scripts/ab_pair_dbg.py"""
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
PICK = sys.argv[1].split(",") if len(sys.argv) > 1 else ["Encoder.1.layers.0.main.0", "Encoder.1.layers.0.main.1", "Decoder.1.layers.1.main.1",
                                                          "Encoder.0.layers.0.main.0", "Encoder.0.layers.0.main.1"]
MODES = sys.argv[2].split(",") if len(sys.argv) > 2 else ["tc_debug=0", "tc_debug=2", "tc_debug=32", "tc_debug=34"]   # option=value, toggled per launch


def setopt(**kw):
    for k, v in kw.items():
        L.check(lib.read_set_option(k.encode(), int(v)))


def t1(fn):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); fn(); b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b)


setopt(tc_pdl=0)
e = UNetEngine(sd, 1, H, W, dev, precision="bf16", use_graph=False)
e.set_inputs_nchw([f.to(dev) for f in feats])
e.run(); torch.cuda.synchronize()
sp = L.stream_ptr()
for name in PICK:
    ly = next(l for l in e.ops if l.name == name)
    ts = {m: [] for m in MODES}
    for rep in range(9):
        for m in MODES:
            setopt(**{m.split("=")[0]: int(m.split("=")[1])})
            ts[m].append(t1(lambda: e.launch_op(ly, sp)) * 1e3)
    setopt(**{m.split("=")[0]: 0 if m.startswith("tc_debug") else 1 for m in MODES})
    print(f"{name:30s}", {m: round(float(np.median(v[2:])), 1) for m, v in ts.items()}, flush=True)
