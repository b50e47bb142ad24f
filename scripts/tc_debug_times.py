"""Time single conv layers of the C3 net under the tc_debug diagnostic knobs (which role bounds the kernel?)."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from read_b200 import synth, _lib as L
from read_b200.engine import UNetEngine
W, H = 1920, 1088
dev = torch.device("cuda", 0); torch.cuda.set_device(0)
lib = L.load()
eng = UNetEngine(synth.synth_state_dict(synth.SEED), 1, H, W, dev, precision="bf16", use_graph=False)
for t in eng.inputs: t.uniform_(0, 1)
eng.run(); torch.cuda.synchronize()
by = {ly.name: ly for ly in eng.layers}
names = sys.argv[1].split(",") if len(sys.argv) > 1 else ["Encoder.0.layers.0.main.0", "Encoder.0.layers.0.main.1", "Encoder.1.layers.0.main.0", "Convs.2", "AFFs.0.conv.0", "feat_extract.0", "feat_extract.1", "AFFs.1.conv.0"]
sp = L.stream_ptr()
def t_layer(ly, reps=6):
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); eng.launch_op(ly, sp); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b) * 1e3)
    return float(np.median(ts[1:]))
for n in names:
    row = []
    gather = by[n].impl == L.CONV_TCGEN05_GATHER
    opt = b"tcg_debug" if gather else b"tc_debug"
    for dbg in ((0, 1, 2, 4, 1 | 2, 1 | 4, 2 | 4, 1 | 2 | 4) if gather else (0, 1, 8, 2, 4, 16, 1 | 2, 1 | 4, 2 | 4, 1 | 2 | 4)):
        L.check(lib.read_set_option(opt, dbg))
        row.append(f"dbg{dbg}={t_layer(by[n]):.1f}")
    L.check(lib.read_set_option(opt, 0))
    print(f"{n:28s} us: " + " ".join(row))
