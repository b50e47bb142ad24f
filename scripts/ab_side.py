"""ABAB: whole-net CUDA-graph replay with the side chain (SCM1 / SCM0 blocks on a second stream, SIDE_CTAS SMs) on and off.
   Result on a B200 (call r3g): off 5.84 ms, side16 6.31, side24 6.12, side32 6.06 - slower; the option stays off.
   python scripts/ab_side.py [side_ctas,side_ctas,...]"""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from read_b200 import synth, _lib as L
from read_b200.engine import UNetEngine

dev = torch.device("cuda", 0)
sd = synth.synth_state_dict(synth.SEED)
H, W = 1088, 1920
g = torch.Generator().manual_seed(3)
feats = [torch.rand((1, 8, H >> l, W >> l), generator=g) for l in range(4)]


def mk(side, ctas=24):
    os.environ["READ_B200_SIDE_CHAIN"] = "1" if side else "0"
    UNetEngine.SIDE_CTAS = ctas
    e = UNetEngine(sd, 1, H, W, dev, precision="bf16", use_graph=True)
    e.set_inputs_nchw([f.to(dev) for f in feats])
    for _ in range(3):
        e.run()
    torch.cuda.synchronize()
    return e


def t1(fn):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); fn(); b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b)


variants = [("off", mk(False))] + [(f"side{c}", mk(True, int(c))) for c in (sys.argv[1].split(",") if len(sys.argv) > 1 else ["24"])]
ref = variants[0][1].output.clone()
print({k: bool(torch.equal(e.output, ref)) for k, e in variants}, flush=True)
ts = {k: [] for k, _ in variants}
for rep in range(15):
    for k, e in variants:
        ts[k].append(t1(e.run))
print({k: round(float(np.median(v[3:])), 4) for k, v in ts.items()}, "ms per net replay")
