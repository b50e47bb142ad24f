"""ABAB timing of a plan-time option: two engines built with OPT=0 / OPT=1, every conv layer timed alternately.
   python scripts/ab_opt.py tc_pair_wide [out.json]"""
import json, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from read_b200 import synth, _lib as L
from read_b200.engine import UNetEngine

OPT = sys.argv[1]
V0, V1 = (int(sys.argv[3]), int(sys.argv[4])) if len(sys.argv) > 4 else (0, 1)
lib = L.load()
dev = torch.device("cuda", 0)
sd = synth.synth_state_dict(synth.SEED)
H, W = 1088, 1920
g = torch.Generator().manual_seed(3)
feats = [torch.rand((1, 8, H >> l, W >> l), generator=g) for l in range(4)]


def setopt(**kw):
    for k, v in kw.items():
        L.check(lib.read_set_option(k.encode(), int(v)))


def mk(**opts):
    setopt(**opts)
    e = UNetEngine(sd, 1, H, W, dev, precision="bf16", use_graph=False)
    e.set_inputs_nchw([f.to(dev) for f in feats])
    e.run()
    torch.cuda.synchronize()
    return e


def t1(fn):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); fn(); b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b)


setopt(tc_pdl=0)
variants = [(f"{OPT}={V0}", mk(**{OPT: V0})), (f"{OPT}={V1}", mk(**{OPT: V1}))]
sp = L.stream_ptr()
out = {"equal": bool(torch.equal(variants[0][1].output, variants[1][1].output)),
       "max_abs_diff": float((variants[0][1].output.float() - variants[1][1].output.float()).abs().max())}
print(out, flush=True)
tot = {k: 0.0 for k, _ in variants}
for ly in variants[0][1].ops:
    if getattr(ly, "kind", "conv") != "conv":
        continue
    ts = {k: [] for k, _ in variants}
    for rep in range(7):
        for (k, e), v in zip(variants, (V0, V1)):
            l2 = next(l for l in e.ops if getattr(l, "name", None) == ly.name)
            setopt(**{OPT: v})
            ts[k].append(t1(lambda: e.launch_op(l2, sp)) * 1e3)
    row = {k: round(float(np.median(v[2:])), 1) for k, v in ts.items()}
    for k in tot: tot[k] += row[k]
    out[ly.name] = row
    vals = list(row.values())
    if abs(vals[0] - vals[1]) > 1.5:
        print(f"{ly.name:30s}", row, flush=True)
print("sum of layers (us):", {k: round(v, 1) for k, v in tot.items()})
setopt(**{OPT: 1}); setopt(tc_pdl=1)
if len(sys.argv) > 2:
    json.dump(out, open(sys.argv[2], "w"), indent=1)
