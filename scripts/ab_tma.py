"""ABAB timing: epilogue output through staged TMA stores (tc_tma_store=1) against per-lane global stores, layer by layer.
   python scripts/ab_tma.py [out.json]"""
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
variants = [("stg", mk(tc_tma_store=0)), ("tma", mk(tc_tma_store=1))]
sp = L.stream_ptr()
out = {"equal": bool(torch.equal(variants[0][1].output, variants[1][1].output))}
print(out, flush=True)
seen, tot = set(), {k: 0.0 for k, _ in variants}
for ly in variants[0][1].ops:
    if not hasattr(ly, "name") or getattr(ly, "kind", "conv") != "conv":
        continue
    ts = {k: [] for k, _ in variants}
    for rep in range(7):
        for k, e in variants:
            l2 = next(l for l in e.ops if getattr(l, "name", None) == ly.name)
            setopt(tc_tma_store=1 if k == "tma" else 0)
            ts[k].append(t1(lambda: e.launch_op(l2, sp)) * 1e3)
    row = {k: round(float(np.median(v[2:])), 1) for k, v in ts.items()}
    for k in tot: tot[k] += row[k]
    out[ly.name] = row
    if abs(row["stg"] - row["tma"]) > 1.5 or ly.name.endswith("layers.0.main.0") or ly.name.endswith("layers.0.main.1"):
        print(f"{ly.name:30s}", row, flush=True)
print("sum of layers (us):", {k: round(v, 1) for k, v in tot.items()})
setopt(tc_tma_store=1, tc_pdl=1)
if len(sys.argv) > 1:
    json.dump(out, open(sys.argv[1], "w"), indent=1)
