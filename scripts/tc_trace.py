"""Role timelines of CTA 0 for single conv layers of the C3 net (READ_DIAG build).
   READ_B200_LIB=read_b200/libread_b200_diag.so python scripts/tc_trace.py [layer,layer..] [mt]"""
import ctypes, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from read_b200 import synth, _lib as L
from read_b200.engine import UNetEngine
W, H = 1920, 1088
dev = torch.device("cuda", 0); torch.cuda.set_device(0)
lib = L.load()
lib.read_set_trace_buffer.argtypes = [ctypes.c_void_p]
lib.read_set_trace_buffer.restype = None
names = sys.argv[1].split(",") if len(sys.argv) > 1 else ["Encoder.0.layers.0.main.0", "Encoder.0.layers.0.main.1", "Encoder.1.layers.0.main.0"]
mt = int(sys.argv[2]) if len(sys.argv) > 2 else 0
L.check(lib.read_set_option(b"tc_mt", mt))
L.check(lib.read_set_option(b"tc_pair", int(sys.argv[3]) if len(sys.argv) > 3 else 1))
L.check(lib.read_set_option(b"tc_pdl", 0))
eng = UNetEngine(synth.synth_state_dict(synth.SEED), 1, H, W, dev, precision="bf16", use_graph=False)
for t in eng.inputs: t.uniform_(0, 1)
eng.run(); torch.cuda.synchronize()
by = {ly.name: ly for ly in eng.layers}
sp = L.stream_ptr()
NAMES = {1: "P:empty ok", 2: "P:tma issued", 3: "I:tempty ok", 4: "I:afull ok", 5: "I:mma+commit issued", 6: "E:tfull ok", 7: "E:tmem ld done",
         8: "E:item done", 9: "E:item start", 10: "E:math done", 11: "E:sts done", 12: "E:fence+sync"}
for n in names:
    buf = torch.zeros(32 * 2048, dtype=torch.int64, device=dev)
    eng.launch_op(by[n], sp); torch.cuda.synchronize()
    lib.read_set_trace_buffer(buf.data_ptr())
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); eng.launch_op(by[n], sp); b.record(); torch.cuda.synchronize()
    lib.read_set_trace_buffer(None)
    tr = buf.cpu().numpy().astype(np.uint64).reshape(32, 2048)
    print(f"==== {n}  mt={mt}  {a.elapsed_time(b)*1e3:.1f} us")
    t0 = None
    for r in range(20):
        cnt = int(tr[r, 0])
        if cnt == 0: continue
        ev = tr[r, 1:1 + cnt]
        code = (ev >> np.uint64(56)).astype(np.int64); clk = (ev & np.uint64(0x00FFFFFFFFFFFFFF)).astype(np.int64)
        if t0 is None: t0 = clk.min()
        t0 = min(t0, clk.min())
    for r in range(20):
        cnt = int(tr[r, 0])
        if cnt == 0: continue
        ev = tr[r, 1:1 + cnt]
        code = (ev >> np.uint64(56)).astype(np.int64); clk = (ev & np.uint64(0x00FFFFFFFFFFFFFF)).astype(np.int64) - t0
        span = clk[-1] - clk[0]
        line = f"role {r:2d}: {cnt:4d} events over {span:8d} cycles;"
        # mean delta INTO each code (time since the role's previous event), steady state = middle half
        lo, hi = cnt // 4, 3 * cnt // 4
        d = np.diff(clk)
        for c in sorted(set(code.tolist())):
            sel = [i for i in range(max(lo, 1), hi) if code[i] == c]
            if sel:
                line += f"  ->{NAMES.get(c, c)}: {np.mean([d[i - 1] for i in sel]):7.0f}"
        print(line)
    # one role's raw timeline sample (issuer 0 and epilogue warp 4), events 40..60
    for r in (0, 2, 4, 5, 12):
        cnt = int(tr[r, 0])
        if cnt < 60: continue
        ev = tr[r, 41:61]
        print(f"   role {r} sample:", " ".join(f"{int(e >> np.uint64(56))}@{int(e & np.uint64(0x00FFFFFFFFFFFFFF)) - int(t0)}" for e in ev))
