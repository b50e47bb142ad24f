"""Top stall locations of one kernel instance in an .ncu-rep:  python scripts/ncu_stalls.py rep kernel_regex [skip]"""
import csv, io, subprocess, sys
rep, rx = sys.argv[1], sys.argv[2]
skip = sys.argv[3] if len(sys.argv) > 3 else "0"
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "-k", f"regex:{rx}", "-s", skip, "-c", "1"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[1]; idx = {h: i for i, h in enumerate(hdr)}; data = rows[2:]
def f(r, k):
    try: return float(r[idx[k]])
    except Exception: return 0.0
tot = sum(f(r, '# Samples') for r in data)
print('kernel', rows[0][1][:60], 'samples', tot, 'warp-instr', sum(f(r, 'Instructions Executed') for r in data))
stalls = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
agg = {k: sum(f(r, k) for r in data) for k in stalls}
print(sorted(((k, v) for k, v in agg.items() if v > 0), key=lambda kv: -kv[1])[:7])
for r in sorted(data, key=lambda r: -f(r, '# Samples'))[:int(sys.argv[4]) if len(sys.argv) > 4 else 18]:
    s = {k: f(r, k) for k in stalls if f(r, k) > 0}
    main = sorted(s.items(), key=lambda kv: -kv[1])[:2]
    print(f"{f(r,'# Samples'):8.0f} {r[idx['Source']][:66]:66s} exec={f(r,'Instructions Executed'):9.0f} {main}")
