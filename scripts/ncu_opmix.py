"""Opcode mix of one kernel instance in an .ncu-rep: python scripts/ncu_opmix.py rep kernel_regex [skip]"""
import csv, io, subprocess, sys, collections, re
rep, rx = sys.argv[1], sys.argv[2]
skip = sys.argv[3] if len(sys.argv) > 3 else "0"
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "-k", f"regex:{rx}", "-s", skip, "-c", "1"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[1]; idx = {h: i for i, h in enumerate(hdr)}; data = rows[2:]
def f(r, k):
    try: return float(r[idx[k]])
    except Exception: return 0.0
mix = collections.Counter(); smp = collections.Counter()
for r in data:
    s = r[idx['Source']].strip()
    s = re.sub(r'^@!?U?P\w+\s+', '', s)
    op = s.split()[0].split('.')[0] if s else '?'
    mix[op] += f(r, 'Instructions Executed'); smp[op] += f(r, '# Samples')
tot = sum(mix.values())
print(rows[0][1][:80], 'warp-instr', tot, 'samples', sum(smp.values()))
for op, n in mix.most_common(40):
    print(f"{op:12s} {n:12.0f} {100*n/tot:5.1f}%  samples {smp[op]:7.0f}")
