import json, re, collections, sys
f = sys.argv[1]
rows = json.load(open(f))
grp = collections.OrderedDict()
for r in rows:
    k = (re.sub(r'(Encoder|Decoder)\.(\d)\.layers\.\d+\.main\.\d', r'\1.\2.*', r['name']), r['impl'])
    g = grp.setdefault(k, [0, 0.0, 0.0]); g[0] += 1; g[1] += r['ms']; g[2] += r['gflop']
thr = float(sys.argv[2]) if len(sys.argv) > 2 else 0.09
for (k, impl), (n, ms, gf) in grp.items():
    if n > 1 or ms > thr: print(f"{k:26s} impl={impl:2d} n={n:2d} ms={ms:7.3f} ({ms/n:6.3f}/layer) {gf/ms if ms else 0:8.1f} TF/s")
print("total ms", round(sum(r['ms'] for r in rows), 3))
