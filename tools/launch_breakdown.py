"""aggregate an ncu gpu__time_duration launch list between the (k-1)-th and k-th k_init (one BVH build)
usage: python tools/launch_breakdown.py file.csv [build_index]"""
import csv, collections, sys
lines = [l for l in open(sys.argv[1]) if not l.startswith('==')]
rows = list(csv.DictReader(lines))
names = [r['Kernel Name'].split('(')[0][-40:] for r in rows]
vals = [float(r['Metric Value'].replace(',', '')) * {'ns': 1e-6, 'us': 1e-3, 'ms': 1}.get(r['Metric Unit'], 1) for r in rows]
idx = [i for i, n in enumerate(names) if 'k_init' in n] + [len(names)]
k = int(sys.argv[2]) if len(sys.argv) > 2 else len(idx) - 3
s, e = idx[k], idx[k + 1]
agg = collections.OrderedDict()
for n, v in zip(names[s:e], vals[s:e]):
    a = agg.setdefault(n, [0, 0.0]); a[0] += 1; a[1] += v
tot = sum(a[1] for a in agg.values())
for n, a in agg.items():
    print(f"{n:42s} n={a[0]:3d} ms={a[1]:7.3f} {a[1]/tot:.3f}")
print("total", round(tot, 3))
