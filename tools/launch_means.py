"""per-kernel totals of an ncu launch list (--metrics gpu__time_duration.sum --csv): python tools/launch_means.py file.csv"""
import collections, csv, sys
rows = list(csv.reader(open(sys.argv[1])))
h = [i for i, r in enumerate(rows) if 'Kernel Name' in r][0]
hd = rows[h]; ik = hd.index('Kernel Name'); iv = hd.index('Metric Value')
agg = collections.OrderedDict()
for r in rows[h + 2:]:
    if len(r) != len(hd): continue
    k = r[ik].split('(')[0].split('::')[-1]; v = float(r[iv].replace(',', ''))
    a = agg.setdefault(k, [0, 0.0]); a[0] += 1; a[1] += v
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{k:32s} launches {n:4d}  total {t/1e3:10.1f} us  mean {t/n/1e3:8.1f} us")
