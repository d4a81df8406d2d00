"""hot SASS blocks of an ncu --page source --csv dump: python tools/sass_hot.py file.csv [min_share]"""
import csv, sys
from collections import Counter
rows = list(csv.reader(open(sys.argv[1])))
minshare = float(sys.argv[2]) if len(sys.argv) > 2 else 0.004
which = int(sys.argv[3]) if len(sys.argv) > 3 else 0
starts = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"] + [len(rows)]
print(rows[starts[which]][1][:90])
hdr = rows[starts[which] + 1]
data = [r for r in rows[starts[which] + 2:starts[which + 1]] if len(r) == len(hdr)]
ia, it, isrc, iss = hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed"), hdr.index("Source"), hdr.index("# Samples")
tot = sum(int(r[ia]) for r in data); tth = sum(int(r[it]) for r in data); ts = sum(int(r[iss]) for r in data)
print("SASS instrs", len(data), "warp instr", tot, "thread instr", tth, "avg active", round(tth / tot, 2), "samples", ts)
blk = []
for i, r in enumerate(data):
    n, th = int(r[ia]), int(r[it])
    if blk and abs(n - blk[-1][2]) <= 0.02 * max(n, blk[-1][2], 1):
        b = blk[-1]; b[1] = i; b[3] += n; b[4] += th; b[5] += int(r[iss])
    else:
        blk.append([i, i, n, n, th, int(r[iss])])
def op(s):
    p = s.split()
    return (p[1] if p[0].startswith('@') else p[0]).split('.')[0]
for b in blk:
    if b[3] > minshare * tot:
        c = Counter(op(data[j][isrc]) for j in range(b[0], b[1] + 1)).most_common(7)
        print(f"[{b[0]:4d}-{b[1]:4d}] n={b[1]-b[0]+1:4d} exec={b[2]:>10d} share={b[3]/tot:6.3f} thr={b[4]/max(b[3],1):5.1f} stall={b[5]/ts:6.3f} {c}")
