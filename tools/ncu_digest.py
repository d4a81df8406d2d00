"""Digest of an .ncu-rep on the GPU box (the reports are too large to bring back): per-launch key metrics, DRAM bytes, and the hot
SASS blocks with lanes per instruction.  usage: python tools/ncu_digest.py <rep> <out.txt> [max_launches_for_sass]"""
import csv, io, subprocess, sys
from collections import Counter

rep, out = sys.argv[1], sys.argv[2]
max_sass = int(sys.argv[3]) if len(sys.argv) > 3 else 40
KEYS = ["Duration", "Registers Per Thread", "Theoretical Occupancy", "Achieved Occupancy", "L1/TEX Hit Rate", "L2 Hit Rate", "DRAM Throughput",
        "Memory Throughput", "Compute (SM) Throughput", "Executed Ipc Active", "Warp Cycles Per Issued Instruction", "Avg. Active Threads Per Warp",
        "Avg. Not Predicated Off Threads Per Warp", "No Eligible", "Eligible Warps Per Scheduler", "Branch Efficiency", "Issue Slots Busy",
        "Local Memory Spilling", "Dynamic Shared Memory", "Static Shared Memory", "Block Limit"]
det = subprocess.run(["ncu", "-i", rep, "--page", "details"], capture_output=True, text=True).stdout
lines = []
for ln in det.splitlines():
    s = ln.strip()
    if "Context" in s and "Device" in s:
        lines.append("\n" + s[:160])
    elif any(s.startswith(k) for k in KEYS):
        lines.append("  " + " ".join(s.split()))
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv", "--metrics", "dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
lines.append("\n# DRAM traffic per launch (dram__bytes_read.sum + dram__bytes_write.sum)")
if len(rows) > 2:
    hdr, units = rows[0], rows[1]
    def col(n): return hdr.index(n) if n in hdr else None
    ck, cr, cw, cd = col("Kernel Name"), col("dram__bytes_read.sum"), col("dram__bytes_write.sum"), col("gpu__time_duration.sum")
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    tscale = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}
    for r in rows[2:]:
        try:
            rd = float(r[cr].replace(",", "")) * scale.get(units[cr], 1); wr = float(r[cw].replace(",", "")) * scale.get(units[cw], 1)
            us = float(r[cd].replace(",", "")) * tscale.get(units[cd], 1)
            lines.append(f"  {r[ck][:60]:60s} {us:10.1f} us  read {rd/1e6:9.1f} MB  write {wr/1e6:9.1f} MB")
        except Exception:
            pass
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
starts = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"] + [len(rows)]
lines.append("\n# hot SASS blocks per launch: [first-last] instructions, executions, share of issued warp instructions, lanes per instruction, stall-sample share, mix")
seen = Counter()
for w in range(min(len(starts) - 1, max_sass)):
    name = rows[starts[w]][1]
    short = name.split("(")[0].split("::")[-1]
    seen[short] += 1
    if seen[short] > 2:
        continue
    hdr = rows[starts[w] + 1]
    data = [r for r in rows[starts[w] + 2:starts[w + 1]] if len(r) == len(hdr)]
    if not data:
        continue
    ia, it, isrc, iss = hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed"), hdr.index("Source"), hdr.index("# Samples")
    tot = sum(int(r[ia]) for r in data); tth = sum(int(r[it]) for r in data); ts = max(sum(int(r[iss]) for r in data), 1)
    lines.append(f"\n{short} (launch {w}): SASS instructions {len(data)}, warp instructions {tot}, thread instructions {tth}, lanes per instruction {tth / max(tot, 1):.2f}")
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
        if b[3] > 0.01 * tot:
            c = Counter(op(data[j][isrc]) for j in range(b[0], b[1] + 1)).most_common(5)
            lines.append(f"  [{b[0]:4d}-{b[1]:4d}] n={b[1]-b[0]+1:4d} exec={b[2]:>9d} share={b[3]/tot:6.3f} lanes={b[4]/max(b[3],1):5.1f} stall={b[5]/ts:6.3f} {c}")
open(out, "w").write("\n".join(lines) + "\n")
print(f"{out}: {len(lines)} lines")
