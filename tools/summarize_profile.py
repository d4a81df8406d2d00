"""Summarise a gpurun_out/<tag> directory into profiles/<tag>_*.{txt,json} (committed evidence).
usage: python tools/summarize_profile.py <tag>"""
import collections, csv, json, os, subprocess, sys

tag = sys.argv[1]
src = os.path.join("gpurun_out", tag)
os.makedirs("profiles", exist_ok=True)

def launches():
    p = os.path.join(src, "launches.csv")
    if not os.path.exists(p):
        return
    lines = [l for l in open(p) if not l.startswith("==")]
    agg = collections.OrderedDict()
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(row["Metric Value"].replace(",", ""))
        v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(row["Metric Unit"], 1.0)
        a = agg.setdefault(row["Kernel Name"].split("(")[0][-48:], [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    with open(f"profiles/{tag}_launches.txt", "w") as f:
        f.write(f"# ncu --metrics gpu__time_duration.sum --clock-control none; python bench.py --steps 2 --warmup 1 ({tag})\n")
        f.write("# cold-cache, serialised launch times: compare SHARES with bench.py's event shares, not absolutes\n")
        for k, a in sorted(agg.items(), key=lambda x: -x[1][1]):
            f.write(f"{k:50s} n={a[0]:4d} total_ms={a[1]:10.3f} avg_ms={a[1]/a[0]:9.4f} share={a[1]/tot:.4f}\n")
    print(open(f"profiles/{tag}_launches.txt").read())

KEYS = ["Duration", "Registers Per Thread", "Theoretical Occupancy", "Achieved Occupancy", "L1/TEX Hit Rate", "L2 Hit Rate", "DRAM Throughput",
        "Memory Throughput", "Compute (SM) Throughput", "Executed Ipc Active", "Warp Cycles Per Issued Instruction", "Avg. Active Threads Per Warp",
        "Avg. Not Predicated Off Threads Per Warp", "No Eligible", "Eligible Warps Per Scheduler", "Branch Efficiency", "Block Limit",
        "Shared Memory Configuration Size", "Dynamic Shared Memory", "Static Shared Memory", "Local Memory Spilling", "L2 Cache Throughput",
        "L1/TEX Cache Throughput", "Max Bandwidth", "Mem Busy", "Issue Slots Busy"]

def ncu(name):
    rep = os.path.join(src, f"{name}.ncu-rep")
    if not os.path.exists(rep):
        return
    det = subprocess.run(["ncu", "-i", rep, "--page", "details"], capture_output=True, text=True).stdout
    out = []
    for ln in det.splitlines():
        s = ln.strip()
        if "Context" in s and "Device" in s:
            out.append("\n" + s)
        elif any(s.startswith(k) for k in KEYS):
            out.append("  " + " ".join(s.split()))
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    traffic = {}
    if rows:
        hdr = rows[0]
        def col(n):
            return hdr.index(n) if n in hdr else None
        ck, cr, cw, cd = col("Kernel Name"), col("dram__bytes_read.sum"), col("dram__bytes_write.sum"), col("gpu__time_duration.sum")
        units = rows[1]
        def scale(u):
            return {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
        for r in rows[2:]:
            if cr is None:
                break
            b = float(r[cr].replace(",", "")) * scale(units[cr]) + float(r[cw].replace(",", "")) * scale(units[cw])
            traffic.setdefault(r[ck].split("(")[0][-40:], []).append(b)
    with open(f"profiles/{tag}_{name}.txt", "w") as f:
        f.write(f"# ncu --set full --clock-control none --import-source on ({tag}/{name}.ncu-rep), selected lines of --page details\n")
        f.write("\n".join(out) + "\n")
        f.write("\n# dram__bytes_read.sum + dram__bytes_write.sum per launch\n")
        for k, v in traffic.items():
            f.write(f"{k}: {[int(x) for x in v]}\n")
    print(open(f"profiles/{tag}_{name}.txt").read())

launches()
for n in ("prof_trace", "prof_traceq", "prof_restir", "prof_shade"):
    ncu(n)
for n in ("bench.json", "bench_ref.json", "gpu.txt", "pytest_gpu.log", "smoke.log"):
    p = os.path.join(src, n)
    if os.path.exists(p):
        open(f"profiles/{tag}_{n}", "w").write(open(p).read())
