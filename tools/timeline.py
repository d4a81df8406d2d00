"""timeline of one north-star frame with the two chains overlapped (profiling level 4: every launch bracketed on its own stream,
start offsets relative to the first launch) + the unprofiled frame time. usage: [SNKPT_NS_STAGGER=a,b] python tools/timeline.py [scene]"""
import importlib, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
snk = importlib.import_module("snake-vk_b200")
name = sys.argv[1] if len(sys.argv) > 1 else "c2"
w, h = (3840, 2160) if name.startswith("c4") else (1920, 1080)
sc, desc = bench.make_scene(snk, name, w, h)
pt = snk.PathTracer(sc.width, sc.height)
pt.upload_scene(sc)
st = snk.api.default_settings(mode=1)
def step(f):
    ubo = bench.frame_ubo(sc, f); prev = bench.frame_ubo(sc, f - 1) if f else ubo
    pt.build_tlas(); pt.set_frame(ubo, prev); pt.trace_primary(jitter=bench.frame_jitter(snk, f, sc.width)); pt.render(st)
for f in range(4):
    step(f)
pt.synchronize()
best = 1e9
for rep in range(3):
    t0 = time.time()
    for f in range(4, 44):
        step(f)
    pt.synchronize()
    best = min(best, (time.time() - t0) / 40 * 1e3)
print(f"{name} stagger={os.environ.get('SNKPT_NS_STAGGER', '-')}: {best:.3f} ms per frame (best of 3 x 40 frames back to back)")
if os.environ.get("TIMELINE", "1") != "0":
    pt.set_profiling(4)
    step(44); pt.synchronize()
    T = pt.launch_timings()
    pt.set_profiling(0)
    for r in T:
        t0 = int(r["nodes"]) / 1e6
        print(f"   {r['name'].decode():16s} start {t0:7.3f}  end {t0 + float(r['ms']):7.3f}  ({float(r['ms']):6.3f} ms)")
pt.close()
