"""per-kernel timing table of a few frames (profiling level 1) + counted work (level 2).
usage: python tools/quick_bench.py [scene ...]   scenes: c2 c2flat c2m c3 cornell"""
import importlib, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
snk = importlib.import_module("snake-vk_b200")
frames = int(os.environ.get("FRAMES", "6"))
for name in (sys.argv[1:] or ["c2"]):
    w, h = (3840, 2160) if name.startswith("c4") else (1920, 1080)  # C4 is the 4K config
    sc, desc = bench.make_scene(snk, name, w, h)
    t0 = time.time()
    pt = snk.PathTracer(sc.width, sc.height)
    pt.upload_scene(sc)
    print(f"== {name}: {desc}; upload+build {time.time()-t0:.2f} s, last build {float(pt.stats()['ms_build_last']):.3f} ms")
    st = snk.api.default_settings(mode=1 if os.environ.get("MODE", "reference") == "northstar" else 0,
                                  ris_iters=int(os.environ.get("RIS", "32")))
    def step(f):
        ubo = bench.frame_ubo(sc, f); prev = bench.frame_ubo(sc, f - 1) if f else ubo
        pt.build_tlas(); pt.set_frame(ubo, prev); pt.trace_primary(); pt.render(st)
        if os.environ.get("TAA"):
            pt.taa_resolve()
    for f in range(3):
        step(f)
    pt.synchronize()
    # the product configuration (no per-launch events; the north-star chains may overlap): wall time of 20 back-to-back frames
    t0 = time.time()
    for f in range(3, 23):
        step(f)
    pt.synchronize()
    print(f"   unprofiled: {(time.time() - t0) / 20 * 1e3:.3f} ms per frame (20 frames back to back)")
    s0 = pt.stats()
    pt.set_profiling(1)
    t0 = time.time()
    for f in range(3, 3 + frames):
        step(f)
    pt.synchronize()
    wall = (time.time() - t0) / frames * 1e3
    s1 = pt.stats()
    rays = sum(int(s1[k]) - int(s0[k]) for k in ("total_rays_primary", "total_rays_shadow", "total_rays_secondary")) / frames
    T = pt.launch_timings()
    per = {}
    for r in T:
        d = per.setdefault(r["name"].decode(), [0.0, 0]); d[0] += float(r["ms"]); d[1] += 1
    pt.set_profiling(2)
    step(3 + frames); pt.synchronize()
    work = {r["name"].decode(): r for r in pt.launch_timings() if int(r["is_trace"])}
    pt.set_profiling(0)
    tot = sum(v[0] / v[1] for v in per.values())
    print(f"   frame {wall:.2f} ms wall, kernels {tot:.2f} ms, {rays/1e6:.1f} M rays -> {rays/wall/1e3:.1f} Mrays/s; overflow {int(s1['stack_overflows'])}")
    for k, v in per.items():
        ms = v[0] / v[1]
        extra = ""
        if k in work:
            w = work[k]
            live = int(w["live"])
            extra = (f"  live {live/1e6:6.2f} M  {live/ms/1e6:6.3f} Grays/s  nodes/ray {int(w['nodes'])/max(live,1):5.1f} tris/ray {int(w['tris'])/max(live,1):4.1f} "
                     f"inst/ray {int(w['insts'])/max(live,1):4.1f}  alg {int(w['alg_bytes'])/ms/1e6:7.1f} GB/s")
        print(f"   {k:18s} {ms:8.3f} ms {ms/tot*100:5.1f}%{extra}")
    pt.close()
