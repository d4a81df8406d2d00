"""where the end-to-end step (bench.py `e2e`) spends its extra time over the device-only step: the frame loop with each host-facing
call added in turn. usage: python tools/e2e_breakdown.py"""
import importlib, os, sys, time
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
snk = importlib.import_module("snake-vk_b200")
W, H = 1920, 1080
sc, desc = bench.make_scene(snk, "c2", W, H)
pt = snk.PathTracer(W, H)
flat = pt.upload_scene(sc) or sc.flatten()
stream = torch.cuda.Stream()
st = snk.api.default_settings(mode=1)
out_host = [torch.empty((H, W, 4), dtype=torch.float16, pin_memory=True).numpy() for _ in range(2)]
with torch.cuda.stream(stream):
    loop = bench.FrameLoop(snk, sc, pt, stream.cuda_stream, st)
    for name in ("plain", "+set_transforms", "+set_instances", "+readback", "all"):
        k = [0]
        def step():
            if name in ("+set_transforms", "all"):
                pt.set_transforms(flat["transforms"], flat["prev_transforms"])
            if name in ("+set_instances", "all"):
                pt.set_instances(flat["instances"], flat["blas_of_instance"], flat["emissive_idx"])
            loop.step()
            if name in ("+readback", "all"):
                pt.read_output_async(out_host[k[0] % 2], slot=k[0] % 2)
                if k[0]:
                    pt.wait_output((k[0] - 1) % 2)
            k[0] += 1
        for _ in range(3):
            step()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.time()
        e0.record(stream)
        for _ in range(20):
            step()
        t_issue = time.time() - t0
        e1.record(stream)
        torch.cuda.synchronize()
        print(f"{name:18s} {e0.elapsed_time(e1) / 20:.3f} ms per step on the device, host issue {t_issue / 20 * 1e3:.3f} ms per step")

# the host's own cost of issuing one frame: the same calls on an 96x64 image of the same scene (the GPU work is negligible there)
pt2 = snk.PathTracer(96, 64)
sc2, _ = bench.make_scene(snk, "c2", 96, 64)
pt2.upload_scene(sc2)
with torch.cuda.stream(stream):
    calls = {"frame_ubo x2 (python)": lambda f: (bench.frame_ubo(sc2, f), bench.frame_ubo(sc2, f - 1)),
             "build_tlas": lambda f: pt2.build_tlas(stream.cuda_stream),
             "set_frame": lambda f: pt2.set_frame(ubo, ubo),
             "trace_primary": lambda f: pt2.trace_primary(jitter=bench.frame_jitter(snk, f, 96), stream=stream.cuda_stream),
             "render": lambda f: pt2.render(st, stream.cuda_stream),
             "set_transforms": lambda f: pt2.set_transforms(flat["transforms"], flat["prev_transforms"]),
             "set_instances": lambda f: pt2.set_instances(flat["instances"], flat["blas_of_instance"], flat["emissive_idx"])}
    ubo = bench.frame_ubo(sc2, 0)
    tot = {k: 0.0 for k in calls}
    N = 200
    for f in range(N + 5):
        for k, fn in calls.items():
            t0 = time.perf_counter()
            fn(f + 1)
            if f >= 5:
                tot[k] += time.perf_counter() - t0
        if f % 8 == 0:
            torch.cuda.synchronize()
    for k, v in tot.items():
        print(f"host cost of {k:24s} {v / N * 1e6:8.1f} us per frame")

# device -> host bandwidth of one RGBA16F image (page-locked destination)
img = torch.empty((H, W, 4), dtype=torch.float16, device="cuda")
dst = torch.empty((H, W, 4), dtype=torch.float16, pin_memory=True)
for _ in range(3):
    dst.copy_(img, non_blocking=True)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    dst.copy_(img, non_blocking=True)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
print(f"D2H of {img.numel() * 2 / 1e6:.1f} MB: {ms:.3f} ms = {img.numel() * 2 / ms / 1e6:.1f} GB/s")
# the frame loop with the read of frame k waited for one frame later (three images in flight)
out3 = [torch.empty((H, W, 4), dtype=torch.float16, pin_memory=True).numpy() for _ in range(3)]
with torch.cuda.stream(stream):
    for depth in (1, 2):
        k = [0]
        def step():
            pt.set_transforms(flat["transforms"], flat["prev_transforms"])
            pt.set_instances(flat["instances"], flat["blas_of_instance"], flat["emissive_idx"])
            loop.step()
            pt.read_output_async(out3[k[0] % 3], slot=k[0] % 3)
            if k[0] >= depth:
                pt.wait_output((k[0] - depth) % 3)
            k[0] += 1
        for _ in range(4):
            step()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(20):
            step()
        e1.record(stream)
        torch.cuda.synchronize()
        print(f"e2e with the host waiting for the image of {depth} frame(s) ago: {e0.elapsed_time(e1) / 20:.3f} ms per step")
