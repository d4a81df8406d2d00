"""bisects the cost of the per-frame read-back in the e2e step. usage: python tools/e2e_bisect.py"""
import importlib, os, sys, time
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
snk = importlib.import_module("snake-vk_b200")
W, H = 1920, 1080
sc, desc = bench.make_scene(snk, "c2", W, H)
pt = snk.PathTracer(W, H)
flat = pt.upload_scene(sc)
stream = torch.cuda.Stream()
copy_stream = torch.cuda.Stream()
st = snk.api.default_settings(mode=1)
out = [torch.empty((H, W, 4), dtype=torch.float16, pin_memory=True) for _ in range(3)]
outn = [t.numpy() for t in out]
img = torch.empty((H, W, 4), dtype=torch.float16, device="cuda")
with torch.cuda.stream(stream):
    loop = bench.FrameLoop(snk, sc, pt, stream.cuda_stream, st)
    flat_img, flat_out = img.view(-1), [t.view(-1) for t in out]
    NCH = 32
    CH = flat_img.numel() // NCH
    for name in ("plain", "torch_copy_other_stream", "plain", "torch_copy_chunked", "plain", "torch_copy_same_stream", "plain", "read_wait1", "plain"):
        loop.f = 0
        pt.reset_history()
        k = [0]
        def step():
            loop.step()
            i = k[0]
            if name == "read_no_wait":
                pt.read_output_async(outn[i % 3], slot=i % 3)
            elif name == "read_wait1":
                pt.read_output_async(outn[i % 3], slot=i % 3)
                if i:
                    pt.wait_output((i - 1) % 3)
            elif name == "torch_copy_other_stream":
                with torch.cuda.stream(copy_stream):
                    out[i % 3].copy_(img, non_blocking=True)
            elif name == "torch_copy_chunked":
                with torch.cuda.stream(copy_stream):
                    for c in range(NCH):
                        flat_out[i % 3][c * CH:(c + 1) * CH].copy_(flat_img[c * CH:(c + 1) * CH], non_blocking=True)
            elif name == "torch_copy_same_stream":
                out[i % 3].copy_(img, non_blocking=True)
            elif name == "torch_copy_after_event":
                ev = stream.record_event()
                copy_stream.wait_event(ev)
                with torch.cuda.stream(copy_stream):
                    out[i % 3].copy_(img, non_blocking=True)
            elif name == "sync_each_frame":
                stream.synchronize()
            k[0] += 1
        for _ in range(4):
            step()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.time()
        e0.record(stream)
        for _ in range(30):
            step()
        e1.record(stream)
        torch.cuda.synchronize()
        wall = (time.time() - t0) / 30 * 1e3
        for s in range(3):
            pt.wait_output(s)
        print(f"{name:26s} {e0.elapsed_time(e1) / 30:.3f} ms per step (events), {wall:.3f} ms wall")
