#!/usr/bin/env python
"""bench.py — the path-tracing frame (BASELINE.json metric) on N B200s of one node.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--mode northstar|reference] [--scene c2]

A step = one frame of the pass on the C2 workload (261,122 triangles in 205 instances, 1920x1080): per-frame TLAS rebuild
(TlasSystem::BuildTlas), primary visibility with the engine's Halton sub-pixel jitter (replaces the raster G-buffer pass),
the ReSTIR DI stages and the path stages. Two estimators (never mixed in one number): `--mode northstar`
(default; BASELINE.json's metric: 1 spp, 4 bounces + ReSTIR DI; its oracle is this repository's own CPU statement — the
reference has no such estimator) and `--mode reference` (the reference's own frame, RT::RecordRenderCmdBuf: ReSTIR + the
8x(1+2) BounceDiffuse tree, <= 65 rays/pixel — the mode whose results are pinned to the reference's shader text); the mode
not selected is measured too and reported under `other_mode`.

  value       all rays traced by all ranks / max-over-ranks device time (CUDA events on the frame's stream), inputs resident
  e2e         the same through the C ABI with HOST buffers: per-frame transform / instance upload + read-back of the image
  roofline    dominant traversal launch: algorithmic bytes (counted replay) / its event duration in the timed region
  other_configs (N = 1)  C2 with 1 M triangles, C3 (10,248 emissive triangles), C4 (3 M triangles, 3840x2160, instanced and flat), C5 build /
              refit / TLAS sweep at 1e6 and 1e7 triangles
  strong (N > 1)  the frame strong-scaled as row bands (snkpt_group_attach_nccl: reservoir rows travel by ncclSend / ncclRecv
              inside snkpt_render): ms per frame and efficiency against one GPU, on C2 1080p and on C4 4K (BASELINE config 4)

N > 1: `value` is sample-sharded (every rank renders the whole frame with its own seed; radiance summed over NVLink with one
NCCL all-reduce per frame on a second stream, the timed region waits for all of them) -> weak scaling.

`--impl reference` times the CPU oracle (the Vulkan reference cannot run here: SURVEY §8c) on a bounded crop of the same frames
with all host threads; it is the only leg besides cpu_baseline that executes oracle/.
"""
from __future__ import annotations

import argparse
import importlib
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
METRIC, UNIT = "rays_per_second_frame_1080p", "Mrays/s"
ESTIMATOR = {"northstar": "1 spp, 4 bounces (GGX/Lambert sampling, MIS NEE, RR from bounce 2) + ReSTIR DI (32+32 candidates, temporal 1, spatial 3)",
             "reference": "the reference's own frame: ReSTIR DI pass 1 + 2 and the 8x(1+2) BounceDiffuse tree (<= 65 rays/pixel)"}
PARITY = {"northstar": "bit-exact against this repository's own CPU statement (oracle/snk_northstar.h, snk_restir_ns.h); the reference has no such estimator, "
                       "so no reference-pinned parity exists for this mode (other_mode carries the reference-pinned frame)",
          "reference": "bit-exact against oracle/snk_raygen.h, which is pinned to the reference's executed shader text (tests/test_cpu_reference_frames.py)"}


def make_scene(snk, name: str, width: int, height: int):
    S = snk.scenes
    if name == "c2":
        return S.c2("instanced", width=width, height=height), "C2 instanced: floor + 204 icospheres = 261,122 tris, 205 instances, 16 point lights + dir light"
    if name == "c2flat":
        return S.c2("flat", width=width, height=height), "C2 flat: 261,122 tris in one BLAS, 16 point lights + dir light"
    if name == "c2m":
        return S.c2("flat", subdiv=4, width=width, height=height), "C2 flat subdiv 4: 1,044,482 tris in one BLAS"
    if name == "c3":
        return S.c3(width=width, height=height), "C3: C2 instanced + 854 emissive cubes (10,248 emissive tris)"
    if name in ("c4", "c4flat"):
        v = "flat" if name == "c4flat" else "instanced"
        return S.c4(v, width=width, height=height), f"C4 {v}: floor + 2343 icospheres = 2,999,042 tris" + (" in one BLAS" if v == "flat" else ", 2344 instances")
    if name == "cornell":
        return S.cornell(width, height, True), "C1 Cornell box, 32 tris"
    if name == "shipped":
        # the engine's own meshes (Core/res/meshes/sphere.glb, cube.glb) from the committed arrays (the GPU box has no /root/reference)
        d = np.load(os.path.join(ROOT, "tests", "golden", "shipped_mesh_arrays.npz"))
        mesh = lambda n: S.Mesh(*(np.ascontiguousarray(d[f"{n}_{k}"]) for k in ("positions", "normals", "uvs", "tangents", "indices", "submeshes")))
        sc = S.shipped_field(mesh("sphere"), mesh("cube"), width=width, height=height)
        tris = sum((sc.meshes[e.mesh].indices.size // 3) for e in sc.entities)
        return sc, f"shipped meshes: 600 instances of Core/res/meshes/sphere.glb + cube.glb over a cube floor = {tris:,} tris, {len(sc.entities)} instances, 18 emissive cubes, 16 point lights"
    raise SystemExit(f"unknown scene {name}")


def make_config(args, desc: str, W: int, H: int) -> dict:
    """the `config` object — identical in the GPU arm and the reference arm"""
    return {"workload": desc, "scene": args.scene, "mode": args.mode, "resolution": [W, H], "estimator": ESTIMATOR[args.mode],
            "frames": "camera dolly 0.1 u/frame along +z, time 1000 f / 60 ms, Halton(2,3) sub-pixel jitter per frame (RenderCommon.h:39-59), TLAS rebuilt every frame",
            "parity": PARITY[args.mode],
            "l2": "per-pixel planes, reservoirs and ray queues (0.4-1.8 GB per frame) exceed the 126 MB L2 between iterations; the BVH itself is cache resident by size"}


def frame_ubo(sc, f: int, seed_offset: int = 0):
    """camera dolly 0.1 u/frame along +z; time_elapsed = 1000 f / 60 ms"""
    pos = np.asarray(sc.camera.pos, np.float32) + np.float32(0.1 * f) * np.array([0, 0, 1], np.float32)
    u = sc.ubo(f, cam_pos=pos)
    if seed_offset:
        u["time_elapsed"] = np.float32(float(u["time_elapsed"]) + 7919.0 * seed_offset)
    return u


def frame_jitter(snk, f: int, width: int):
    """GetFrameJitter (RenderCommon.h:50-59) as the sub-pixel offset of the camera rays, in pixels"""
    return snk.api.frame_jitter(f, width)[1]


class ClockSampler:
    """nvidia-smi clocks + throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._pump, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return None
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, power = [], [], set(), []
        for ln in self.lines:
            p = [x.strip() for x in ln.split(",")]
            if len(p) < 9:
                continue
            try:
                sm.append(float(p[1])), mx.append(float(p[2])), power.append(float(p[3]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), p[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return None
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "power_w": statistics.median(power),
                "reasons": sorted(reasons), "samples": len(sm)}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(p) as f:
            d = json.load(f)
        v = d.get("hbm_gbs") or d.get("hbm_gb_s")
        if v:
            return float(v), "measured (MEASURED_PEAKS.json)"
    except Exception:
        pass
    return 6650.0, "fallback (B200_PROFILING.md)"


# ------------------------------------------------------------------------------------------------ CPU oracle legs
def oracle_sample(snk, sc, mode: str, rows: tuple[int, int], frames: int, threads: int):
    """time the CPU oracle on rows [rows) of `frames` consecutive frames; returns (rays, seconds)"""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_api as O
    osc = O.OracleScene(sc)
    fr = O.OracleFrame(osc)
    prev = None
    rays = 0
    t_total = 0.0
    for f in range(frames):
        ubo = frame_ubo(sc, f)
        fr.set(ubo, ubo if prev is None else prev, O.default_settings(mode=0 if mode == "reference" else 1))
        j = frame_jitter(snk, f, sc.width)
        t0 = time.perf_counter()
        fr.gbuffer(jitter=(float(j[0]), float(j[1])), rows=rows, threads=threads)
        fr.render(rows=rows, threads=threads)
        t_total += time.perf_counter() - t0
        c, a = fr.ray_counts(reset=True)
        rays += c + a + (rows[1] - rows[0]) * sc.width
        prev = ubo
    return rays, t_total


def run_reference_arm(args, snk):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    sc, desc = make_scene(snk, args.scene, args.width, args.height)
    threads = os.cpu_count() or 1
    # at least 16 rows per host thread and never fewer than 256, so that every thread has work (16 rows on 16 threads ran the
    # same code at a third of its throughput)
    rows_per_step = min(args.height, max(256, args.ref_rows, 16 * threads))
    y0 = args.height // 2 - rows_per_step // 2
    rows = (y0, y0 + rows_per_step)
    for _ in range(min(max(args.warmup, 1), 2)):
        oracle_sample(snk, sc, args.mode, rows, 1, threads)
    rays, secs = 0, 0.0
    for _ in range(args.steps):
        r, s = oracle_sample(snk, sc, args.mode, rows, 2, threads)
        rays += r
        secs += s
    v = rays / secs / 1e6
    sample = (f"rows [{rows[0]},{rows[1]}) of the {args.width}x{args.height} frame on {threads} host threads, 2 consecutive frames per step "
              f"(frame 1 exercises temporal reuse); CPU restatement of the reference's shaders (oracle/), not the Vulkan reference")
    emit({
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * secs / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic", "config": make_config(args, desc, args.width, args.height),
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "CPU oracle (restatement of the reference GLSL; the Vulkan reference cannot run headless here, SURVEY §8c)",
    })


# ------------------------------------------------------------------------------------------------ GPU arm
def kname(rec) -> str:
    """kernel name of a launch-timing record (fixed 32-byte field, NUL terminated)"""
    return bytes(rec["name"]).split(b"\0", 1)[0].decode()


class CudaArray:
    """__cuda_array_interface__ shim so torch can alias a library-owned device buffer (NCCL plumbing)"""

    def __init__(self, ptr: int, shape, typestr="<f4"):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (ptr, False), "version": 3, "strides": None}


RAYS = ("total_rays_primary", "total_rays_shadow", "total_rays_secondary")


class FrameLoop:
    """one context + scene + stream: step() = one frame as the engine issues it"""

    def __init__(self, snk, sc, pt, stream, settings, seed_off=0):
        self.snk, self.sc, self.pt, self.stream, self.settings, self.seed_off = snk, sc, pt, stream, settings, seed_off
        self.f = 0
        self.after_frame = None  # hook: called with the frame number after the frame is enqueued (sample accumulation)

    def step(self):
        f = self.f
        self.f += 1
        ubo = frame_ubo(self.sc, f, self.seed_off)
        prev = frame_ubo(self.sc, f - 1, self.seed_off) if f > 0 else ubo
        pt = self.pt
        pt.build_tlas(self.stream)           # TlasSystem::BuildTlas — rebuilt every frame like the reference
        pt.set_frame(ubo, prev)
        pt.trace_primary(jitter=frame_jitter(self.snk, f, self.sc.width), stream=self.stream)
        pt.render(self.settings, self.stream)
        if self.after_frame:
            self.after_frame(f)

    def ray_total(self):
        st = self.pt.stats()
        return sum(int(st[k]) for k in RAYS), int(st["kernel_launches"])


def timed(torch, loop, steps, barrier, drain=None):
    """K frames bracketed by CUDA events on the frame's stream (barrier + synchronize on both sides) -> ms, rays, launches"""
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    r0, l0 = loop.ray_total()
    barrier()
    e0.record()
    for _ in range(steps):
        loop.step()
    if drain:
        drain()
    e1.record()
    barrier()
    r1, l1 = loop.ray_total()
    return e0.elapsed_time(e1), r1 - r0, l1 - l0


def per_kernel(timings):
    per = {}
    for r in timings:
        d = per.setdefault(kname(r), {"ms": 0.0, "n": 0, "is_trace": int(r["is_trace"])})
        d["ms"] += float(r["ms"])
        d["n"] += 1
    return per


def counted_work(loop, frames=2):
    """algorithmic work per traversal launch from counted replays (profiling level 2)"""
    loop.pt.set_profiling(2)
    for _ in range(frames):
        loop.step()
    loop.pt.synchronize()
    counted = {}
    for r in loop.pt.launch_timings():
        if int(r["is_trace"]):
            d = counted.setdefault(kname(r), {"bytes": 0, "live": 0, "nodes": 0, "tris": 0, "insts": 0, "n": 0})
            d["bytes"] += int(r["alg_bytes"]); d["live"] += int(r["live"]); d["nodes"] += int(r["nodes"])
            d["tris"] += int(r["tris"]); d["insts"] += int(r["insts"]); d["n"] += 1
    loop.pt.set_profiling(0)
    return counted


def serialized_kernels(loop, frames=4):
    """every kernel timed alone (profiling level 1: one launch at a time, the two chains of the north-star frame not overlapped)"""
    loop.pt.set_profiling(1)
    for _ in range(frames):
        loop.step()
    loop.pt.synchronize()
    per = per_kernel(loop.pt.launch_timings())
    loop.pt.set_profiling(0)
    return per


def traversal_table(per, counted):
    out = {}
    for k, v in per.items():
        if v["is_trace"] and k in counted and counted[k]["live"]:
            c = counted[k]
            ms = v["ms"] / v["n"]
            live = c["live"] / c["n"]
            out[k] = {"ms": round(ms, 4), "mrays": round(live / 1e6, 3), "grays_per_s": round(live / (ms * 1e-3) / 1e9, 3),
                      "nodes_per_ray": round(c["nodes"] / c["live"], 2), "tris_per_ray": round(c["tris"] / c["live"], 2),
                      "instances_per_ray": round(c["insts"] / c["live"], 2), "alg_gbs": round(c["bytes"] / c["n"] / (ms * 1e-3) / 1e9, 1)}
    return out


def run_gpu_arm(args, snk):
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — this arm has no CPU path (use --impl reference for the CPU oracle)")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    sc, desc = make_scene(snk, args.scene, args.width, args.height)
    W, H = sc.width, sc.height
    # one explicit stream for the frames, the timing events and the NCCL calls: the library enqueues on the stream it is given
    # (its internal chain streams fork from and join to it), so events recorded on this stream bracket exactly the work timed
    work = torch.cuda.Stream(device=local)
    torch.cuda.set_stream(work)
    stream = work.cuda_stream
    peak, peak_src = measured_peaks()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def reduce_max_sum(vals):
        """[time-like, additive...] -> max over ranks of the first, sum of the rest"""
        if world == 1:
            return vals
        t = torch.tensor(vals, device=f"cuda:{local}", dtype=torch.float64)
        tmax = t.clone()
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return [float(tmax[0])] + [float(v) for v in t[1:]]

    pt = snk.PathTracer(W, H, device=local)
    flat = pt.upload_scene(sc)
    seed_off = rank if world > 1 else 0

    # Sample sharding (N > 1): frame f's radiance is copied to one of two accumulation images and summed over the ranks on a
    # second stream while frame f+1 renders (the render never waits for NVLink); drain_samples() makes the timed stream wait for
    # every all-reduce still in flight, so each timed region contains K frames AND their K all-reduces.
    rad = None
    if world > 1:
        ptr, _ = pt.device_ptr(0x102)
        rad = torch.as_tensor(CudaArray(ptr, (H, W, 4)), device=f"cuda:{local}")
    acc = [torch.empty_like(rad) for _ in range(2)] if rad is not None else None
    comm = torch.cuda.Stream(device=local) if rad is not None else None
    acc_done = [None, None]

    def accumulate_samples(f):
        k = f & 1
        if acc_done[k] is not None:
            work.wait_event(acc_done[k])    # the all-reduce that used this image two frames ago
        acc[k].copy_(rad)
        comm.wait_event(work.record_event())
        with torch.cuda.stream(comm):
            dist.all_reduce(acc[k])
            acc_done[k] = comm.record_event()

    def drain_samples():
        for ev in acc_done:
            if ev is not None:
                work.wait_event(ev)

    def measure(mode_name: str, with_e2e: bool):
        mode = snk.api.MODE_REFERENCE if mode_name == "reference" else snk.api.MODE_NORTHSTAR
        loop = FrameLoop(snk, sc, pt, stream, snk.api.default_settings(mode=mode), seed_off)
        if rad is not None:
            loop.after_frame = accumulate_samples
        pt.set_profiling(0)
        for _ in range(args.warmup):
            loop.step()
        if rad is not None and args.warmup:
            # the pipelined accumulation image of the last warm-up frame = a plain all-reduce of that frame's radiance
            drain_samples()
            chk = rad.clone()
            dist.all_reduce(chk)
            torch.cuda.synchronize()
            if not torch.allclose(acc[(loop.f - 1) & 1], chk, rtol=1e-6, atol=0.0, equal_nan=True):
                raise SystemExit("bench.py: overlapped sample accumulation differs from a direct all-reduce")
        # ---- the timed region: product configuration (chains overlapped), traversal launches bracketed on their own streams
        pt.set_profiling(3)
        clocks = ClockSampler(local)
        if rank == 0:
            clocks.start()
        ms_total, rays, launches = timed(torch, loop, args.steps, barrier, drain_samples if rad is not None else None)
        clk = clocks.stop() if rank == 0 else None
        live = per_kernel(pt.launch_timings())
        pt.set_profiling(0)
        ms_total, rays, launches = reduce_max_sum([ms_total, float(rays), float(launches)])
        res = {"value": rays / (ms_total * 1e-3) / 1e6, "ms_per_step": ms_total / args.steps, "rays_per_frame": rays / args.steps / world,
               "gpu_launches": int(launches), "clocks": clk}

        # ---- end to end through the C ABI with host buffers (rank-local; max over ranks), over the same number of steps
        if with_e2e:
            # two page-locked images: frame k is copied out (own copy stream) while frame k+1 renders, the way the engine's
            # frames in flight present one frame while recording the next; every step still lands one finished image on the host
            out_host = [torch.empty((H, W, 4), dtype=torch.float16, pin_memory=True) for _ in range(2)]
            out_nps = [t.numpy() for t in out_host]
            h2d = 2 * 240 + 2 * flat["transforms"].nbytes + flat["instances"].nbytes + flat["blas_of_instance"].nbytes + flat["emissive_idx"].nbytes
            reads = rank == 0 or rad is None
            kk = [0]
            plain_step = loop.step

            def e2e_step():
                k = kk[0]
                kk[0] += 1
                pt.set_transforms(flat["transforms"], flat["prev_transforms"])  # TransformBufferSystem: cur + prev, host → device
                pt.set_instances(flat["instances"], flat["blas_of_instance"], flat["emissive_idx"])
                plain_step()
                if reads:
                    pt.read_output_async(out_nps[k % 2], slot=k % 2)  # device → host: the RGBA16F image the engine would present
                    if k:
                        pt.wait_output((k - 1) % 2)                   # the previous frame is now on the host

            def e2e_drain():
                if rad is not None:
                    drain_samples()
                if reads and kk[0]:
                    pt.wait_output((kk[0] - 1) % 2)

            loop.step = e2e_step
            for _ in range(2):
                loop.step()
            e2e_drain()
            e2e_ms, e2e_rays, _ = timed(torch, loop, args.steps, barrier, e2e_drain)
            loop.step = plain_step
            e2e_ms, e2e_rays = reduce_max_sum([e2e_ms, float(e2e_rays)])
            res["e2e"] = {"value": e2e_rays / (e2e_ms * 1e-3) / 1e6, "unit": UNIT, "ms_per_step": e2e_ms / args.steps, "steps": args.steps,
                          "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(out_nps[0].nbytes)}

        # ---- roofline: a second timed loop of the same K frames in which every kernel runs ALONE (profiling level 1: one launch at
        # a time, the two chains not overlapped), so that a launch's duration is its own; algorithmic bytes from counted replays
        loop.after_frame = None
        solo = serialized_kernels(loop, frames=args.steps)
        counted = counted_work(loop)
        tot_ms = max(sum(x["ms"] / x["n"] for x in solo.values()), 1e-9)
        res["kernels"] = {k: {"ms_per_launch": round(v["ms"] / v["n"], 4), "share": round(v["ms"] / v["n"] / tot_ms, 4)} for k, v in solo.items()}
        res["kernels_note"] = ("each kernel timed ALONE over a second loop of `steps` frames (profiling level 1); in the timed region of `value` the ReSTIR chain and "
                               "the path chain of the north-star frame overlap on two streams, so the frame is shorter than the sum of these")
        res["traversal"] = traversal_table(solo, counted)
        res["roofline"] = None
        names = [k for k, v in solo.items() if v["is_trace"] and k in counted and counted[k]["live"]]
        if names:
            top = max(names, key=lambda k: solo[k]["ms"])
            ms_launch = solo[top]["ms"] / solo[top]["n"]
            c = counted[top]
            bytes_launch = c["bytes"] / c["n"]
            achieved = bytes_launch / (ms_launch * 1e-3) / 1e9
            all_bytes = sum(counted[k]["bytes"] / counted[k]["n"] for k in names)
            all_live = sum(counted[k]["live"] / counted[k]["n"] for k in names)
            all_ms = sum(solo[k]["ms"] / solo[k]["n"] for k in names)
            roof = {"bound": "hbm", "kernel": f"k_trace_q ({top})" if top.startswith("trace.q") else f"k_trace_pt ({top})",
                    "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": None, "traffic_source": None,
                    "peak_source": peak_src, "alg_bytes_per_launch": bytes_launch, "ms_per_launch": ms_launch, "launches_timed": solo[top]["n"],
                    "timing": "CUDA events around this launch on the stream it runs on, mean over a timed loop of `steps` frames in which launches do not overlap",
                    "ms_per_launch_overlapped": (live[top]["ms"] / live[top]["n"]) if top in live else None,
                    "rays_per_launch": c["live"] / c["n"], "nodes_per_ray": c["nodes"] / c["live"], "tris_per_ray": c["tris"] / c["live"],
                    "instances_per_ray": c["insts"] / c["live"], "grays_per_s": c["live"] / c["n"] / (ms_launch * 1e-3) / 1e9,
                    "all_traversal_launches": {"ms_per_frame": all_ms, "achieved": all_bytes / (all_ms * 1e-3) / 1e9, "frac": all_bytes / (all_ms * 1e-3) / 1e9 / peak,
                                               "grays_per_s": all_live / (all_ms * 1e-3) / 1e9},
                    "note": "algorithmic bytes = 32+16|4 per ray (+4 reference, +4 instance plane in queue launches) + 80/node + 48/triangle record + 64/instance (SURVEY §8d); the BVH is "
                            "L1/L2 resident, the kernel is instruction-issue bound (profiles/), so the HBM fraction is context, not the limiter"}
            tr = os.path.join(ROOT, "profiles", "traffic.json")
            if os.path.exists(tr):
                try:
                    t = json.load(open(tr))
                    roof["traffic"] = t.get(top)
                    roof["traffic_source"] = "static: " + t.get("_source", "dram__bytes_read.sum + dram__bytes_write.sum of one committed ncu --set full capture (profiles/)") + "; not re-measured in this run"
                except Exception:
                    pass
            res["roofline"] = roof
        return res

    main_res = measure(args.mode, with_e2e=True)
    other = "reference" if args.mode == "northstar" else "northstar"
    other_res = None
    if not args.single_mode:
        pt.reset_history()
        other_res = measure(other, with_e2e=False)

    # ---- strong scaling (N > 1): the frame as row bands of one image, reservoir rows exchanged inside the library over NCCL
    strong = None
    if world > 1 and not args.no_strong:
        pt.reset_history()
        strong = {"C2_1080p": strong_scaling(args, snk, torch, dist, sc, pt, stream, rank, world, local, barrier, reduce_max_sum, verify=True)}
        pt.close()
        pt = None
        sc4, desc4 = make_scene(snk, "c4", 3840, 2160)
        strong["C4_4K"] = strong_scaling(args, snk, torch, dist, sc4, None, stream, rank, world, local, barrier, reduce_max_sum, verify=False)
        strong["C4_4K"]["workload"] = desc4
        strong["note"] = ("row bands of ONE image (snkpt_group_attach_nccl), band heights balanced by the rays each band traced in two probe rounds (snkpt_group_set_bands; --equal-bands keeps H/N rows): temporal reservoirs of 30 rows and final reservoirs of 16 rows "
                          "per band edge travel by ncclSend/ncclRecv on the frame's stream; only the G-buffer keeps a 31-row halo. efficiency = whole-frame ms / (N x band ms)")

    # ---- the other BASELINE configs on one GPU
    others = None
    if world == 1 and not args.no_other_configs:
        others = other_configs(args, snk, torch, stream)

    # ---- CPU baseline (rank 0, N = 1): the oracle on a bounded crop of the same frames
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        nrows = min(H, max(256, 16 * threads))
        y0 = H // 2 - nrows // 2
        r, s = oracle_sample(snk, sc, args.mode, (y0, y0 + nrows), 2, threads)
        cpu = {"value": r / s / 1e6, "unit": UNIT, "cores": threads, "kind": "port",
               "sample": f"rows [{y0},{y0 + nrows}) of the {W}x{H} frame, frames 0-1, {r} rays in {s:.1f} s (CPU restatement, oracle = cpu_restatement)"}

    if rank == 0:
        line = {
            "metric": METRIC, "value": main_res["value"], "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": main_res["ms_per_step"], "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": make_config(args, desc, W, H),
            "rays_per_frame": main_res["rays_per_frame"], "shard": "samples" if world > 1 else "none",
            "e2e": main_res.get("e2e"), "gpu_launches": main_res["gpu_launches"], "clocks": main_res["clocks"], "roofline": main_res["roofline"],
            "cpu_baseline": cpu, "kernels": main_res["kernels"], "kernels_note": main_res["kernels_note"], "traversal": main_res["traversal"],
        }
        if other_res is not None:
            line["other_mode"] = {"mode": other, "parity": PARITY[other], "value": other_res["value"], "unit": UNIT, "ms_per_step": other_res["ms_per_step"],
                                  "rays_per_frame": other_res["rays_per_frame"], "roofline": other_res["roofline"], "kernels": other_res["kernels"],
                                  "traversal": other_res["traversal"], "clocks": other_res["clocks"]}
        if strong is not None:
            line["strong"] = strong
        if others is not None:
            line["other_configs"] = others
        emit(line)
    if pt is not None:
        pt.close()
    if world > 1:
        dist.destroy_process_group()


def strong_scaling(args, snk, torch, dist, sc, whole, stream, rank, world, local, barrier, reduce_max_sum, verify):
    """row bands of one image over `world` GPUs: ms per frame, efficiency against the whole frame on one GPU, exchanged bytes"""
    W, H = sc.width, sc.height
    settings = snk.api.default_settings(mode=snk.api.MODE_NORTHSTAR)
    own_whole = whole is None
    if own_whole:
        whole = snk.PathTracer(W, H, device=local)
        whole.upload_scene(sc)
    # the whole frame on one GPU (every rank measures it on its own GPU at the same time; max over ranks)
    lw = FrameLoop(snk, sc, whole, stream, settings)
    for _ in range(args.warmup):
        lw.step()
    steps = max(5, min(args.steps, 20))
    ms_w, _, _ = timed(torch, lw, steps, barrier)
    ms_w = reduce_max_sum([ms_w])[0] / steps
    def make_group(starts):
        band = (int(starts[rank]), int(starts[rank + 1]))
        pb = snk.PathTracer(W, H, device=local, rows=band)
        pb.upload_scene(sc)
        pb.group_set_bands(starts)
        # the 128-byte NCCL id travels through torch.distributed (plumbing); the exchange itself lives in the library
        ids = [snk.api.group_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(ids, src=0)
        pb.group_attach_nccl(rank, world, ids[0], 30, 16)
        return pb, band, FrameLoop(snk, sc, pb, stream, settings)

    def gather(x):
        mine = torch.tensor([float(x)], device=f"cuda:{local}", dtype=torch.float64)
        allv = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(allv, mine)
        return [float(t[0]) for t in allv]

    # band heights: equal bands first; the rays each band traced over a few frames give a piecewise-constant cost per row
    # (plus a fixed share per row for the stages that touch every pixel), and the boundaries are moved to equal cost —
    # rows of sky cost a fraction of rows full of geometry (snkpt_group_set_bands). Two rounds settle it.
    starts = [snk.api.group_band(H, world, r)[0] for r in range(world)] + [H]
    balance_log = []
    for _ in range(0 if args.equal_bands else 2):
        pb, band, lb = make_group(starts)
        for _ in range(3):
            lb.step()
        _, rays_probe, _ = timed(torch, lb, 3, barrier)
        rays = gather(rays_probe / 3)
        pb.group_detach(); pb.close()
        rows = np.diff(np.asarray(starts, np.float64))
        mean_density = sum(rays) / H
        cost_row = np.concatenate([np.full(int(rows[r]), rays[r] / rows[r] + 0.5 * mean_density) for r in range(world)])  # 0.5: per-pixel stages
        cum = np.concatenate([[0.0], np.cumsum(cost_row)])
        new = [0] + [int(np.searchsorted(cum, cum[-1] * r / world)) for r in range(1, world)] + [H]
        for r in range(1, world + 1):  # at least 8 rows per band
            new[r] = max(new[r], new[r - 1] + 8)
        new[world] = H
        for r in range(world - 1, 0, -1):
            new[r] = min(new[r], new[r + 1] - 8)
        balance_log.append({"starts": [int(v) for v in starts], "mrays_per_band": [round(v / 1e6, 3) for v in rays]})
        starts = new
    pb, band, lb = make_group(starts)
    ok = 1.0
    if verify:
        # parity of the group path: my band of frames 0-3 equals the same rows of a whole-frame context, bit for bit
        whole.reset_history()
        lw2 = FrameLoop(snk, sc, whole, stream, settings)
        for f in range(4):
            lw2.step(); lb.step()
            torch.cuda.synchronize()
            for name in ("radiance", "reservoirs_final"):
                a = pb.read_aux(name)[band[0]:band[1]]
                b = whole.read_aux(name)[band[0]:band[1]]
                if a.tobytes() != b.tobytes():
                    ok = 0.0
        okt = torch.tensor([ok], device=f"cuda:{local}")
        dist.all_reduce(okt, op=dist.ReduceOp.MIN)
        ok = float(okt[0])
        if ok != 1.0:
            raise SystemExit("bench.py: a row band of the NCCL group differs from the same rows of the whole frame")
    for _ in range(args.warmup):
        lb.step()
    ms_b, rays_b, launches_b = timed(torch, lb, steps, barrier)
    # every rank's own time and ray count (bands differ in content: load balance) beside the max that defines the frame
    mine = torch.tensor([ms_b / steps, rays_b / steps], device=f"cuda:{local}", dtype=torch.float64)
    per_rank = [torch.zeros_like(mine) for _ in range(world)]
    dist.all_gather(per_rank, mine)
    ms_b, rays_b, launches_b = reduce_max_sum([ms_b, float(rays_b), float(launches_b)])
    ms_b /= steps
    sent = pb.group_info()[2]
    pb.group_detach()
    pb.close()
    if own_whole:
        whole.close()
    return {"n_gpus": world, "resolution": [W, H], "ms_per_frame": ms_b, "ms_per_frame_1gpu": ms_w, "efficiency_vs_n1": ms_w / (world * ms_b),
            "speedup": ms_w / ms_b, "mrays_per_s": rays_b / steps / (ms_b * 1e-3) / 1e6, "gpu_launches_per_frame_all_ranks": launches_b / steps,
            "verified_bit_exact_vs_whole_frame": bool(verify), "nccl_bytes_sent_rank0_total": int(sent), "steps": steps,
            "per_rank_ms": [round(float(t[0]), 4) for t in per_rank], "per_rank_mrays": [round(float(t[1]) / 1e6, 3) for t in per_rank],
            "band_starts": [int(v) for v in starts], "band_balancing": balance_log}


def other_configs(args, snk, torch, stream):
    """one GPU: the north-star frame on the other BASELINE scenes + the C5 build / refit / TLAS sweep"""
    out = {}
    settings = snk.api.default_settings(mode=snk.api.MODE_NORTHSTAR)

    def barrier():
        torch.cuda.synchronize()

    for name, (w, h) in (("c2m", (1920, 1080)), ("c3", (1920, 1080)), ("shipped", (1920, 1080)), ("c4", (3840, 2160)), ("c4flat", (3840, 2160))):
        sc, desc = make_scene(snk, name, w, h)
        pt = snk.PathTracer(w, h)
        pt.upload_scene(sc)
        loop = FrameLoop(snk, sc, pt, stream, settings)
        for _ in range(3):
            loop.step()
        steps = 10 if not name.startswith("c4") else 6
        ms, rays, _ = timed(torch, loop, steps, barrier)
        solo = serialized_kernels(loop, frames=2)
        counted = counted_work(loop, frames=1)
        out[name] = {"workload": desc, "resolution": [w, h], "mode": "northstar", "ms_per_frame": ms / steps, "mrays_per_s": rays / (ms * 1e-3) / 1e6,
                     "rays_per_frame": rays / steps, "traversal": traversal_table(solo, counted), "steps": steps}
        pt.close()
    # C5: BLAS build / refit over flat soups, TLAS rebuild over rigidly animated instances (TLAS.cpp:109-186, TlasSystem.cpp:27-49)
    S = snk.scenes
    peak, _ = measured_peaks()
    c5 = {"blas": [], "tlas": []}
    for n in (1_000_000, 10_000_000):
        mesh = S.c5_soup(n)
        pt = snk.PathTracer(64, 64)
        entry = pt.upload_mesh(mesh)
        blas = int(entry["blas_id"][0])
        builds = []
        for _ in range(5):
            pt._ck(pt.L.snkpt_rebuild_blas(pt.h, blas, None), "snkpt_rebuild_blas")
            builds.append(float(pt.stats()["ms_build_last"]))
        pos = mesh.positions.copy()
        pos[:, 1] += 0.05 * np.sin(pos[:, 0] * 3.0).astype(np.float32)
        pt._ck(pt.L.snkpt_update_positions(pt.h, snk.api._p(entry), snk.api._p(pos), pos.shape[0]), "snkpt_update_positions")
        refits = []
        for _ in range(5):
            pt._ck(pt.L.snkpt_refit_blas(pt.h, blas, None), "snkpt_refit_blas")
            refits.append(float(pt.stats()["ms_build_last"]))
        b, r, ntri = min(builds), min(refits), mesh.num_tris
        c5["blas"].append({"tris": ntri, "build_ms": round(b, 3), "build_mtris_per_s": round(ntri / b / 1e3, 1), "build_frac_hbm": round(420.0 * ntri / b / 1e6 / peak, 4),
                           "refit_ms": round(r, 3), "refit_frac_hbm": round(84.0 * ntri / r / 1e6 / peak, 4)})
        pt.close()
    for n_inst in (1024, 4096):
        rng = np.random.default_rng(5)
        sphere = S.icosphere_mesh(3)
        ents = [S.Entity(0, S.trs(rng.uniform(-40, 40, 3), rng.uniform(0, 360, 3), (1, 1, 1)), [0]) for _ in range(n_inst)]
        sc = S.Scene("tlas", 64, 64, [sphere], ents, S.make_materials([((0.8, 0.8, 0.8), 0.5, 0.0, 0)]), S.Camera((0, 0, -60), (0, 0, 1)))
        pt = snk.PathTracer(64, 64)
        flat = pt.upload_scene(sc)
        dev = []
        for f in range(8):
            tr = flat["transforms"].copy()
            tr[:, 13] += 0.1 * f  # rigid animation: translate every instance
            pt.set_transforms(tr, flat["transforms"])
            pt.build_tlas()
            pt.synchronize()
            dev.append(float(pt.stats()["ms_build_last"]))
        c5["tlas"].append({"instances": n_inst, "rebuild_ms": round(min(dev), 3)})
        pt.close()
    c5["note"] = "algorithmic bytes: 420 B/triangle build, 84 B/triangle refit (SURVEY §8d); min of 5 runs, device time from the library's events"
    out["c5"] = c5
    return out


JSON_OUT = sys.stdout


def emit(line):
    """the ONE JSON line of the contract, on the process's real stdout"""
    JSON_OUT.write(json.dumps(line) + "\n")
    JSON_OUT.flush()


def main():
    # stdout carries the JSON line and nothing else: libraries print there too (NCCL's version banner under
    # NCCL_DEBUG=VERSION, which the GPU boxes set), so file descriptor 1 is pointed at stderr for everything but emit()
    global JSON_OUT
    sys.stdout.flush()
    JSON_OUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--mode", default="northstar", choices=["reference", "northstar"],
                    help="northstar = BASELINE.json's metric (1 spp, 4 bounces + ReSTIR DI); reference = the reference's own frame")
    ap.add_argument("--single-mode", action="store_true", help="do not also measure the other estimator mode")
    ap.add_argument("--scene", default="c2")
    ap.add_argument("--width", type=int, default=1920)
    ap.add_argument("--height", type=int, default=1080)
    ap.add_argument("--shard", default="samples", choices=["samples"], help="N > 1: `value` is sample-sharded; the row-band frame is reported under `strong`")
    ap.add_argument("--ref-rows", type=int, default=256, help="rows of the frame the CPU oracle renders per reference-arm step")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-other-configs", action="store_true", help="N = 1: skip the C2-1M / C3 / C4 / C5 sub-records")
    ap.add_argument("--no-strong", action="store_true", help="N > 1: skip the row-band (strong scaling) sub-record")
    ap.add_argument("--equal-bands", action="store_true", help="N > 1: keep the row bands equal instead of balancing them by traced rays")
    args = ap.parse_args()
    snk = importlib.import_module("snake-vk_b200")
    if args.impl == "reference":
        run_reference_arm(args, snk)
    else:
        run_gpu_arm(args, snk)


if __name__ == "__main__":
    main()
