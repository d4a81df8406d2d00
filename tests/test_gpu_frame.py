"""GPU parity of the reference-mode frame (ReSTIR pass 1 + pass 2 + BounceDiffuse) against the CPU oracle.

Both sides follow the same float32 numerics contract (oracle/snk_math.h, csrc/vecmath.cuh), so reservoirs and
radiance are compared bit-for-bit; the only admitted differences are pixels where a float64 sin/cos/tan rounds
differently between glibc and CUDA (probability ~1e-8 per call)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def bits_equal(a, b):
    a, b = np.ascontiguousarray(a), np.ascontiguousarray(b)
    if a.dtype.fields:
        eq = np.ones(a.shape, bool)
        for name in a.dtype.names:
            if name == "pad":
                continue
            x, y = a[name], b[name]
            e = (x.view(np.uint32) == y.view(np.uint32)) | ((x != x) & (y != y)) if x.dtype.kind == "f" else (x == y)
            eq &= e.reshape(a.shape + (-1,)).all(-1)
        return eq
    e = (a.view(np.uint32) == b.view(np.uint32)) | ((a != a) & (b != b))
    return e.reshape(a.shape[:2] + (-1,)).all(-1)


def run_both(snk, oracle, sc, frames, flags=0, cam_step=(0.05, 0.0, 0.1)):
    osc = oracle.OracleScene(sc)
    fr = oracle.OracleFrame(osc)
    rt = snk.RT().Init(sc)
    prev = None
    out = []
    for f in range(frames):
        pos = np.asarray(sc.camera.pos, np.float32) + np.asarray(cam_step, np.float32) * f
        ubo = sc.ubo(f, cam_pos=pos)
        p = ubo if prev is None else prev
        fr.set(ubo, p, oracle.default_settings(flags=flags))
        fr.gbuffer()
        fr.render()
        rt.RecordRenderCmdBuf(ubo, p, flags=flags)
        out.append((rt.pt, fr))
        prev = ubo
        yield f, rt.pt, fr


@pytest.mark.parametrize("scene_name", ["small", "cornell_pl"])
def test_reference_frame_matches_oracle(snk, oracle, scene_name):
    sc = snk.scenes.small_test_scene(96, 64) if scene_name == "small" else snk.scenes.cornell(96, 96, with_pointlight=True)
    for f, pt, fr in run_both(snk, oracle, sc, frames=3):
        for name in ["prim", "depth", "normal", "motion"]:
            assert np.array_equal(pt.read_aux(name).view(np.uint8), fr.read(name).view(np.uint8)), f"frame {f}: G-buffer {name}"
        for half in ["reservoirs_temporal", "reservoirs_final"]:
            eq = bits_equal(pt.read_aux(half), fr.read(half))
            assert eq.mean() >= 0.999, f"frame {f}: {half}: {(~eq).sum()} of {eq.size} reservoirs differ"
        d_eq = bits_equal(pt.read_aux("direct")[..., :3], fr.read("direct")[..., :3])
        assert d_eq.mean() >= 0.999, f"frame {f}: direct light differs on {(~d_eq).sum()} pixels"
        a, b = pt.read_aux("radiance"), fr.read("radiance")
        r_eq = bits_equal(a[..., :3], b[..., :3])
        assert r_eq.mean() >= 0.998, f"frame {f}: radiance differs on {(~r_eq).sum()} of {r_eq.size} pixels"
        # the fp16 image the engine would blit
        o16 = pt.read_output("f16").astype(np.float32)
        assert np.allclose(o16[..., :3], np.clip(a[..., :3], -65504, 65504).astype(np.float16).astype(np.float32), equal_nan=True)
        st = pt.stats()
        oc, oa = fr.ray_counts(reset=True)
        assert int(st["stack_overflows"]) == 0
        assert int(st["rays_secondary"]) == oc, "closest-hit ray count differs from the oracle's"
        assert int(st["rays_shadow"]) == oa, "any-hit ray count differs from the oracle's"
        if f == 2:
            res = fr.read("reservoirs_final")
            assert (res["pl_idx"] != 0xFFFFFFFF).mean() > 0.01, "temporal/spatial reuse produced no point-light samples"


def test_ray_budget_closed_box(snk, oracle):
    """SURVEY §8c KAT 7: <= 65 rays per pixel (24 closest + 41 any); the GPU counts equal the oracle's."""
    sc = snk.scenes.cornell(64, 64, with_pointlight=True)
    for f, pt, fr in run_both(snk, oracle, sc, frames=2):
        st = pt.stats()
        npix = 64 * 64
        assert int(st["rays_secondary"]) <= 24 * npix and int(st["rays_shadow"]) <= 41 * npix
        oc, oa = fr.ray_counts(reset=True)
        assert (int(st["rays_secondary"]), int(st["rays_shadow"])) == (oc, oa)


def test_flags_no_restir_no_bounce(snk, oracle):
    sc = snk.scenes.small_test_scene(64, 48)
    for flags in (1, 2, 3):
        for f, pt, fr in run_both(snk, oracle, sc, frames=2, flags=flags):
            eq = bits_equal(pt.read_aux("radiance")[..., :3], fr.read("radiance")[..., :3])
            assert eq.mean() >= 0.999, f"flags {flags} frame {f}: {(~eq).sum()} pixels differ"


def test_async_uploads_and_readback_pipeline(snk, oracle):
    """per-frame uploads without stream synchronisation + two read-backs in flight give the same images as the blocking calls"""
    sc = snk.scenes.small_test_scene(96, 64)
    flat = sc.flatten()
    ref = snk.RT().Init(sc)
    rt = snk.RT().Init(sc)
    bufs = [np.zeros((64, 96, 4), np.float16) for _ in range(2)]
    want, got = [], []
    prev = None
    for f in range(5):
        pos = np.asarray(sc.camera.pos, np.float32) + np.asarray((0.05, 0.0, 0.1), np.float32) * f
        ubo = sc.ubo(f, cam_pos=pos)
        p = ubo if prev is None else prev
        prev = ubo
        ref.RecordRenderCmdBuf(ubo, p, mode=snk.api.MODE_NORTHSTAR)
        want.append(ref.pt.read_output("f16").copy())
        # the engine's per-frame buffer updates, then the frame, then a non-blocking read-back
        rt.pt.set_transforms(flat["transforms"], flat["prev_transforms"])
        rt.pt.set_instances(flat["instances"], flat["blas_of_instance"], flat["emissive_idx"])
        rt.pt.set_lights(flat["lights"])
        rt.pt.set_materials(flat["materials"])
        rt.pt.build_tlas()
        rt.RecordRenderCmdBuf(ubo, p, mode=snk.api.MODE_NORTHSTAR)
        rt.pt.read_output_async(bufs[f % 2], slot=f % 2)
        if f:
            rt.pt.wait_output((f - 1) % 2)
            got.append(bufs[(f - 1) % 2].copy())
    rt.pt.wait_output(4 % 2)
    got.append(bufs[4 % 2].copy())
    for f in range(5):
        assert np.array_equal(want[f].view(np.uint16), got[f].view(np.uint16)), f"frame {f}"


@pytest.mark.parametrize("mode", ["reference", "northstar"])
def test_resize_equals_fresh_context(snk, oracle, mode):
    """snkpt_resize (window resize, EditorLayer.cpp:351-366): all per-pixel state is recreated and the history is invalid, so the
    frames after a resize equal those of a context created at the new size — image, reservoirs and ray counts, bit for bit —
    and a read-back that was still in flight at the old size does not disturb it."""
    mode_id = snk.api.MODE_REFERENCE if mode == "reference" else snk.api.MODE_NORTHSTAR
    small, big = snk.scenes.small_test_scene(64, 48), snk.scenes.small_test_scene(96, 64)

    def frames(rt, sc, n):
        out, prev = [], None
        for f in range(n):
            ubo = sc.ubo(f)
            rt.RecordRenderCmdBuf(ubo, ubo if prev is None else prev, mode=mode_id)
            prev = ubo
            st = rt.pt.stats()
            out.append((rt.pt.read_output("f32").copy(), rt.pt.read_aux("reservoirs_final").copy(), rt.pt.read_aux("reservoirs_temporal").copy(),
                        int(st["rays_secondary"]), int(st["rays_shadow"])))
        return out

    rt = snk.RT().Init(small)
    frames(rt, small, 2)                      # history at the old size
    pending = np.empty((48, 64, 4), np.float16)
    rt.pt.read_output_async(pending, slot=0)  # still in flight when the resize arrives
    rt.pt.resize(96, 64)
    got = frames(rt, big, 3)
    fresh = snk.RT().Init(big)
    want = frames(fresh, big, 3)
    for f, (g, w) in enumerate(zip(got, want)):
        assert g[0].shape == (64, 96, 4)
        assert np.array_equal(g[0].view(np.uint32), w[0].view(np.uint32)), f"frame {f} after the resize differs from a fresh context"
        assert np.array_equal(g[1].view(np.uint8), w[1].view(np.uint8)) and np.array_equal(g[2].view(np.uint8), w[2].view(np.uint8))
        assert g[3:] == w[3:]
    # back to the first size: equal to what a fresh small context renders
    rt.pt.resize(64, 48)
    again = frames(rt, small, 2)
    ref = frames(snk.RT().Init(small), small, 2)
    for g, w in zip(again, ref):
        assert np.array_equal(g[0].view(np.uint32), w[0].view(np.uint32))
    with pytest.raises(snk.SnkError):
        rt.pt.resize(0, 10)
    rt.pt.close(); fresh.pt.close()


def test_accumulate_rows_and_diagnostics(snk, oracle):
    """the entry points without a reference call site of their own (second table): progressive accumulation,
    a row band set after creation, raw plane pointers / sizes, per-launch timings and the order-independent work counters."""
    sc = snk.scenes.small_test_scene(64, 48)
    rt = snk.RT().Init(sc)
    pt = rt.pt
    pt.synchronize()
    # ---- snkpt_accumulate: running mean in float32, in the kernel's own operation order
    mean = None
    prev = None
    for f in range(4):
        ubo = sc.ubo(f)
        rt.RecordRenderCmdBuf(ubo, ubo if prev is None else prev, mode=snk.api.MODE_NORTHSTAR)
        prev = ubo
        pt.accumulate(reset=(f == 0))
        r = pt.read_aux("radiance")
        if f == 0:
            mean = r.copy()
        else:
            mean = (mean + (r - mean) * np.float32(np.float32(1.0) / np.float32(f + 1))).astype(np.float32)
        acc = pt.read_aux("accum")
        assert np.array_equal(acc[..., :3].view(np.uint32), mean[..., :3].view(np.uint32)), f"accumulation after frame {f}"
        out = pt.read_output("f16")
        assert np.array_equal(out[..., :3], mean[..., :3].astype(np.float16))
    # ---- raw pointers and sizes
    p, n = pt.device_ptr(0x100)
    assert p and n == 64 * 48 * 8
    p, n = pt.device_ptr(0x101)
    assert p and n == 64 * 48 * 16
    assert pt.read_raw(20, snk.api.BLAS_INFO_DTYPE).size >= len(sc.meshes)  # snkpt_aux_size + snkpt_read_aux: one BLAS per submesh
    # ---- per-launch timings (level 1) and work counters (level 2)
    pt.set_profiling(2)
    ubo = sc.ubo(4)
    rt.RecordRenderCmdBuf(ubo, prev, mode=snk.api.MODE_NORTHSTAR)
    recs = pt.launch_timings()
    names = [r["name"].decode() for r in recs]
    assert "primary" in names and any(nm.startswith("trace.q") for nm in names) and "ns_resolve" in names
    assert all(r["ms"] > 0 for r in recs)
    tr = [r for r in recs if r["is_trace"]]
    st = pt.stats()
    assert sum(int(r["live"]) for r in tr) == int(st["rays_shadow"]) + int(st["rays_secondary"])
    assert all(int(r["nodes"]) >= int(r["live"]) for r in tr if r["live"])  # every live ray tests at least the TLAS root
    assert pt.launch_timings().size == 0  # returned and cleared
    pt.set_profiling(0)
    # snkpt_count_work replays single rays: its sums equal the launch record of the same rays traced at level 2
    rng = np.random.default_rng(3)
    rays = np.zeros(5000, snk.scenes.RAY_DTYPE)
    rays["o"] = rng.uniform((-8, 0.5, -8), (8, 6, 8), (5000, 3)).astype(np.float32)
    d = rng.normal(size=(5000, 3)); rays["d"] = (d / np.linalg.norm(d, axis=1, keepdims=True)).astype(np.float32)
    rays["tmin"], rays["tmax"] = 0.001, 1000.0
    pt.set_profiling(2)
    hits = pt.trace_rays(rays)
    rec = [r for r in pt.launch_timings() if r["is_trace"]]
    pt.set_profiling(0)
    work = pt.count_work(rays, hits)
    assert len(rec) == 1 and int(rec[0]["live"]) == 5000
    assert (int(rec[0]["nodes"]), int(rec[0]["tris"]), int(rec[0]["insts"])) == tuple(int(v) for v in work.sum(0, dtype=np.uint64))
    assert (work[:, 0] >= 1).all() and (work[hits["t"] >= 0, 1] >= 1).all()
    # ---- snkpt_set_rows after creation: without ReSTIR every pixel is independent, so a band equals those rows of the frame
    flags = snk.api.RT_NO_RESTIR
    ubo = sc.ubo(7)
    pt.reset_history()
    rt.RecordRenderCmdBuf(ubo, ubo, mode=snk.api.MODE_NORTHSTAR, flags=flags)
    full = pt.read_aux("radiance").copy()
    pt.set_rows(10, 30)
    pt.reset_history()
    rt.RecordRenderCmdBuf(ubo, ubo, mode=snk.api.MODE_NORTHSTAR, flags=flags)
    band = pt.read_aux("radiance")
    assert np.array_equal(band[10:30].view(np.uint32), full[10:30].view(np.uint32))
    assert pt.read_output("f16").shape == (20, 64, 4)
    with pytest.raises(snk.SnkError):
        pt.set_rows(30, 10)
    pt.close()
