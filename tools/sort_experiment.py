"""Does ray ordering matter for the traversal kernel? Traces the same incoherent secondary rays of C2 / C2 flat in several
orders through snkpt_trace_rays (device time from the per-launch CUDA events) — the measurement behind DESIGN §4.2's
ray-sorting entry.  usage: python tools/sort_experiment.py [c2|c2flat ...]"""
import importlib, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
snk = importlib.import_module("snake-vk_b200")
S = snk.scenes


def unit(v):
    return v / np.linalg.norm(v, axis=-1, keepdims=True)


def camera_rays(sc, W, H):
    fwd = unit(np.asarray(sc.camera.forward, np.float64)); up0 = np.array([0.0, 1.0, 0.0])
    right = unit(np.cross(fwd, up0)); up = np.cross(right, fwd)
    th = np.tan(np.radians(sc.camera.fov * 0.25))
    xs, ys = np.meshgrid((np.arange(W) + 0.5) / W * 2 - 1, 1 - (np.arange(H) + 0.5) / H * 2)
    d = unit(fwd + xs[..., None] * th * W / H * right + ys[..., None] * th * up).reshape(-1, 3)
    r = np.zeros(W * H, S.RAY_DTYPE)
    r["o"] = np.asarray(sc.camera.pos, np.float32); r["d"] = d.astype(np.float32); r["tmin"] = 0.01; r["tmax"] = 1000.0
    return r


def bounce(rng, rays, hits):
    """cosine-distributed directions around -d at every hit (a stand-in for the diffuse bounce)"""
    ok = hits["t"] >= 0
    o = rays["o"][ok].astype(np.float64) + rays["d"][ok].astype(np.float64) * hits["t"][ok][:, None]
    n = -rays["d"][ok].astype(np.float64)
    u1, u2 = rng.random(o.shape[0]), rng.random(o.shape[0])
    a = np.where(np.abs(n[:, :1]) > 0.9, [[0.0, 1.0, 0.0]], [[1.0, 0.0, 0.0]])
    t = unit(np.cross(a, n)); b = np.cross(n, t)
    r_, phi = np.sqrt(u1), 2 * np.pi * u2
    d = unit(t * (r_ * np.cos(phi))[:, None] + b * (r_ * np.sin(phi))[:, None] + n * np.sqrt(1 - u1)[:, None])
    out = np.zeros(o.shape[0], S.RAY_DTYPE)
    out["o"] = (o + n * 0.01).astype(np.float32); out["d"] = d.astype(np.float32); out["tmin"] = 0.001; out["tmax"] = 1000.0
    return out


def morton30(p):
    lo, hi = p.min(0), p.max(0)
    q = np.clip(((p - lo) / np.maximum(hi - lo, 1e-9) * 1024).astype(np.int64), 0, 1023)
    def spread(x):
        x = (x | (x << 16)) & 0x030000FF; x = (x | (x << 8)) & 0x0300F00F; x = (x | (x << 4)) & 0x030C30C3; x = (x | (x << 2)) & 0x09249249
        return x
    return (spread(q[:, 0]) << 2) | (spread(q[:, 1]) << 1) | spread(q[:, 2])


def timed_trace(pt, rays, any_hit=False):
    best = 1e9
    for _ in range(3):
        pt.set_profiling(1)
        h = pt.trace_rays(rays, any_hit=any_hit)
        t = pt.launch_timings()
        ms = float(sum(r["ms"] for r in t if r["is_trace"]))
        best = min(best, ms)
    return best, h


for name in (sys.argv[1:] or ["c2", "c2flat"]):
    sc = S.c2("flat", 4) if name == "c2flat" else (S.c3() if name == "c3" else S.c2())
    pt = snk.PathTracer(sc.width, sc.height)
    pt.upload_scene(sc)
    rng = np.random.default_rng(5)
    r0 = camera_rays(sc, sc.width, sc.height)
    ms0, h0 = timed_trace(pt, r0)
    r1 = bounce(rng, r0, h0)
    ms1, h1 = timed_trace(pt, r1)
    r2 = bounce(rng, r1, h1)
    print(f"== {name}: primary {r0.size/ms0/1e6:.2f} Grays/s ({ms0:.3f} ms), bounce 1 in pixel order {r1.size/ms1/1e6:.2f} Grays/s ({r1.size} rays)")
    for label, rays in (("bounce 1", r1), ("bounce 2", r2)):
        octant = (rays["d"][:, 0] < 0).astype(np.int64) | ((rays["d"][:, 1] < 0).astype(np.int64) << 1) | ((rays["d"][:, 2] < 0).astype(np.int64) << 2)
        mo = morton30(rays["o"].astype(np.float64))
        tip = morton30(rays["o"].astype(np.float64) + rays["d"].astype(np.float64) * 2.0)
        orders = {
            "producer order (pixel / path order)": np.arange(rays.size),
            "random shuffle": rng.permutation(rays.size),
            "direction octant": np.argsort(octant, kind="stable"),
            "origin Morton (30 bit)": np.argsort(mo, kind="stable"),
            "octant, then origin Morton": np.argsort((octant << 30) | mo, kind="stable"),
            "origin cell (15 bit) then Morton of a point 2 units along the ray": np.argsort(((mo >> 15) << 30) | tip, kind="stable"),
        }
        base = None
        for k, order in orders.items():
            for any_hit in (False, True):
                ms, _ = timed_trace(pt, rays[order], any_hit)
                if base is None:
                    base = {}
                base.setdefault(any_hit, ms)
                print(f"   {label:9s} {'any    ' if any_hit else 'closest'} {k:66s} {ms:7.3f} ms  {rays.size/ms/1e6:5.2f} Grays/s  x{base[any_hit]/ms:4.2f}")
    pt.close()
