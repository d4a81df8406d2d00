"""Statistical check of the north-star estimator on C1 (Cornell box, 512x512; BASELINE.json config 0): relative MSE per channel
= mean((a - ref)^2) / mean(ref)^2 over the non-emissive pixels (north_star: < 1e-3 at 4096 spp) of
(a) the plain path tracer (next-event estimation at x0, no ReSTIR), another seed than the reference — the Monte-Carlo noise at `spp`,
(b) ReSTIR DI with initial candidates + visibility only, (c) full ReSTIR DI (temporal 1, spatial 3; MIS weights without visibility:
biased), (d) the same with SNKPT_RT_UNBIASED_REUSE,
each against a reference of `ref_spp` frames of the plain path tracer (so the reference's own noise is far below the gate).
usage: python tools/convergence.py [spp [ref_spp [size]]]"""
import importlib, json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
snk = importlib.import_module("snake-vk_b200")


def mean_image(sc, spp, seed0, **kw):
    W, H = sc.width, sc.height
    pt = snk.PathTracer(W, H)
    pt.upload_scene(sc)
    st = snk.api.default_settings(mode=snk.api.MODE_NORTHSTAR, **kw)
    t0 = time.time()
    for f in range(spp):
        ubo = sc.ubo(seed0 + f)
        ubo["prev_frame_invalid"] = int(f == 0)
        pt.set_frame(ubo, ubo)
        if f == 0:
            pt.trace_primary()
        pt.render(st)
        pt.accumulate(reset=(f == 0))
    pt.synchronize()
    img = pt.read_aux("accum")[..., :3].astype(np.float64)
    body = ((pt.read_aux("flags") & 1) == 0) & (pt.read_aux("prim") != 0xFFFFFFFF)
    ms = (time.time() - t0) / spp * 1e3
    pt.close()
    return img, body, ms


def rmse(a, b, m):
    return [float(((a[..., c] - b[..., c])[m] ** 2).mean() / (b[..., c][m].mean() ** 2)) for c in range(3)]


MODES = {"plain": (0, dict(flags=1)), "restir_initial": (200000, dict(num_temporal=0, num_spatial=0)), "restir_full": (300000, dict()),
         "restir_unbiased": (300000, dict(flags=4))}  # flags: SNKPT_RT_NO_RESTIR = 1, SNKPT_RT_UNBIASED_REUSE = 4


def run(spp=4096, ref_spp=262144, size=512, modes=tuple(MODES)):
    sc = snk.scenes.cornell(size, size, with_pointlight=True)
    ref, body, ms_r = mean_image(sc, ref_spp, 1000000, flags=snk.api.RT_NO_RESTIR)
    out = {"scene": f"C1 Cornell {size}x{size}, point light + emissive ceiling quad, 4 bounces", "spp": spp, "reference_spp": ref_spp,
           "ms_per_frame": {"reference": round(ms_r, 3)}, "mean_radiance": {"reference": ref[body].mean(0).tolist()}}
    for m in modes:
        seed0, kw = MODES[m]
        img, _, ms = mean_image(sc, spp, seed0, **kw)
        out["ms_per_frame"][m] = round(ms, 3)
        out[f"relMSE_{m}_vs_reference"] = rmse(img, ref, body)
        out["mean_radiance"][m] = img[body].mean(0).tolist()
    return out


if __name__ == "__main__":
    a = [int(v) for v in sys.argv[1:]]
    print(json.dumps(run(*a), indent=1))
