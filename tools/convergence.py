"""Statistical check of the north-star estimator on C1 (Cornell box, 512x512; BASELINE.json config 0): 4096 accumulated frames of
(a) the plain path tracer (next-event estimation at x0, no ReSTIR) with two different seeds — the noise floor,
(b) ReSTIR DI with initial candidates + visibility only, (c) full ReSTIR DI (temporal 1, spatial 3),
compared per channel by relative MSE = mean((a - b)^2) / mean(b)^2 over the non-emissive pixels (north_star: < 1e-3).
usage: python tools/convergence.py [spp]"""
import importlib, json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
snk = importlib.import_module("snake-vk_b200")
spp = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
W = H = 512
sc = snk.scenes.cornell(W, H, with_pointlight=True)


def mean_image(seed0, **kw):
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


plain_a, body, ms_a = mean_image(0, flags=1)
plain_b, _, _ = mean_image(100000, flags=1)
init, _, ms_i = mean_image(200000, num_temporal=0, num_spatial=0)
full, _, ms_f = mean_image(300000)
out = {"scene": "C1 Cornell 512x512, point light + emissive ceiling quad, 4 bounces", "spp": spp,
       "ms_per_frame": {"plain": round(ms_a, 3), "restir_initial": round(ms_i, 3), "restir_full": round(ms_f, 3)},
       "relMSE_plain_vs_plain_other_seed": rmse(plain_a, plain_b, body),
       "relMSE_restir_initial_only_vs_plain": rmse(init, plain_b, body),
       "relMSE_restir_full_vs_plain": rmse(full, plain_b, body),
       "mean_radiance": {"plain": plain_b[body].mean(0).tolist(), "restir_initial": init[body].mean(0).tolist(), "restir_full": full[body].mean(0).tolist()}}
print(json.dumps(out, indent=1))
