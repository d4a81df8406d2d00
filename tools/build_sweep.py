"""C5: BLAS full build / refit over flat soups of N triangles, TLAS rebuild over rigidly animated instances.
usage: python tools/build_sweep.py [max_tris]   → JSON lines on stdout"""
import importlib, json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
snk = importlib.import_module("snake-vk_b200")
S = snk.scenes
max_tris = int(float(sys.argv[1])) if len(sys.argv) > 1 else 10_000_000
PEAK = 6650.0

def soup_scene(mesh):
    return S.Scene("soup", 64, 64, [mesh], [S.Entity(0, np.eye(4, dtype=np.float32), [0])], S.make_materials([((0.8, 0.8, 0.8), 0.5, 0.0, 0)]),
                   S.Camera((0, 0, -50), (0, 0, 1)))

for n in [100_000, 300_000, 1_000_000, 3_000_000, 10_000_000]:
    if n > max_tris:
        break
    if os.environ.get("ONLY") and int(float(os.environ["ONLY"])) != n:
        continue
    t0 = time.time()
    mesh = S.c5_soup(n)
    gen_s = time.time() - t0
    pt = snk.PathTracer(64, 64)
    t0 = time.time()
    entry = pt.upload_mesh(mesh)
    up_s = time.time() - t0
    blas = int(entry["blas_id"][0])
    builds = []
    for _ in range(5):
        pt._ck(pt.L.snkpt_rebuild_blas(pt.h, blas, None), "rebuild")
        builds.append(float(pt.stats()["ms_build_last"]))
    ntri = mesh.num_tris
    tbl = pt.blas_table()[blas]
    # refit after a per-vertex sinusoidal displacement
    pos = mesh.positions.copy()
    pos[:, 1] += 0.05 * np.sin(pos[:, 0] * 3.0).astype(np.float32)
    pt._ck(pt.L.snkpt_update_positions(pt.h, entry.ctypes.data_as(__import__("ctypes").c_void_p), pos.ctypes.data_as(__import__("ctypes").c_void_p), pos.shape[0]), "update")
    refits = []
    for _ in range(5):
        pt._ck(pt.L.snkpt_refit_blas(pt.h, blas, None), "refit")
        refits.append(float(pt.stats()["ms_build_last"]))
    b, r = min(builds), min(refits)
    print(json.dumps({"what": "blas", "tris": ntri, "nodes": int(tbl["node_count"]), "build_ms": round(b, 3), "build_mtris_s": round(ntri / b / 1e3, 1),
                      "build_alg_gbs": round(420.0 * ntri / b / 1e6, 1), "build_frac_hbm": round(420.0 * ntri / b / 1e6 / PEAK, 4),
                      "refit_ms": round(r, 3), "refit_alg_gbs": round(84.0 * ntri / r / 1e6, 1), "refit_frac_hbm": round(84.0 * ntri / r / 1e6 / PEAK, 4),
                      "host_gen_s": round(gen_s, 2), "upload_s": round(up_s, 2)}), flush=True)
    pt.close()

# TLAS rebuild: n rigidly animated instances of one sphere BLAS
for n_inst in (() if os.environ.get("ONLY") else (256, 1024, 4096)):
    rng = np.random.default_rng(5)
    sphere = S.icosphere_mesh(3)
    ents = [S.Entity(0, S.trs(rng.uniform(-40, 40, 3), rng.uniform(0, 360, 3), (1, 1, 1)), [0]) for _ in range(n_inst)]
    sc = S.Scene("tlas", 64, 64, [sphere], ents, S.make_materials([((0.8, 0.8, 0.8), 0.5, 0.0, 0)]), S.Camera((0, 0, -60), (0, 0, 1)))
    pt = snk.PathTracer(64, 64)
    flat = pt.upload_scene(sc)
    dev, wall = [], []
    for f in range(8):
        tr = flat["transforms"].copy()
        tr[:, 13] += 0.1 * f  # rigid animation: translate every instance
        t0 = time.perf_counter()
        pt.set_transforms(tr, flat["transforms"])
        pt.build_tlas()
        pt.synchronize()
        wall.append((time.perf_counter() - t0) * 1e3)
        dev.append(float(pt.stats()["ms_build_last"]))
    print(json.dumps({"what": "tlas", "instances": n_inst, "device_ms": round(min(dev), 3), "wall_ms_incl_upload": round(min(wall), 3)}), flush=True)
    pt.close()
