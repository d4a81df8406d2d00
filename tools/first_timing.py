import sys, importlib, time, numpy as np
sys.path.insert(0, '.')
snk = importlib.import_module("snake-vk_b200")
name = sys.argv[1] if len(sys.argv) > 1 else "c2"
sc = {"c2": lambda: snk.scenes.c2("instanced"), "c2flat": lambda: snk.scenes.c2("flat"), "c3": snk.scenes.c3,
      "c2m": lambda: snk.scenes.c2("flat", subdiv=4), "cornell": lambda: snk.scenes.cornell(512, 512, True)}[name]()
t = time.time(); rt = snk.RT().Init(sc); print("upload+build s", time.time() - t, "tris", sc.num_tris)
print("last build ms", float(rt.pt.stats()["ms_build_last"]), rt.pt.blas_table()[["node_count", "tri_count"]][:4])
prev = None
for f in range(4):
    ubo = sc.ubo(f, cam_pos=np.asarray(sc.camera.pos, np.float32) + np.float32(0.1 * f) * np.array([0, 0, 1], np.float32))
    rt.RecordRenderCmdBuf(ubo, ubo if prev is None else prev); prev = ubo
    st = rt.pt.stats()
    rays = int(st["rays_primary"]) + int(st["rays_shadow"]) + int(st["rays_secondary"])
    ms = float(st["ms_primary"]) + float(st["ms_total"])
    print(f"frame {f}: primary {float(st['ms_primary']):.3f} ms, restir1 {float(st['ms_restir1']):.3f}, restir2 {float(st['ms_restir2']):.3f}, "
          f"bounce {float(st['ms_bounce']):.3f}, rays {rays/1e6:.1f} M ({int(st['rays_shadow'])/1e6:.1f} shadow, {int(st['rays_secondary'])/1e6:.1f} closest) -> {rays/ms/1e3:.1f} Mrays/s, overflow {int(st['stack_overflows'])}")
img = rt.pt.read_aux("radiance")[..., :3]
print("mean", img.mean(axis=(0, 1)), "nan", np.isnan(img).sum())
try:
    from PIL import Image
    t = np.clip(img / (1 + img), 0, 1) ** (1 / 2.2)
    Image.fromarray((t * 255).astype(np.uint8)).save(f"gpurun_out/{name}.png")
except Exception as e:
    print("no image:", e)
