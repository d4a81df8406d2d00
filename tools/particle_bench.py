"""Particle collision pass on C2: particles/s and rays/s of ParticleSystem::UpdateParticles (3 sub-steps per
frame, the reference's 100 000 particles by default) + the CPU oracle on a sample.  usage: python tools/particle_bench.py [count]"""
import importlib, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import bench
snk = importlib.import_module("snake-vk_b200")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000
sc, desc = bench.make_scene(snk, "c2", 1920, 1080)
pt = snk.PathTracer(64, 64)
pt.upload_scene(sc)
rng = np.random.default_rng(1)
p = np.zeros(n, snk.api.PARTICLE_DTYPE)
p["position_radius"][:, 0] = rng.uniform(-30, 30, n); p["position_radius"][:, 1] = rng.uniform(0.5, 12, n); p["position_radius"][:, 2] = rng.uniform(-15, 15, n)
p["position_radius"][:, 3] = 0.15
p["velocity"][:, :3] = rng.uniform(-2, 2, (n, 3)); p["velocity"][:, 3] = 1
pt.particles_create(n, init=False)
pt.particles_write(p)
for _ in range(5):
    pt.particles_update(substeps=3)
pt.synchronize()
r0 = pt.particles_rays()
frames = 30
pt.set_profiling(1)
t0 = time.time()
for _ in range(frames):
    pt.particles_update(substeps=3)
pt.synchronize()
wall = (time.time() - t0) / frames * 1e3
rays = (pt.particles_rays() - r0) / frames
upd = [float(r["ms"]) for r in pt.launch_timings() if r["name"].decode().startswith("ptcl_update")]
awake = float((pt.particles_read()["velocity"][:, 3] > 0.5).mean())
print(f"particles {n} on {desc}: frame (3 sub-steps) {wall:.3f} ms wall, update kernel {np.mean(upd):.3f} ms per sub-step, "
      f"{n * 3 / wall / 1e3:.1f} M particle-steps/s, {rays / 1e6:.2f} M rays/frame -> {rays / wall / 1e3:.1f} Mrays/s; awake {awake:.2f}")
if os.environ.get("CPU"):
    import oracle_api as O
    osc = O.OracleScene(sc)
    m = min(n, 20000)
    q = p[:m].view(O.PARTICLE_DTYPE)
    t0 = time.time()
    q, r = O.particles_update(osc, q, substeps=3, threads=os.cpu_count() or 1)
    s = time.time() - t0
    print(f"CPU oracle ({os.cpu_count()} threads): {m} particles x 3 sub-steps in {s * 1e3:.1f} ms -> {m * 3 / s / 1e6:.2f} M particle-steps/s, {r / s / 1e6:.2f} Mrays/s")
