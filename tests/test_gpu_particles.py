"""Particle collision pass (snkpt_particles_*, csrc/particles.cu) against the CPU statement (oracle/snk_particles.h): Init,
and several frames of three sub-steps of particles raining onto the instanced test scene — particle buffers and ray counts
must agree bit for bit."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def rain(oracle, n, seed=5):
    rng = np.random.default_rng(seed)
    p = np.zeros(n, oracle.PARTICLE_DTYPE)
    p["position_radius"][:, 0] = rng.uniform(-5, 5, n)
    p["position_radius"][:, 1] = rng.uniform(0.3, 5.0, n)
    p["position_radius"][:, 2] = rng.uniform(-3, 3, n)
    p["position_radius"][:, 3] = 0.15
    p["velocity"][:, :3] = rng.uniform(-1.5, 1.5, (n, 3))
    p["velocity"][:, 3] = 1.0
    p["colour"][:, :3] = rng.uniform(0, 1, (n, 3))
    return p


def same(a, b):
    return (a.view(np.uint32).reshape(a.size, 12) == b.view(np.uint32).reshape(b.size, 12)).all(-1)


def test_init_matches_oracle(snk, oracle):
    pt = snk.PathTracer(32, 32)
    pt.particles_create(40000, init=True)
    assert same(pt.particles_read(), oracle.particles_init(40000)).all()
    pt.close()


def test_particle_updates_match_oracle(snk, oracle):
    sc = snk.scenes.small_test_scene(64, 48)
    osc = oracle.OracleScene(sc)
    rt = snk.RT().Init(sc)
    n = 6000
    p = rain(oracle, n)
    rt.pt.particles_create(n, init=False)
    rt.pt.particles_write(p.view(snk.api.PARTICLE_DTYPE))
    rays_cpu = 0
    for frame in range(4):
        rt.pt.particles_update(substeps=3)
        p, r = oracle.particles_update(osc, p, substeps=3)
        rays_cpu += r
        got = rt.pt.particles_read().view(oracle.PARTICLE_DTYPE)
        eq = same(got, p)
        assert eq.all(), f"frame {frame}: {(~eq).sum()} of {n} particles differ; first {np.nonzero(~eq)[0][:5]}"
        assert np.isfinite(got["position_radius"]).all()
    assert rt.pt.particles_rays() == rays_cpu
    moved = np.abs(p["position_radius"][:, :3] - rain(oracle, n)["position_radius"][:, :3]).max(-1)
    assert (moved > 1e-3).mean() > 0.9 and p["position_radius"][:, 1].min() > -0.2, "particles moved and none fell through the floor"
    # custom parameters reach the kernel
    prm = snk.api.physics_params(friction_coefficient=0.9, particle_restitution=0.6)
    rt.pt.particles_update(prm, substeps=2)
    p2, _ = oracle.particles_update(osc, p, prm, substeps=2)
    assert same(rt.pt.particles_read().view(oracle.PARTICLE_DTYPE), p2).all()


def test_particles_need_a_tlas(snk):
    pt = snk.PathTracer(32, 32)
    with pytest.raises(RuntimeError):
        pt.particles_update()
    pt.particles_create(16)
    with pytest.raises(RuntimeError):
        pt.particles_update()
    pt.close()
