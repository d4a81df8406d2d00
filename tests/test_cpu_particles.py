"""Particle collision pass oracle (oracle/snk_particles.h, restating Particle.glsl / Particle.comp / ParticleUpdate.rgen+rchit):
known answers for the spatial hash, the grid tables, Init, and the physics of simple situations. CPU only."""
import numpy as np
import pytest


def cell_key(cell, n):
    x, y, z = (int(c) & 0xFFFFFFFF for c in cell)
    h = ((x * 73856093) & 0xFFFFFFFF) ^ ((y * 10357) & 0xFFFFFFFF) ^ ((z * 83492791) & 0xFFFFFFFF)
    return h % n


def floor_scene(snk):
    S = snk.scenes
    floor = S.quad_mesh((-50, 0, -50), (50, 0, -50), (50, 0, 50), (-50, 0, 50))
    return S.Scene("floor", 32, 32, [floor], [S.Entity(0, np.eye(4, dtype=np.float32), [0])], S.make_materials([((0.8, 0.8, 0.8), 0.5, 0.0, 0)]),
                   S.Camera((0, 2, -6), (0, 0, 1)))


def one(oracle, pos, vel=(0, 0, 0), radius=0.15, awake=1.0, colour=(0, 0, 1)):
    p = np.zeros(1, oracle.PARTICLE_DTYPE)
    p["position_radius"][0] = (*pos, radius)
    p["velocity"][0] = (*vel, awake)
    p["colour"][0] = (*colour, 0)
    return p


def test_init_matches_the_shader(oracle):
    n = 30000
    p = oracle.particles_init(n)
    i = np.arange(n)
    grid = np.stack([i % 100, (i % 10000) // 100, i // 10000], -1).astype(np.float32)
    want = ((grid - np.float32([50, 0, 50])) * np.float32(0.5) + np.float32([-100, 20, 190])).astype(np.float32)
    assert np.array_equal(p["position_radius"][:, :3], want) and (p["position_radius"][:, 3] == np.float32(0.15)).all()
    arg = (i.astype(np.float32) * np.float32(8.34534)).astype(np.float32)  # the shader multiplies in float32 (ulp 0.016 at 250 000)
    assert np.allclose(p["velocity"][:, 0], 0.1 * np.sin(arg.astype(np.float64)), atol=2e-7)
    assert (p["velocity"][:, 3] == 1).all()
    assert np.array_equal(p["colour"][0, :3], (0, 0, 1)) and np.array_equal(p["colour"][15000, :3], (0, 1, 0)) and np.array_equal(p["colour"][29999, :3], (1, 0, 0))


def test_spatial_hash_and_grid_tables(oracle):
    rng = np.random.default_rng(2)
    n = 5000
    p = np.zeros(n, oracle.PARTICLE_DTYPE)
    p["position_radius"][:, :3] = rng.uniform(-20, 20, (n, 3)).astype(np.float32)  # negative cells too: int32 wrap-around in the hash
    keys, start = oracle.particles_grid(p)
    cells = (p["position_radius"][:, :3] * np.float32(5.0)).astype(np.int32)     # ivec3(pos * 5): truncation
    want = np.array([cell_key(c, n) for c in cells], np.uint32)
    assert np.array_equal(np.sort(want), keys[:, 0])
    order = np.argsort(want, kind="stable")                                       # ascending particle index within one key
    assert np.array_equal(keys[:, 1], order.astype(np.uint32)) and np.array_equal(keys[:, 0], want[order])
    for k in np.unique(want)[:200]:
        assert start[k] == np.searchsorted(keys[:, 0], k)
    assert (start[np.setdiff1d(np.arange(n), want)] == 0xFFFFFFFF).all()


def test_free_fall_then_rest_on_the_floor(snk, oracle):
    osc = oracle.OracleScene(floor_scene(snk))
    p = one(oracle, (0.0, 1.0, 0.0), vel=(0.2, 0.0, 0.0))
    g = np.float32(9.8) * np.float32(1.0 / 120.0)
    q, rays = oracle.particles_update(osc, p, substeps=1)
    # no contact: 3 collision rays + gravity ray + constraint ray, one gravity impulse, friction takes a little speed
    assert rays == 5 and q["velocity"][0, 1] < 0 and abs(float(q["velocity"][0, 1]) + float(g)) < 2e-3 and q["velocity"][0, 3] == 1
    assert q["position_radius"][0, 1] < 1.0 and q["position_radius"][0, 0] > 0.0
    ys = []
    for _ in range(150):
        p, _ = oracle.particles_update(osc, p, substeps=3)
        ys.append(float(p["position_radius"][0, 1]))
        assert np.isfinite(p["position_radius"]).all() and p["position_radius"][0, 1] > 0.05, "the particle must not fall through the floor"
    assert abs(ys[-1] - ys[-20]) < 1e-3 and ys[-1] < 0.3, "it comes to rest on the floor"
    assert p["velocity"][0, 3] == 0 and np.array_equal(p["velocity"][0, :3], (0, 0, 0)), "and falls asleep (velocity.w = 0)"


def padded(oracle, *ps, n=4099):
    """the hash is taken modulo the particle count (Particle.glsl:14-16): with only two particles most of the 27 neighbour cells
    share a key and the same neighbour is processed many times (a property of the reference). Pad with far-away sleepers."""
    pad = np.zeros(n - len(ps), oracle.PARTICLE_DTYPE)
    pad["position_radius"][:, 0] = 1000.0 + 3.0 * np.arange(pad.size)
    pad["position_radius"][:, 1] = 500.0
    pad["position_radius"][:, 3] = 0.15
    return np.concatenate(list(ps) + [pad])


def test_two_particles_repel_symmetrically(snk, oracle):
    osc = oracle.OracleScene(floor_scene(snk))
    p = padded(oracle, one(oracle, (-0.1, 5.0, 0.0), vel=(0.5, 0, 0), colour=(1, 0, 0)), one(oracle, (0.1, 5.0, 0.0), vel=(-0.5, 0, 0), colour=(0, 0, 1)))
    q, _ = oracle.particles_update(osc, p, substeps=1)
    assert q["velocity"][0, 0] < 0.5 and q["velocity"][1, 0] > -0.5, "approach speed is reduced / reversed"
    assert np.isclose(q["velocity"][0, 0], -q["velocity"][1, 0], atol=1e-6) and np.isclose(q["position_radius"][0, 0], -q["position_radius"][1, 0], atol=1e-6)
    assert q["position_radius"][1, 0] - q["position_radius"][0, 0] > 0.2, "overlap is pushed apart"
    t = np.float32(1.0 / 120.0)
    assert np.allclose(q["colour"][0, :3], np.float32([1, 0, 0]) * (1 - t) + np.float32([0, 0, 1]) * t, atol=1e-7), "colours mix by the timestep"
    far = padded(oracle, one(oracle, (-3.0, 5.0, 0.0)), one(oracle, (3.0, 5.0, 0.0)))
    r, _ = oracle.particles_update(osc, far, substeps=1)
    assert np.array_equal(r["colour"], far["colour"]) and r["velocity"][0, 0] == 0.0, "no interaction beyond the neighbouring cells"
    # two particles alone: the 27 neighbour cells collapse onto 2 keys and the same neighbour is merged several times — as in the reference
    two = np.concatenate([one(oracle, (-0.1, 5.0, 0.0), colour=(1, 0, 0)), one(oracle, (0.1, 5.0, 0.0), colour=(0, 0, 1))])
    w, _ = oracle.particles_update(osc, two, substeps=1)
    k = round(np.log(float(w["colour"][0, 0])) / np.log(1 - 1 / 120))
    assert 2 <= k <= 27 and np.isclose(w["colour"][0, 0], (1 - 1 / 120) ** k, rtol=1e-5)


def test_sleeping_particles(snk, oracle):
    osc = oracle.OracleScene(floor_scene(snk))
    # resting on the floor: the gravity ray hits, no impulse, the particle stays asleep and does not move
    p = one(oracle, (0.0, 0.15, 0.0), awake=0.0)
    q, rays = oracle.particles_update(osc, p, substeps=3)
    assert np.array_equal(q["position_radius"], p["position_radius"]) and np.array_equal(q["velocity"], p["velocity"])
    assert rays == 3 * 1, "an asleep particle at rest only traces the gravity ray (its offset is below the constraint threshold)"
    # asleep in mid-air: the first sub-step leaves it in place but gravity's impulse wakes it (velocity.w = 1), then it falls
    p = one(oracle, (0.0, 3.0, 0.0), awake=0.0)
    q, _ = oracle.particles_update(osc, p, substeps=1)
    assert np.array_equal(q["position_radius"], p["position_radius"]) and np.array_equal(q["velocity"][0], (0, 0, 0, 1))
    q, _ = oracle.particles_update(osc, q, substeps=1)
    assert q["position_radius"][0, 1] < 3.0 and q["velocity"][0, 1] < 0
