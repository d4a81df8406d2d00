"""GPU parity: BVH build + traversal + watertight intersection against the CPU oracle.

Everything goes through the C ABI (snkpt_trace_rays / snkpt_trace_primary). Bar: bit-exact
(t, u, v, prim, inst) — the triangle test is written op-for-op like the oracle's, and closest-hit
with the (inst, prim) tie-break is traversal-order independent."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def random_rays(rng, n, lo, hi, scenes):
    rays = np.zeros(n, scenes.RAY_DTYPE)
    rays["o"] = rng.uniform(lo, hi, (n, 3)).astype(np.float32)
    d = rng.normal(size=(n, 3))
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    rays["d"] = d.astype(np.float32)
    rays["tmin"] = 0.001
    rays["tmax"] = 1000.0
    return rays


def hits_equal(a, b):
    miss = (a["t"] < 0) & (b["t"] < 0)
    same = miss | ((a["t"] == b["t"]) & (a["u"] == b["u"]) & (a["v"] == b["v"]) & (a["prim"] == b["prim"]) & (a["inst"] == b["inst"]))
    return same


@pytest.mark.parametrize("scene_name", ["cornell", "small"])
def test_random_rays_match_bruteforce(snk, oracle, scene_name):
    sc = snk.scenes.cornell(64, 64) if scene_name == "cornell" else snk.scenes.small_test_scene()
    osc = oracle.OracleScene(sc)
    pt = snk.PathTracer(sc.width, sc.height)
    pt.upload_scene(sc)
    rng = np.random.default_rng(1)
    lo, hi = ((0, 0, -2), (5.5, 5.5, 5.5)) if scene_name == "cornell" else ((-8, 0, -8), (8, 6, 8))
    rays = random_rays(rng, 20000, lo, hi, snk.scenes)
    ref = osc.trace(rays, brute=True)
    got = pt.trace_rays(rays)
    same = hits_equal(got, ref)
    assert same.all(), f"{(~same).sum()} of {same.size} closest hits differ; first: {got[~same][:3]} vs {ref[~same][:3]}"
    ref_any = osc.trace(rays, any_hit=True, brute=True)
    got_any = pt.trace_rays(rays, any_hit=True)
    assert np.array_equal(got_any["t"] >= 0, ref_any["t"] >= 0)
    pt.close()


def test_edge_and_vertex_rays_are_watertight(snk, oracle):
    """rays aimed exactly at shared edges / vertices of the Cornell quads must hit (no cracks) and resolve
    to the same primitive as the oracle (lowest (inst, prim) at equal t)."""
    sc = snk.scenes.cornell(64, 64)
    osc = oracle.OracleScene(sc)
    pt = snk.PathTracer(64, 64)
    pt.upload_scene(sc)
    rng = np.random.default_rng(2)
    n = 4096
    rays = np.zeros(n, snk.scenes.RAY_DTYPE)
    o = np.tile(np.array([2.75, 2.75, -8.0], np.float32), (n, 1))
    # points on the diagonal of the back wall (shared edge of its two triangles) and on wall/floor seams
    s = rng.uniform(0.0, 5.5, n).astype(np.float32)
    target = np.stack([s, s, np.full(n, 5.5, np.float32)], 1)
    target[n // 2:, 1] = 0.0  # seam between back wall and floor
    d = target - o
    rays["o"], rays["d"], rays["tmin"], rays["tmax"] = o, d.astype(np.float32), 0.0, 10.0
    ref = osc.trace(rays, brute=True)
    got = pt.trace_rays(rays)
    assert (ref["t"] >= 0).all()
    same = hits_equal(got, ref)
    assert same.all(), f"{(~same).sum()} edge rays differ"
    pt.close()


def test_empty_and_degenerate_inputs(snk):
    sc = snk.scenes.cornell(32, 32)
    pt = snk.PathTracer(32, 32)
    # tracing before any TLAS exists is a state error, not a crash
    with pytest.raises(snk.SnkError):
        pt.trace_rays(np.zeros(4, snk.scenes.RAY_DTYPE))
    pt.upload_scene(sc)
    assert pt.trace_rays(np.zeros(0, snk.scenes.RAY_DTYPE)).size == 0
    rays = np.zeros(3, snk.scenes.RAY_DTYPE)
    rays["o"] = (2.75, 2.75, -8)
    rays["d"] = [(0, 0, 1), (np.nan, 0, 1), (0, 0, 0)]
    rays["tmin"], rays["tmax"] = 0.0, 100.0
    h = pt.trace_rays(rays)
    assert h["t"][0] > 0 and h["t"][1] < 0 and h["t"][2] < 0
    # tmax before the wall → miss; tmin beyond it → miss
    rays["d"] = (0, 0, 1)
    rays["tmax"] = [5.0, 10.5, 100.0]
    rays["tmin"] = [0.0, 0.0, 13.6]
    h = pt.trace_rays(rays)
    assert (h["t"] < 0).all()
    pt.close()


def test_gbuffer_matches_oracle_cornell(snk, oracle):
    sc = snk.scenes.cornell(128, 128)
    osc = oracle.OracleScene(sc)
    fr = oracle.OracleFrame(osc)
    ubo = sc.ubo(0)
    fr.set(ubo, ubo)
    fr.gbuffer()
    pt = snk.PathTracer(sc.width, sc.height)
    pt.upload_scene(sc)
    pt.set_frame(ubo, ubo)
    pt.trace_primary()
    for name in ["prim", "inst", "hit_t", "bary", "depth", "normal", "albedo", "rma", "motion", "flags"]:
        a, b = pt.read_aux(name), fr.read(name)
        assert np.array_equal(a.view(np.uint8), b.view(np.uint8)), f"G-buffer plane {name}: {(a != b).sum()} texels differ"
    pt.close()


@pytest.mark.parametrize("variant", ["instanced", "flat"])
def test_primary_ids_c2(snk, oracle, variant):
    """C2 (261,122 triangles) at 480x270: primary-hit primitive ids agree on >= 99.99 % of pixels."""
    sc = snk.scenes.c2(variant, width=480, height=270)
    osc = oracle.OracleScene(sc)
    fr = oracle.OracleFrame(osc)
    ubo = sc.ubo(0)
    fr.set(ubo, ubo)
    fr.gbuffer()
    pt = snk.PathTracer(sc.width, sc.height)
    pt.upload_scene(sc)
    pt.set_frame(ubo, ubo)
    pt.trace_primary()
    a, b = pt.read_aux("prim"), fr.read("prim")
    ai, bi = pt.read_aux("inst"), fr.read("inst")
    agree = ((a == b) & (ai == bi)).mean()
    assert agree >= 0.9999, f"prim-id agreement {agree:.6f}"
    assert (b != 0xFFFFFFFF).mean() > 0.5
    t = pt.blas_table()
    assert t["tri_count"].sum() == sum(m.num_tris for m in sc.meshes)
    pt.close()


def test_refit_matches_rebuild_and_bruteforce(snk, oracle):
    """animated vertices: snkpt_update_positions + snkpt_refit_blas (topology kept, boxes re-quantised bottom-up) must
    give the same hits as the brute-force oracle on the deformed mesh, bit for bit, and as a full rebuild."""
    import ctypes as C
    S = snk.scenes
    rng = np.random.default_rng(9)
    mesh = S.c5_soup(20000)
    ident = np.eye(4, dtype=np.float32)
    mats = S.make_materials([((0.8, 0.8, 0.8), 0.5, 0.0, 0)])

    def scene_of(m):
        return S.Scene("refit", 32, 32, [m], [S.Entity(0, ident, [0]), S.Entity(0, S.trs((3, 1, 2), (10, 40, 70), (1.2, 0.8, 1.0)), [0])], mats,
                       S.Camera((0, 0, -30), (0, 0, 1)))

    pt = snk.PathTracer(32, 32)
    flat = pt.upload_scene(scene_of(mesh))
    entry = np.zeros((), snk.api.MESH_ENTRY_DTYPE)  # first mesh: offsets 0, BLAS 0
    entry["num_submeshes"] = 1
    rays = random_rays(rng, 30000, (-14, -14, -14), (14, 14, 14), S)
    for step in range(3):
        pos = mesh.positions.copy()
        pos[:, 1] += (0.3 * (step + 1)) * np.sin(pos[:, 0] * 2.0 + step).astype(np.float32)
        pos[:, 0] *= np.float32(1.0 + 0.1 * step)
        moved = S.Mesh(pos, mesh.normals, mesh.uvs, mesh.tangents, mesh.indices, mesh.submeshes)
        pt._ck(pt.L.snkpt_update_positions(pt.h, entry.ctypes.data_as(C.c_void_p), pos.ctypes.data_as(C.c_void_p), pos.shape[0]), "update_positions")
        pt._ck(pt.L.snkpt_refit_blas(pt.h, 0, None), "refit_blas")
        pt.build_tlas()
        got = pt.trace_rays(rays)
        ref = oracle.OracleScene(scene_of(moved)).trace(rays, threads=16)
        same = hits_equal(got, ref)
        assert same.all(), f"step {step}: {(~same).sum()} of {same.size} hits differ after refit"
        assert (ref["t"] >= 0).mean() > 0.03
        lo, hi = pt.blas_table()[0]["bounds_lo"], pt.blas_table()[0]["bounds_hi"]
        assert (lo <= pos.min(0)).all() and (hi >= pos.max(0)).all()
    pt._ck(pt.L.snkpt_rebuild_blas(pt.h, 0, None), "rebuild_blas")
    pt.build_tlas()
    assert hits_equal(pt.trace_rays(rays), got).all()
    assert int(pt.stats()["stack_overflows"]) == 0
    pt.close()


def _soup(rng, n, spread, size):
    """n random triangles: centres in [-spread, spread]^3, edge vectors of length ~size"""
    c = rng.uniform(-spread, spread, (n, 1, 3))
    p = (c + rng.normal(scale=size, size=(n, 3, 3))).reshape(-1, 3).astype(np.float32)
    nrm = np.tile(np.float32([[0, 1, 0]]), (3 * n, 1))
    return p, nrm, np.arange(3 * n, dtype=np.uint32)


@pytest.mark.parametrize("case", ["n1", "n2", "n3", "n9", "n17", "n4096", "n4097", "identical", "identical3000", "coplanar_grid", "mixed_scale", "degenerate"])
def test_builder_edge_cases_match_bruteforce(snk, oracle, case):
    """BLAS sizes around the branches of the builder (single leaf, fewer primitives than the 8 slots of a wide node, the
    one-CTA path's 4096-primitive limit and the first size of the general path), Morton-code duplicates (identical and
    coincident primitives), primitives of very different size and zero-area triangles: every closest hit and every any-hit
    answer equals the brute-force oracle's, bit for bit, for an identity and a rotated, non-uniformly scaled instance."""
    S = snk.scenes
    import zlib
    rng = np.random.default_rng(zlib.crc32(case.encode()))
    if case.startswith("n"):
        p, nrm, idx = _soup(rng, int(case[1:]), 4.0, 0.6)
    elif case in ("identical", "identical3000"):  # 300 / 3000 copies of the same triangle + a few others: equal Morton keys, equal boxes
        p1, _, _ = _soup(rng, 1, 1.0, 1.5)                # (3000 copies: a hierarchy deeper than the packet traversal's shared stack)
        p2, _, _ = _soup(rng, 20, 4.0, 0.6)
        p = np.concatenate([np.tile(p1, (3000 if case == "identical3000" else 300, 1)), p2]).astype(np.float32)
        nrm = np.tile(np.float32([[0, 1, 0]]), (p.shape[0], 1)); idx = np.arange(p.shape[0], dtype=np.uint32)
    elif case == "coplanar_grid":  # 32 x 32 quads in the plane y = 0.5: one Morton axis carries no information
        g = np.arange(33, dtype=np.float32) * 0.25 - 4.0
        xx, zz = np.meshgrid(g, g)
        p = np.stack([xx.ravel(), np.full(xx.size, 0.5, np.float32), zz.ravel()], -1).astype(np.float32)
        q = (np.arange(32)[:, None] * 33 + np.arange(32)[None, :]).ravel()
        idx = np.stack([q, q + 1, q + 34, q, q + 34, q + 33], -1).astype(np.uint32).ravel()
        nrm = np.tile(np.float32([[0, 1, 0]]), (p.shape[0], 1))
    elif case == "mixed_scale":  # two huge triangles over 3000 tiny ones
        pa, _, _ = _soup(rng, 3000, 4.0, 0.02)
        pb = np.float32([[-6, -1, -6], [6, -1, -6], [0, -1, 7], [-6, 6, 0], [6, 6, 0], [0, -6, 0.5]])
        p = np.concatenate([pa, pb]).astype(np.float32)
        nrm = np.tile(np.float32([[0, 1, 0]]), (p.shape[0], 1)); idx = np.arange(p.shape[0], dtype=np.uint32)
    else:  # degenerate: zero-area (collinear / point) triangles between ordinary ones
        p, nrm, idx = _soup(rng, 64, 4.0, 0.6)
        p = p.copy()
        p[3:6] = p[3]                      # a point
        p[9:12] = [p[9], p[9] + 1, p[9] + 2]  # collinear
    mesh = S.Mesh.single(p, nrm, idx)
    ents = [S.Entity(0, np.eye(4, dtype=np.float32), [0]), S.Entity(0, S.trs((1.5, -0.5, 2), (25, 50, 75), (1.3, 0.7, 1.1)), [0])]
    sc = S.Scene(case, 32, 32, [mesh], ents, S.make_materials([((0.8, 0.8, 0.8), 0.5, 0.0, 0)]), S.Camera((0, 0, -20), (0, 0, 1)))
    pt = snk.PathTracer(32, 32)
    pt.upload_scene(sc)
    rays = random_rays(rng, 6000, (-7, -7, -7), (7, 7, 7), S)
    # half of the rays aim at a vertex of the mesh so that small inputs are hit at all
    tgt = p[rng.integers(0, p.shape[0], 3000)] + rng.normal(scale=0.05, size=(3000, 3)).astype(np.float32)
    d = tgt - rays["o"][:3000]
    rays["d"][:3000] = (d / np.linalg.norm(d, axis=1, keepdims=True)).astype(np.float32)
    osc = oracle.OracleScene(sc)
    ref = osc.trace(rays, brute=True)
    got = pt.trace_rays(rays)
    same = hits_equal(got, ref)
    assert same.all(), f"{case}: {(~same).sum()} of {same.size} closest hits differ; first: {got[~same][:3]} vs {ref[~same][:3]}"
    assert (ref["t"] >= 0).any()
    got_any = pt.trace_rays(rays, any_hit=True)
    assert np.array_equal(got_any["t"] >= 0, ref["t"] >= 0)
    tbl = pt.blas_table()[0]
    assert (tbl["bounds_lo"] <= p.min(0)).all() and (tbl["bounds_hi"] >= p.max(0)).all()
    assert int(pt.stats()["stack_overflows"]) == 0
    # the same hierarchies through the PACKET traversal of the camera rays (gbuffer.cu: one stack per 8x4 tile): ids equal the oracle's
    ubo = sc.ubo(0)
    pt.set_frame(ubo, ubo)
    pt.trace_primary()
    fr = oracle.OracleFrame(osc)
    fr.set(ubo, ubo, oracle.default_settings())
    fr.gbuffer()
    for name in ("prim", "inst"):
        assert np.array_equal(pt.read_aux(name), fr.read(name)), f"{case}: primary {name} ids differ"
    assert np.array_equal(pt.read_aux("hit_t").view(np.uint32), fr.read("hit_t").view(np.uint32))
    assert int(pt.stats()["stack_overflows"]) == 0
    pt.close()
