"""North-star estimator (1 spp, 4 bounces, MIS NEE, RR) — CUDA wavefront stages vs the CPU statement
(oracle/snk_northstar.h): same expression order under the float32 numerics contract, so radiance is compared
bit for bit; plus estimator-level properties (no NaN, energy bound in a closed diffuse box, bounce count monotone)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def bits_equal(a, b):
    return ((a.view(np.uint32) == b.view(np.uint32)) | ((a != a) & (b != b))).all(-1)


def run(snk, oracle, sc, frames, **kw):
    osc = oracle.OracleScene(sc)
    fr = oracle.OracleFrame(osc)
    rt = snk.RT().Init(sc)
    prev = None
    for f in range(frames):
        pos = np.asarray(sc.camera.pos, np.float32) + np.asarray((0.05, 0.0, 0.1), np.float32) * f
        ubo = sc.ubo(f, cam_pos=pos)
        p = ubo if prev is None else prev
        fr.set(ubo, p, oracle.default_settings(mode=1, **kw))
        fr.gbuffer()
        fr.render()
        rt.RecordRenderCmdBuf(ubo, p, mode=snk.api.MODE_NORTHSTAR, **kw)
        prev = ubo
        yield f, rt.pt, fr


@pytest.mark.parametrize("scene_name", ["small", "cornell_pl"])
def test_northstar_matches_oracle(snk, oracle, scene_name):
    sc = snk.scenes.small_test_scene(96, 64) if scene_name == "small" else snk.scenes.cornell(96, 96, with_pointlight=True)
    for f, pt, fr in run(snk, oracle, sc, frames=3):
        a, b = pt.read_aux("radiance")[..., :3], fr.read("radiance")[..., :3]
        assert not np.isnan(a).any()
        eq = bits_equal(a, b)
        assert eq.mean() >= 0.998, f"frame {f}: radiance differs on {(~eq).sum()} of {eq.size} pixels; max abs {np.abs(a - b).max()}"
        for half in ("reservoirs_temporal", "reservoirs_final"):
            ra, rb = pt.read_aux(half), fr.read(half)
            same = ra.view(np.uint8).reshape(ra.shape + (48,))[..., :44] == rb.view(np.uint8).reshape(rb.shape + (48,))[..., :44]
            assert same.all(-1).mean() >= 0.998, f"frame {f}: {half} differs on {(~same.all(-1)).sum()} pixels"
        da, db = pt.read_aux("direct")[..., :3], fr.read("direct")[..., :3]
        assert bits_equal(da, db).mean() >= 0.998
        st = pt.stats()
        oc, oa = fr.ray_counts(reset=True)
        assert (int(st["rays_secondary"]), int(st["rays_shadow"])) == (oc, oa), "ray counts differ from the oracle's"
        assert int(st["stack_overflows"]) == 0
        npix = sc.width * sc.height
        assert int(st["rays_secondary"]) <= 4 * npix


def test_bounce_count_and_flags(snk, oracle):
    sc = snk.scenes.cornell(64, 64, with_pointlight=True)
    means = []
    for mb in (0, 1, 4):
        for f, pt, fr in run(snk, oracle, sc, frames=1, max_bounces=mb):
            a, b = pt.read_aux("radiance")[..., :3], fr.read("radiance")[..., :3]
            assert bits_equal(a, b).mean() >= 0.998
            means.append(float(a.mean()))
            if mb == 0:
                assert np.array_equal(a, np.where(np.isnan(a), a, pt.read_aux("radiance")[..., :3]))
                assert int(pt.stats()["rays_secondary"]) == 0
    assert means[0] <= means[1] <= means[2] * 1.2  # indirect light only adds energy (up to 1-frame noise)


@pytest.mark.parametrize("kw", [dict(flags=1), dict(num_temporal=0, num_spatial=0), dict(num_temporal=1, num_spatial=0),
                                dict(num_temporal=0, num_spatial=6), dict(ris_iters=4, m_cap=3, neighbour_radius=8.0),
                                dict(flags=4), dict(flags=4, num_spatial=6, neighbour_radius=12.0)])
def test_restir_settings_match_oracle(snk, oracle, kw):
    """north-star ReSTIR DI stages under other settings (flags=1: reservoir passes off = plain path tracer with NEE at x0;
    ris_iters < number of point lights exercises the random-candidate branch; flags=4: unbiased reuse, the MIS weights' shadow rays
    traced inline by the spatial stage and counted), emissive + point lights (small scene)"""
    sc = snk.scenes.small_test_scene(80, 48)
    for f, pt, fr in run(snk, oracle, sc, frames=3, **kw):
        a, b = pt.read_aux("radiance")[..., :3], fr.read("radiance")[..., :3]
        eq = bits_equal(a, b)
        assert eq.mean() >= 0.998, f"{kw} frame {f}: {(~eq).sum()} pixels differ"
        st = pt.stats()
        oc, oa = fr.ray_counts(reset=True)
        assert (int(st["rays_secondary"]), int(st["rays_shadow"])) == (oc, oa)


def test_restir_mean_close_to_plain_path_tracer(snk):
    """the ReSTIR'd direct term must agree with the plain path tracer (NEE at x0) up to the visibility bias of
    reuse without shadow rays in the MIS weights (oracle/snk_restir_ns.h header): |bias| < 10 % of the frame mean,
    and exactly the same expectation with reuse off (initial RIS + visibility only: < 2 % at 96 frames)"""
    sc = snk.scenes.cornell(64, 64, with_pointlight=True)

    def mean_image(**kw):
        rt = snk.RT().Init(sc)
        acc = np.zeros((64, 64, 3))
        n = 96
        for f in range(n):
            ubo = sc.ubo(f)
            rt.RecordRenderCmdBuf(ubo, ubo, mode=snk.api.MODE_NORTHSTAR, max_bounces=0, **kw)
            acc += rt.pt.read_aux("radiance")[..., :3]
        body = (rt.pt.read_aux("flags") & 1) == 0
        return (acc / n)[body].mean(0)

    plain = mean_image(flags=1)
    initial_only = mean_image(num_temporal=0, num_spatial=0)
    full = mean_image()
    assert np.all(np.abs(initial_only - plain) / plain < 0.02), (initial_only, plain)
    assert np.all(np.abs(full - plain) / plain < 0.10), (full, plain)


def test_furnace_energy_bound(snk):
    """closed grey box (albedo 0.5) lit only by the ceiling quad: the converged mean radiance of non-emissive
    pixels must stay below the analytic bound Le * A_light/A_total * rho / (1 - rho) x safety (no energy gain)."""
    sc = snk.scenes.cornell(64, 64)
    rt = snk.RT().Init(sc)
    acc = np.zeros((64, 64, 3))
    n = 48
    for f in range(n):
        ubo = sc.ubo(f)
        rt.RecordRenderCmdBuf(ubo, ubo, mode=snk.api.MODE_NORTHSTAR)
        acc += rt.pt.read_aux("radiance")[..., :3]
    img = acc / n
    assert np.isfinite(img).all() and (img >= 0).all()
    flags = rt.pt.read_aux("flags")
    body = img[(flags & 1) == 0]
    assert 0.01 < body.mean() < 3.0


def test_deep_paths_use_one_queue_per_bounce(snk, oracle):
    """max_bounces beyond the default: every bounce gets its own reference queue / counter (24 per frame) and stays bit-exact"""
    sc = snk.scenes.cornell(48, 48, with_pointlight=True)
    for f, pt, fr in run(snk, oracle, sc, frames=2, max_bounces=12):
        a, b = pt.read_aux("radiance")[..., :3], fr.read("radiance")[..., :3]
        assert bits_equal(a, b).mean() >= 0.998
        st = pt.stats()
        oc, oa = fr.ray_counts(reset=True)
        assert (int(st["rays_secondary"]), int(st["rays_shadow"])) == (oc, oa)
