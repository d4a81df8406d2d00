"""Albedo textures (snkpt_set_texture; RayClosestHit.rchit:100-116, GBuffer.frag:56-58): G-buffer albedo plane byte-exact and
frames bit-exact against the CPU oracle on a textured scene, in both estimator modes; slot release restores the plain albedo."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def bits_equal(a, b):
    return ((a.view(np.uint32) == b.view(np.uint32)) | ((a != a) & (b != b))).all(-1)


@pytest.mark.parametrize("mode", [0, 1])
def test_textured_frames_match_oracle(snk, oracle, mode):
    sc = snk.scenes.textured_test_scene(96, 64)
    osc = oracle.OracleScene(sc)
    fr = oracle.OracleFrame(osc)
    rt = snk.RT().Init(sc)
    prev = None
    for f in range(2):
        pos = np.asarray(sc.camera.pos, np.float32) + np.asarray((0.05, 0.0, 0.1), np.float32) * f
        ubo = sc.ubo(f, cam_pos=pos)
        p = ubo if prev is None else prev
        fr.set(ubo, p, oracle.default_settings(mode=mode))
        fr.gbuffer()
        fr.render()
        rt.RecordRenderCmdBuf(ubo, p, mode=mode)
        prev = ubo
        assert np.array_equal(rt.pt.read_aux("albedo"), fr.read("albedo")), f"frame {f}: G-buffer albedo differs"
        na, nb = rt.pt.read_aux("normal").view(np.uint16), fr.read("normal").view(np.uint16)
        assert np.array_equal(na, nb), f"frame {f}: G-buffer normal (normal map through the interpolated TBN) differs on {(na != nb).any(-1).sum()} pixels"
        a, b = rt.pt.read_aux("radiance")[..., :3], fr.read("radiance")[..., :3]
        eq = bits_equal(a, b)
        assert eq.mean() >= 0.998, f"mode {mode} frame {f}: radiance differs on {(~eq).sum()} pixels, max abs {np.abs(a - b).max()}"


def test_normal_map_changes_the_gbuffer_normal(snk):
    sc = snk.scenes.textured_test_scene(64, 48)
    rt = snk.RT().Init(sc)
    ubo = sc.ubo(0)
    rt.RecordRenderCmdBuf(ubo, ubo, mode=0)
    mapped = rt.pt.read_aux("normal").astype(np.float32)[..., :3]
    rt.pt.set_texture(2, None)  # drop the normal map
    rt.RecordRenderCmdBuf(ubo, ubo, mode=0)
    plain = rt.pt.read_aux("normal").astype(np.float32)[..., :3]
    hit = rt.pt.read_aux("prim") != 0xFFFFFFFF
    assert not np.array_equal(mapped, plain)
    assert np.allclose(np.linalg.norm(mapped[hit], axis=-1), 1.0, atol=2e-3), "mapped normals are unit vectors"
    cosang = (mapped[hit] * plain[hit]).sum(-1)
    assert cosang.min() > 0.5 and (cosang < 0.999).mean() > 0.2, "the map tilts the normal without flipping it"


def test_texture_slot_release(snk):
    sc = snk.scenes.textured_test_scene(64, 48)
    rt = snk.RT().Init(sc)
    ubo = sc.ubo(0)
    rt.RecordRenderCmdBuf(ubo, ubo, mode=0)
    with_tex = rt.pt.read_aux("albedo").copy()
    for i in range(len(sc.textures)):
        rt.pt.set_texture(i, None)
    rt.RecordRenderCmdBuf(ubo, ubo, mode=0)
    without = rt.pt.read_aux("albedo")
    plain = snk.scenes.textured_test_scene(64, 48)
    plain.textures = []
    rt2 = snk.RT().Init(plain)
    rt2.RecordRenderCmdBuf(ubo, ubo, mode=0)
    assert np.array_equal(without, rt2.pt.read_aux("albedo"))
    assert not np.array_equal(with_tex, without)
    with pytest.raises(RuntimeError):
        rt.pt.set_texture(5000, np.zeros((2, 2, 4), np.uint8))
