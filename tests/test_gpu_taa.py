"""TAA resolve (snkpt_taa_resolve, csrc/taa.cu) against the CPU statement of TAA_Resolve.comp (oracle/snk_taa.h):
bit-exact RGBA16F output over a moving-camera sequence, history / previous-velocity hand-over, reset behaviour."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_taa_sequence_matches_oracle(snk, oracle):
    sc = snk.scenes.small_test_scene(96, 64)
    rt = snk.RT().Init(sc)
    hist = pvel = None
    prev = None
    for f in range(4):
        pos = np.asarray(sc.camera.pos, np.float32) + np.asarray((0.08, 0.0, 0.15), np.float32) * f
        ubo = sc.ubo(f, cam_pos=pos)
        rt.RecordRenderCmdBuf(ubo, ubo if prev is None else prev, mode=snk.api.MODE_NORTHSTAR)
        prev = ubo
        cur = rt.pt.read_output("f16")
        vel = rt.pt.read_aux("motion")
        if hist is None:  # first resolve: history := current frame, previous velocity := velocity
            hist, pvel = cur.copy(), vel.copy()
        want = oracle.taa_resolve(cur, vel, hist, pvel)
        rt.pt.taa_resolve()
        got = rt.pt.read_output("f16")
        assert np.array_equal(got.view(np.uint16), want.view(np.uint16)), f"frame {f}: {(got.view(np.uint16) != want.view(np.uint16)).any(-1).sum()} pixels differ"
        assert np.array_equal(rt.pt.read_aux("taa_history").view(np.uint16), want.view(np.uint16))
        assert np.array_equal(rt.pt.read_aux("taa_prev_velocity").view(np.uint16), vel.view(np.uint16))
        if f:
            assert not np.array_equal(got.view(np.uint16), cur.view(np.uint16)), "the resolve must change a moving frame"
        hist, pvel = want, vel
    # after a history reset the next resolve seeds itself from the current frame again
    rt.pt.reset_history()
    ubo = sc.ubo(9)
    rt.RecordRenderCmdBuf(ubo, ubo, mode=snk.api.MODE_NORTHSTAR)
    cur, vel = rt.pt.read_output("f16"), rt.pt.read_aux("motion")
    rt.pt.taa_resolve()
    assert np.array_equal(rt.pt.read_output("f16").view(np.uint16), oracle.taa_resolve(cur, vel, cur, vel).view(np.uint16))


def test_taa_odd_size_and_band_refusal(snk, oracle):
    sc = snk.scenes.small_test_scene(45, 27)  # not multiples of 8: remainder strip keeps the rendered frame
    rt = snk.RT().Init(sc)
    ubo = sc.ubo(0)
    rt.RecordRenderCmdBuf(ubo, ubo, mode=snk.api.MODE_REFERENCE)
    cur, vel = rt.pt.read_output("f16"), rt.pt.read_aux("motion")
    rt.pt.taa_resolve()
    got = rt.pt.read_output("f16")
    assert np.array_equal(got.view(np.uint16), oracle.taa_resolve(cur, vel, cur, vel).view(np.uint16))
    assert np.array_equal(got[24:].view(np.uint16), cur[24:].view(np.uint16))
    band = snk.PathTracer(64, 64, rows=(16, 48))
    with pytest.raises(RuntimeError):
        band.taa_resolve()
    band.close()
