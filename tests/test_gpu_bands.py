"""Row-band sharding on the GPU path: two contexts on one device render the two halves of the frame with
halo rows (snkpt_set_halo) and swap final-reservoir halos like two ranks would; the stitched result must be
bit-identical to a whole-frame context (same kernels, so any difference is a halo / row-range bug)."""
import importlib

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_two_bands_equal_whole_frame(snk):
    sh = importlib.import_module("snake-vk_b200.sharding")
    sc = snk.scenes.small_test_scene(96, 64)
    H, W = sc.height, sc.width
    bands = sh.RowBands(H, 2, radius=6, motion_margin=4)
    st = snk.api.default_settings(neighbour_radius=6.0)
    whole = snk.PathTracer(W, H)
    whole.upload_scene(sc)
    parts = []
    for r in range(2):
        pt = snk.PathTracer(W, H, rows=bands.band(r))
        pt.upload_scene(sc)
        pt.set_halo(bands.pass1_halo, bands.gbuffer_halo)
        parts.append(pt)
    prev = None
    for f in range(4):
        pos = np.asarray(sc.camera.pos, np.float32) + np.asarray((0.02, 0.0, 0.05), np.float32) * f
        ubo = sc.ubo(f, cam_pos=pos)
        p = ubo if prev is None else prev
        for pt in [whole] + parts:
            pt.set_frame(ubo, p)
            pt.trace_primary()
            pt.render(st)
        prev = ubo
        # the halo exchange two ranks would do over NCCL, through the host here
        res = [pt.read_aux("reservoirs_final") for pt in parts]
        for r in range(2):
            for peer, send, recv in bands.exchange_plan(r):
                rows = res[peer][recv[0]:recv[1]]
                parts[r].write_reservoirs(0, rows, first=recv[0] * W)
        stitched = np.concatenate([parts[r].read_aux("radiance")[slice(*bands.band(r))] for r in range(2)])
        want = whole.read_aux("radiance")
        assert stitched.tobytes() == want.tobytes(), f"frame {f}: {(stitched != want).any(-1).sum()} pixels differ"
        fin = np.concatenate([parts[r].read_aux("reservoirs_final")[slice(*bands.band(r))] for r in range(2)])
        assert fin.tobytes() == whole.read_aux("reservoirs_final").tobytes(), f"frame {f}: final reservoirs differ"
    out = np.concatenate([parts[r].read_output("f32") for r in range(2)])
    assert out.tobytes() == whole.read_output("f32").tobytes()
    for pt in [whole] + parts:
        assert int(pt.stats()["stack_overflows"]) == 0
        pt.close()
