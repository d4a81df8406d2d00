"""Row-band groups (include/snkpt.h "Row-band GROUPS"): N contexts of one process on ONE device, each driven by its own host
thread, render the N bands of the frame and exchange reservoir rows inside snkpt_render (temporal reservoirs mid frame, final
reservoirs at the end). The stitched frame must equal a whole-frame context bit for bit — radiance, both reservoir halves
on the band rows, the fp16 output — for several frames with camera motion. The NCCL back end of the same exchange is
exercised by bench.py --gpus N (one process per GPU), which asserts the same equality on its first frames."""
import threading

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def render_bands(snk, parts, ubo, prev, settings):
    errs = []

    def work(pt):
        try:
            pt.set_frame(ubo, prev)
            pt.trace_primary()
            pt.render(settings)
            pt.synchronize()
        except Exception as e:  # noqa: BLE001
            errs.append(e)

    ts = [threading.Thread(target=work, args=(pt,)) for pt in parts]
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    assert not errs, errs


@pytest.mark.parametrize("world,radius,margin,starts", [(2, 6, 4, None), (3, 6, 4, None), (4, 30, 16, None),
                                                        (3, 6, 4, (0, 10, 40, 64)),            # unequal bands (snkpt_group_set_bands)
                                                        (4, 30, 16, (0, 100, 104, 200, 256))])  # a band thinner than the reuse radius
def test_local_group_equals_whole_frame(snk, world, radius, margin, starts):
    sc = snk.scenes.small_test_scene(96, 64) if world < 4 else snk.scenes.small_test_scene(128, 256)
    H, W = sc.height, sc.width
    band = (lambda r: (starts[r], starts[r + 1])) if starts else (lambda r: snk.api.group_band(H, world, r))
    st = snk.api.default_settings(mode=snk.api.MODE_NORTHSTAR, neighbour_radius=float(radius))
    whole = snk.PathTracer(W, H)
    whole.upload_scene(sc)
    parts = []
    for r in range(world):
        pt = snk.PathTracer(W, H, rows=band(r))
        pt.upload_scene(sc)
        if starts:
            pt.group_set_bands(starts)
        parts.append(pt)
    snk.api.group_attach_local(parts, radius, margin)
    assert parts[1].group_info()[:2] == (1, world)
    prev = None
    for f in range(4):
        pos = np.asarray(sc.camera.pos, np.float32) + np.asarray((0.02, 0.0, 0.05), np.float32) * f
        ubo = sc.ubo(f, cam_pos=pos)
        p = ubo if prev is None else prev
        prev = ubo
        whole.set_frame(ubo, p)
        whole.trace_primary()
        whole.render(st)
        render_bands(snk, parts, ubo, p, st)
        for name in ("radiance", "direct", "reservoirs_final", "reservoirs_temporal"):
            got = np.concatenate([parts[r].read_aux(name)[slice(*band(r))] for r in range(world)])
            want = whole.read_aux(name)
            assert got.tobytes() == want.tobytes(), f"world {world} frame {f}: {name} differs on {(got != want).reshape(H, W, -1).any(-1).sum()} pixels"
    out = np.concatenate([parts[r].read_output("f32") for r in range(world)])
    assert out.tobytes() == whole.read_output("f32").tobytes()
    for pt in [whole] + parts:
        assert int(pt.stats()["stack_overflows"]) == 0
    assert parts[0].group_info()[2] == 0  # the local back end copies, it does not send
    for pt in parts:
        pt.group_detach()
    for pt in [whole] + parts:
        pt.close()


def test_group_rejects_wrong_rows_and_reference_mode(snk):
    sc = snk.scenes.small_test_scene(64, 48)
    a = snk.PathTracer(64, 48, rows=(0, 20))
    b = snk.PathTracer(64, 48, rows=(20, 48))
    with pytest.raises(snk.api.SnkError):
        snk.api.group_attach_local([a, b], 6, 4)   # not the equal bands of snkpt_group_band
    with pytest.raises(snk.api.SnkError):
        a.group_set_bands((0, 20, 20, 48))          # an empty band
    with pytest.raises(snk.api.SnkError):
        a.group_set_bands((0, 20, 40))              # does not end at the image height
    a.group_set_bands((0, 20, 48)); b.group_set_bands((0, 20, 48))
    snk.api.group_attach_local([a, b], 6, 4)       # the same rows are a group once the table says so
    a.group_detach(); b.group_detach()
    a.close(); b.close()
    parts = [snk.PathTracer(64, 48, rows=snk.api.group_band(48, 2, r)) for r in range(2)]
    for pt in parts:
        pt.upload_scene(sc)
    snk.api.group_attach_local(parts, 6, 4)
    ubo = sc.ubo(0)
    parts[0].set_frame(ubo, ubo)
    parts[0].trace_primary()
    with pytest.raises(snk.api.SnkError):
        parts[0].render(snk.api.default_settings(mode=snk.api.MODE_REFERENCE))
    with pytest.raises(snk.api.SnkError):
        parts[0].render(snk.api.default_settings(mode=snk.api.MODE_NORTHSTAR, neighbour_radius=30.0))  # wider than the group exchanges
    with pytest.raises(snk.api.SnkError):
        parts[0].set_halo(3, 5)
    for pt in parts:
        pt.close()
