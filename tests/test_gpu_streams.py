"""Stream ordering of the C ABI: the host may run a frame ahead of the GPU, so entry points that rewrite data a
frame in flight still reads (vertex positions, BVH nodes / triangles, reservoirs, textures) must order themselves after
the stream the frames run on, and queries on another stream after the uploads of the previous one. A context driven
WITHOUT any host synchronisation between frames, mutations and queries on two caller streams must produce the images of a
context that synchronises after every call."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def animated_positions(mesh, step):
    pos = mesh.positions.copy()
    pos[:, 1] += np.float32(0.08 * step) * np.sin(pos[:, 0] * 3.0 + step).astype(np.float32)
    return pos


@pytest.mark.parametrize("mode", ["reference", "northstar"])
def test_mutations_between_frames_without_host_sync(snk, mode):
    torch = pytest.importorskip("torch")
    sc = snk.scenes.small_test_scene(96, 64)
    s_frame, s_other = torch.cuda.Stream(), torch.cuda.Stream()
    gpu_mode = snk.api.MODE_NORTHSTAR if mode == "northstar" else snk.api.MODE_REFERENCE
    st = snk.api.default_settings(mode=gpu_mode)
    mesh_id = 1  # the sphere mesh of the test scene
    mesh = sc.meshes[mesh_id]

    def run(sync_everything):
        pt = snk.PathTracer(sc.width, sc.height)
        flat = pt.upload_scene(sc)
        i0, v0, b0 = flat["mesh_entries"][mesh_id]
        entry = np.zeros((), snk.api.MESH_ENTRY_DTYPE)
        entry["data_start_indices_idx"], entry["data_start_vertex_idx"], entry["num_submeshes"] = i0, v0, 1
        entry["blas_id"][0] = b0
        imgs, hits = [], []
        prev = None
        rays = np.zeros(4096, snk.scenes.RAY_DTYPE)
        rng = np.random.default_rng(3)
        rays["o"] = rng.uniform((-4, 0.5, -6), (4, 3, -2), (4096, 3))
        d = rng.normal(size=(4096, 3)); rays["d"] = d / np.linalg.norm(d, axis=1, keepdims=True)
        rays["tmin"], rays["tmax"] = 1e-3, 100.0
        for f in range(5):
            ubo = sc.ubo(f, cam_pos=np.asarray(sc.camera.pos, np.float32) + np.float32(0.05 * f))
            p = ubo if prev is None else prev
            prev = ubo
            pt.build_tlas(s_frame.cuda_stream)
            pt.set_frame(ubo, p)
            pt.trace_primary(stream=s_frame.cuda_stream)
            pt.render(st, s_frame.cuda_stream)
            if sync_everything:
                pt.synchronize(); torch.cuda.synchronize()
            # mutate the sphere mesh right behind the frame (no host synchronisation in the asynchronous run)
            pos = animated_positions(mesh, f + 1)
            pt._ck(pt.L.snkpt_update_positions(pt.h, snk.api._p(entry), snk.api._p(pos), pos.shape[0]), "snkpt_update_positions")
            pt._ck(pt.L.snkpt_refit_blas(pt.h, int(b0), s_other.cuda_stream), "snkpt_refit_blas")
            if sync_everything:
                pt.synchronize(); torch.cuda.synchronize()
            # a query on ANOTHER stream than the one the TLAS build of the next frame will use
            pt.build_tlas(s_frame.cuda_stream)
            h = np.zeros(rays.shape[0], snk.scenes.HIT_DTYPE)
            pt._ck(pt.L.snkpt_trace_rays(pt.h, snk.api._p(rays), rays.shape[0], 0, snk.api._p(h), s_other.cuda_stream), "snkpt_trace_rays")
            hits.append(h)
            imgs.append(None)  # read at the end: reading would synchronise
            if sync_everything:
                pt.synchronize(); torch.cuda.synchronize()
        torch.cuda.synchronize()
        out = pt.read_aux("radiance").copy()
        res = pt.read_aux("reservoirs_final").copy()
        assert int(pt.stats()["stack_overflows"]) == 0
        pt.close()
        return out, res, hits

    a_img, a_res, a_hits = run(False)
    b_img, b_res, b_hits = run(True)
    assert a_img.tobytes() == b_img.tobytes(), f"{(a_img != b_img).any(-1).sum()} pixels differ between the asynchronous and the synchronised run"
    assert a_res.tobytes() == b_res.tobytes()
    for f, (x, y) in enumerate(zip(a_hits, b_hits)):
        assert x.tobytes() == y.tobytes(), f"query after frame {f}: {(x != y).sum()} hits differ"


@pytest.mark.parametrize("mode", ["reference", "northstar"])
def test_frames_running_ahead_match_synchronised_run(snk, mode):
    """The TLAS rebuild and the camera rays of frame f+1 run ahead of frame f on an internal stream (second copies of the TLAS storage
    and of the G-buffer planes) when nothing they read changed since the previous frame; a scene change puts them back in order for
    one frame. Nine frames enqueued without any host synchronisation — unchanged scene, a moved instance before frames 4 and 5,
    unchanged again — must leave the same accumulated image, radiance, reservoirs and previous-frame planes as a run that
    synchronises after every call (where running ahead cannot overlap anything)."""
    torch = pytest.importorskip("torch")
    sc = snk.scenes.small_test_scene(640, 360)  # large enough for the GPU to fall behind the host: frames really are in flight
    s_frame = torch.cuda.Stream()
    gpu_mode = snk.api.MODE_NORTHSTAR if mode == "northstar" else snk.api.MODE_REFERENCE
    st = snk.api.default_settings(mode=gpu_mode)

    def run(sync_everything):
        pt = snk.PathTracer(sc.width, sc.height)
        flat = pt.upload_scene(sc)
        tr = flat["transforms"].copy()
        prev = None
        for f in range(9):
            if f in (4, 5):  # move the last instance: the rebuild and the camera rays of this frame must see the new transform
                old = tr.copy()
                tr[-1, 13] += np.float32(0.15)
                pt.set_transforms(tr, old)
            ubo = sc.ubo(f, cam_pos=np.asarray(sc.camera.pos, np.float32) + np.float32(0.04 * f))
            p = ubo if prev is None else prev
            prev = ubo
            pt.build_tlas(s_frame.cuda_stream)
            pt.set_frame(ubo, p)
            pt.trace_primary(stream=s_frame.cuda_stream)
            pt.render(st, s_frame.cuda_stream)
            pt.accumulate(reset=(f == 0), stream=s_frame.cuda_stream)
            if sync_everything:
                pt.synchronize(); torch.cuda.synchronize()
        torch.cuda.synchronize()
        out = [pt.read_aux(k).copy() for k in ("accum", "radiance", "reservoirs_final", "reservoirs_temporal", "prev_normal", "prev_depth", "normal", "prim")]
        assert int(pt.stats()["stack_overflows"]) == 0
        pt.close()
        return out

    a, b = run(False), run(True)
    for name, x, y in zip(("accum", "radiance", "reservoirs_final", "reservoirs_temporal", "prev_normal", "prev_depth", "normal", "prim"), a, b):
        assert x.tobytes() == y.tobytes(), f"{mode}: plane '{name}' differs between the run-ahead and the synchronised run"
