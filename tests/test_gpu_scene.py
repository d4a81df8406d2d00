"""Scene front-end end to end (include/snkscene.h → snkscene_push → libsnkpt): an entity hierarchy with a parent that rotates,
a moving camera, point lights parented to a mover and an emissive cube, rendered for three frames; the CUDA frames must equal
the CPU oracle fed with the same flat buffers bit for bit (so everything the front-end produced reached the kernels intact),
and the motion plane must see the animated parent."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


class FlatScene:
    """adapter: the front-end's buffers + the mesh pools, in the shape OracleScene expects from scenes.Scene"""

    def __init__(self, base, flat):
        self.width, self.height, self.textures = base.width, base.height, []
        pools = base.flatten()
        self._flat = dict(pools)
        for k in ("transforms", "prev_transforms", "instances", "blas_of_instance", "emissive_idx", "materials", "lights"):
            self._flat[k] = flat[k]

    def flatten(self):
        return self._flat


def bits_equal(a, b):
    return ((a.view(np.uint32) == b.view(np.uint32)) | ((a != a) & (b != b))).all(-1)


@pytest.mark.parametrize("mode", [0, 1])
def test_frontend_frames_match_oracle(snk, oracle, mode):
    S = snk.scenes
    W, H = 96, 64
    base = S.small_test_scene(W, H)            # only its meshes / materials are used
    pt = snk.PathTracer(W, H)
    fe = snk.frontend.SceneFrontEnd()
    meshes = [fe.register_mesh(pt.upload_mesh(m), m.submeshes) for m in base.meshes]
    for i in range(base.materials.shape[0]):
        fe.set_material(i, base.materials[i:i + 1])
    emissive_slot = int(np.nonzero(base.materials["flags"] & S.MAT_FLAG_EMISSIVE)[0][0])
    floor = fe.create_entity()
    fe.set_static_mesh(floor, meshes[0], [0])
    hub = fe.create_entity(pos=(0.0, 2.0, 1.0), scale=(1.5, 1.0, 1.0))
    arms = []
    for k in range(4):
        a = fe.create_entity(parent=hub, pos=(2.0 * np.cos(k * 1.57), 0.3 * k, 2.0 * np.sin(k * 1.57)), euler=(10.0 * k, 25.0, 0.0), scale=(0.6,) * 3)
        fe.set_static_mesh(a, meshes[1], [1 + k % 2])
        arms.append(a)
    cube = fe.create_entity(parent=arms[0], pos=(0.0, 1.8, 0.0), scale=(0.2,) * 3)
    fe.set_static_mesh(cube, meshes[2], [emissive_slot])
    lamp = fe.create_entity(parent=hub, pos=(0.0, 2.5, -1.0))
    fe.set_pointlight(lamp, (8.0, 7.0, 6.0))
    lamp2 = fe.create_entity(pos=(-3.0, 3.0, -2.0))
    fe.set_pointlight(lamp2, (4.0, 6.0, 9.0))
    fe.set_directional_light((1.0, 1.0, 1.0), 0.8, 0.9)
    cam = fe.create_entity(pos=(0.0, 2.5, -9.0), euler=(0.0, 180.0, 0.0))
    fe.set_camera(cam, 90.0, W / H)
    settings = snk.api.default_settings(mode=mode)
    fr = None
    for f in range(3):
        fe.set_transform(hub, euler=(0.0, 12.0 * f, 0.0))            # the hub turns: children, lamp and cube follow
        fe.set_transform(cam, pos=(0.05 * f, 2.5, -9.0 + 0.1 * f))
        fe.frame_start(1000.0 * f / 60.0)
        flat = fe.flat()
        fe.push(pt)
        pt.trace_primary()
        pt.render(settings)
        if fr is None:
            osc = oracle.OracleScene(FlatScene(base, flat))
            fr = oracle.OracleFrame(osc, W, H)
        else:  # the oracle scene is immutable: rebuild it from this frame's buffers, keep the frame state (reservoirs, G-buffer history)
            osc2 = oracle.OracleScene(FlatScene(base, flat))
            oracle.lib().snkor_frame_set_scene(fr.h, osc2.h)
            fr.osc, osc = osc2, osc2
        fr.set(flat["ubo"], flat["prev_ubo"], oracle.default_settings(mode=mode))
        fr.gbuffer()
        fr.render()
        fe.frame_end()
        assert np.array_equal(pt.read_aux("prim"), fr.read("prim")), f"frame {f}: primary ids"
        a, b = pt.read_aux("radiance")[..., :3], fr.read("radiance")[..., :3]
        assert bits_equal(a, b).mean() >= 0.998, f"mode {mode} frame {f}: {(~bits_equal(a, b)).sum()} pixels differ"
        if f:
            assert np.array_equal(pt.read_aux("motion"), fr.read("motion"))
            moving = pt.read_aux("inst") >= 1
            assert np.abs(pt.read_aux("motion").astype(np.float32))[moving].max() > 1e-3, "the rotating hub must show up in the motion vectors"
    pt.close()
    fe.close()


def test_cpp_host_program_renders():
    """the C++ program over host/snk_rt.hpp + snkscene.h (no Python in the loop): two frames, image read back, floor lit"""
    import os, subprocess
    exe = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "snake-vk_b200", "example_rt")
    assert os.path.exists(exe), "build it with make -C snake-vk_b200"
    r = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stderr
    lit = int(r.stdout.split()[1])
    assert lit > 64 * 48 // 4, r.stdout
