"""N3 on the GPU: the meshes the engine ships (Core/res/meshes/cube.glb, sphere.glb, read by
snake-vk_b200/assets.py:load_glb as AssetLoader.cpp:337-466 reads them) and the scene file of the editor's project template
(my_scene.json) go through assets -> scene front-end -> C ABI and are rendered in both modes; the frames must equal the CPU
oracle fed with the same buffers bit for bit. The GPU box has no /root/reference: the mesh arrays are the committed fixture
tests/golden/shipped_mesh_arrays.npz (tools/make_golden.py:shipped_mesh_arrays; test_cpu_assets.py re-derives it from the .glb
files where the reference is mounted) and the scene facts tests/golden/shipped_scene.json."""
import importlib
import json
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def shipped_mesh(S, name):
    d = np.load(os.path.join(GOLD, "shipped_mesh_arrays.npz"))
    return S.Mesh(*(np.ascontiguousarray(d[f"{name}_{k}"]) for k in ("positions", "normals", "uvs", "tangents", "indices", "submeshes")))


class FlatScene:
    """the front-end's buffers + the mesh pools, in the shape OracleScene expects from scenes.Scene"""

    def __init__(self, base, flat):
        self.width, self.height, self.textures = base.width, base.height, []
        self._flat = dict(base.flatten())
        for k in ("transforms", "prev_transforms", "instances", "blas_of_instance", "emissive_idx", "materials", "lights"):
            self._flat[k] = flat[k]

    def flatten(self):
        return self._flat


def bits_equal(a, b):
    return ((a.view(np.uint32) == b.view(np.uint32)) | ((a != a) & (b != b))).all(-1)


def render_and_compare(snk, oracle, base, fe, pt, mode, frames, animate):
    settings = snk.api.default_settings(mode=mode)
    fr = None
    lit = 0.0
    for f in range(frames):
        animate(f)
        fe.frame_start(1000.0 * f / 60.0)
        flat = fe.flat()
        fe.push(pt)
        pt.trace_primary()
        pt.render(settings)
        osc = oracle.OracleScene(FlatScene(base, flat))
        if fr is None:
            fr = oracle.OracleFrame(osc, base.width, base.height)
        else:
            oracle.lib().snkor_frame_set_scene(fr.h, osc.h)
            fr.osc = osc
        fr.set(flat["ubo"], flat["prev_ubo"], oracle.default_settings(mode=mode))
        fr.gbuffer()
        fr.render()
        fe.frame_end()
        for name in ("prim", "inst", "depth", "normal", "albedo"):
            assert pt.read_aux(name).tobytes() == fr.read(name).tobytes(), f"mode {mode} frame {f}: G-buffer plane {name}"
        a, b = pt.read_aux("radiance")[..., :3], fr.read("radiance")[..., :3]
        eq = bits_equal(a, b)
        assert eq.mean() >= 0.9995, f"mode {mode} frame {f}: {(~eq).sum()} of {eq.size} pixels differ"
        st = pt.stats()
        oc, oa = fr.ray_counts(reset=True)
        assert (int(st["rays_secondary"]), int(st["rays_shadow"])) == (oc, oa)
        assert int(st["stack_overflows"]) == 0
        lit = float((a > 0).any(-1).mean())
    return lit


@pytest.mark.parametrize("mode", [0, 1])
def test_project_template_scene_renders_like_the_oracle(snk, oracle, mode):
    """my_scene.json: one entity with the core sphere mesh at the origin, the default material, directional light (0, 0)"""
    S = snk.scenes
    assets = importlib.import_module("snake-vk_b200.assets")
    W, H = 160, 120
    sphere = shipped_mesh(S, "sphere")
    assert (sphere.positions.shape[0], sphere.num_tris) == (559, 960)
    mats = S.make_materials([((0.8, 0.8, 0.8), 0.5, 0.0, 0)])
    base = S.Scene("template", W, H, [sphere], [S.Entity(0, np.eye(4, dtype=np.float32), [0])], mats, S.Camera((0, 0, -4), (0, 0, 1)))
    facts = json.load(open(os.path.join(GOLD, "shipped_scene.json")))["scene"]
    scene = dict(facts, entities=[dict(x, pos=tuple(x["pos"]), scale=tuple(x["scale"]), orientation=tuple(x["orientation"])) for x in facts["entities"]])
    pt = snk.PathTracer(W, H)
    fe = snk.frontend.SceneFrontEnd()
    mesh_id = fe.register_mesh(pt.upload_mesh(sphere), sphere.submeshes)
    fe.set_material(0, mats[0:1])
    ids = assets.instantiate_scene(scene, fe, {assets.CORE_SPHERE_MESH: (mesh_id, 1)}, {assets.CORE_MATERIAL: 0})
    assert len(ids) == 1
    # the template's light direction (spherical 0, 0) is +z: look at the lit side of the sphere from z = +3 (default orientation faces -z)
    cam = fe.create_entity(pos=(0.0, 0.3, 3.0))
    fe.set_camera(cam, 90.0, W / H)
    lit = render_and_compare(snk, oracle, base, fe, pt, mode, 3, lambda f: fe.set_transform(cam, pos=(0.05 * f, 0.3, 3.0 - 0.1 * f)))
    assert lit > 0.05, "the sphere is lit by the template's directional light"
    pt.close()
    fe.close()


@pytest.mark.parametrize("mode", [0, 1])
def test_shipped_cube_and_sphere_scene(snk, oracle, mode):
    """both shipped meshes instanced several times (the cube also as floor and as an emissive light), a point light, camera dolly"""
    S = snk.scenes
    W, H = 160, 96
    sphere, cube = shipped_mesh(S, "sphere"), shipped_mesh(S, "cube")
    assert (cube.positions.shape[0], cube.num_tris) == (24, 12)
    mats = S.make_materials([((0.7, 0.7, 0.7), 0.6, 0.0, 0), ((0.9, 0.3, 0.2), 0.4, 0.1, 0), ((0.2, 0.5, 0.9), 0.3, 0.6, 0),
                             ((12.0, 10.0, 8.0), 0.5, 0.0, S.MAT_FLAG_EMISSIVE)])
    base = S.Scene("shipped", W, H, [sphere, cube], [S.Entity(0, np.eye(4, dtype=np.float32), [0])], mats, S.Camera((0, 2, -8), (0, 0, 1)))
    pt = snk.PathTracer(W, H)
    fe = snk.frontend.SceneFrontEnd()
    m_sphere = fe.register_mesh(pt.upload_mesh(sphere), sphere.submeshes)
    m_cube = fe.register_mesh(pt.upload_mesh(cube), cube.submeshes)
    for i in range(mats.shape[0]):
        fe.set_material(i, mats[i:i + 1])
    floor = fe.create_entity(pos=(0.0, -1.0, 0.0), scale=(12.0, 1.0, 12.0))
    fe.set_static_mesh(floor, m_cube, [0])
    hub = fe.create_entity(pos=(0.0, 1.5, 0.0))
    for k in range(5):
        e = fe.create_entity(parent=hub, pos=(3.0 * np.cos(k * 1.2566), 0.4 * k, 3.0 * np.sin(k * 1.2566)), euler=(15.0 * k, 30.0, 5.0), scale=(0.8 + 0.1 * k,) * 3)
        fe.set_static_mesh(e, m_sphere if k % 2 == 0 else m_cube, [1 + k % 2])
    lamp_cube = fe.create_entity(pos=(0.0, 4.5, 1.0), scale=(0.4, 0.1, 0.4))
    fe.set_static_mesh(lamp_cube, m_cube, [3])
    lamp = fe.create_entity(pos=(-3.0, 3.5, -2.0))
    fe.set_pointlight(lamp, (9.0, 8.0, 7.0))
    fe.set_directional_light((1.0, 0.95, 0.9), 0.9, 0.7)
    cam = fe.create_entity(pos=(0.0, 2.5, -8.0), euler=(0.0, 180.0, 0.0))
    fe.set_camera(cam, 90.0, W / H)

    def animate(f):
        fe.set_transform(hub, euler=(0.0, 10.0 * f, 0.0))
        fe.set_transform(cam, pos=(0.04 * f, 2.5, -8.0 + 0.1 * f))

    lit = render_and_compare(snk, oracle, base, fe, pt, mode, 3, animate)
    assert lit > 0.3
    pt.close()
    fe.close()
