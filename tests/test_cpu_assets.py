"""On-disk formats of the engine (snake-vk_b200/assets.py): ByteSerializer streams (.meshdata, .mat), JSON
(.smesh, scene files). Pinned on (a) byte layouts written out by hand from ByteSerializer.h / AssetLoader.cpp, (b) the scene
file the editor's project template ships (facts committed in tests/golden/shipped_scene.json, re-read from /root/reference when
mounted), (c) write → read round trips. CPU only."""
import json
import os
import struct

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")


@pytest.fixture()
def assets(snk):
    import importlib
    return importlib.import_module("snake-vk_b200.assets")


def test_material_byte_layout(assets, tmp_path):
    # AssetLoader.cpp:127-143: u64 uuid | u64 n + name | vec3 albedo | emissive roughness metallic ao | 5 x u64 texture uuids
    raw = struct.pack("<Q", 77) + struct.pack("<Q", 3) + b"red" + struct.pack("<3f", 0.9, 0.1, 0.2) + struct.pack("<4f", 2.0, 0.3, 1.0, 0.5)
    raw += struct.pack("<5Q", 11, 0, 0, 0, 12)
    p = tmp_path / "red.mat"
    p.write_bytes(raw)
    m = assets.load_material(str(p))
    assert m["uuid"] == 77 and m["name"] == "red" and np.allclose(m["albedo"], (0.9, 0.1, 0.2)) and m["albedo_tex"] == 11 and m["ao_tex"] == 12
    assert np.allclose((m["emissive"], m["roughness"], m["metallic"], m["ao"]), (2.0, 0.3, 1.0, 0.5))
    q = tmp_path / "copy.mat"
    assets.save_material(str(q), 77, "red", (0.9, 0.1, 0.2), 2.0, 0.3, 1.0, 0.5, albedo_tex=11, ao_tex=12)
    assert q.read_bytes() == raw


def test_meshdata_round_trip_and_layout(snk, assets, tmp_path):
    S = snk.scenes
    mesh = S.merge_meshes([S.box_mesh(), S.icosphere_mesh(1)], [np.eye(4, dtype=np.float32), S.trs((3, 0, 0))])
    mesh.submeshes = np.array([[36, 24, 0, 0, 0], [mesh.indices.size - 36, mesh.positions.shape[0] - 24, 24, 36, 1]], np.uint32)
    mesh.uvs = np.random.default_rng(1).random(mesh.uvs.shape).astype(np.float32)
    p = str(tmp_path / "two.meshdata")
    assets.save_meshdata(p, mesh, uuid=0xABCDEF0123456789, name="two parts", material_uuids=[3, 99], texture_uuids=[4])
    raw = open(p, "rb").read()
    # header by hand (AssetLoader.cpp:145-155): uuid, name container, submesh container (2 x 20 B), num_vertices, num_indices, positions
    assert struct.unpack_from("<Q", raw, 0)[0] == 0xABCDEF0123456789 and struct.unpack_from("<Q", raw, 8)[0] == 9 and raw[16:25] == b"two parts"
    assert struct.unpack_from("<Q", raw, 25)[0] == 40 and struct.unpack_from("<5I", raw, 33) == (36, 24, 0, 0, 0)
    nv, ni = struct.unpack_from("<2I", raw, 73)
    assert (nv, ni) == (mesh.positions.shape[0], mesh.indices.size) and struct.unpack_from("<Q", raw, 81)[0] == 12 * nv
    d = assets.load_meshdata(p)
    m = d["mesh"]
    assert d["uuid"] == 0xABCDEF0123456789 and d["name"] == "two parts" and list(d["materials"]) == [3, 99] and list(d["textures"]) == [4]
    for a, b in ((m.positions, mesh.positions), (m.normals, mesh.normals), (m.tangents, mesh.tangents), (m.uvs, mesh.uvs), (m.indices, mesh.indices), (m.submeshes, mesh.submeshes)):
        assert np.array_equal(a, b)
    open(p, "wb").write(raw[:200])
    with pytest.raises(ValueError):
        assets.load_meshdata(p)  # truncated file: the reference asserts (ByteSerializer.h:92)


def test_smesh_and_scene_round_trip(assets, tmp_path):
    p = str(tmp_path / "a.smesh")
    assets.save_smesh(p, 10, 20, "ball")
    assert assets.load_smesh(p) == dict(uuid=10, data_uuid=20, name="ball") and json.load(open(p))["DataUUID"] == 20
    scene = dict(uuid=5, dir_colour=(1.0, 0.5, 0.25), dir_spherical=(0.8, 0.9), entities=[
        dict(uuid=100, name="hub", parent=0, pos=(0.0, 2.0, 1.0), scale=(1.5, 1.0, 1.0), orientation=(0.0, 30.0, 0.0)),
        dict(uuid=101, name="arm", parent=100, pos=(2.0, 0.0, 0.0), scale=(0.5, 0.5, 0.5), orientation=(10.0, 0.0, 0.0), mesh=1, materials=[3]),
        dict(uuid=102, name="lamp", parent=100, pos=(0.0, 2.5, -1.0), pointlight=dict(colour=(8.0, 7.0, 6.0), attenuation=(0.5, 0.25, 0.05)))])
    q = str(tmp_path / "s.json")
    assets.save_scene(q, scene)
    back = assets.load_scene(q)
    assert back["uuid"] == 5 and back["dir_colour"] == (1.0, 0.5, 0.25) and [e["uuid"] for e in back["entities"]] == [100, 101, 102]
    assert back["entities"][1]["mesh"] == 1 and back["entities"][1]["materials"] == [3] and back["entities"][1]["parent"] == 100
    assert back["entities"][2]["pointlight"] == dict(colour=(8.0, 7.0, 6.0), attenuation=(0.5, 0.25, 0.05))
    doc = json.load(open(q))
    assert set(doc) == {"SceneUUID", "Entities", "DirLight"} and doc["Entities"][1]["Transform"]["Abs"] is False and doc["Entities"][0]["ParentUUID"] == 0


def test_shipped_project_template_scene(snk, assets):
    facts = json.load(open(os.path.join(GOLD, "shipped_scene.json")))["scene"]
    assert facts["uuid"] == 10979679584156046672 and len(facts["entities"]) == 1
    e = facts["entities"][0]
    assert e["uuid"] == 171721715750585366 and e["mesh"] == assets.CORE_SPHERE_MESH and e["materials"] == [] and e["parent"] == 0
    ref = "/root/reference/Editor/res/project-template/res/scenes/my_scene.json"
    if os.path.exists(ref):  # authoring container: the loader reproduces the committed facts from the real file
        got = assets.load_scene(ref)
        assert json.loads(json.dumps(got)) == facts
    # the scene drives the front-end: the core sphere asset, default material for the missing `Materials` key
    fe = snk.frontend.SceneFrontEnd()
    entry_dt = np.dtype([("a", "<u4"), ("b", "<u4"), ("n", "<u4"), ("blas", "<u4", (64,))])
    entry = np.zeros((), entry_dt); entry["n"] = 1
    sphere = fe.register_mesh(entry, np.array([[2880, 559, 0, 0, 0]], np.uint32))   # sphere.glb: 960 triangles, 559 vertices
    mat = snk.scenes.make_materials([((0.8, 0.8, 0.8), 0.5, 0.0, 0)])
    fe.set_material(0, mat[0:1])
    scene = dict(facts, entities=[dict(x, pos=tuple(x["pos"]), scale=tuple(x["scale"]), orientation=tuple(x["orientation"])) for x in facts["entities"]])
    ids = assets.instantiate_scene(scene, fe, {assets.CORE_SPHERE_MESH: (sphere, 1)}, {assets.CORE_MATERIAL: 0})
    cam = fe.create_entity(pos=(0, 0, -4), euler=(0, 180, 0))
    fe.set_camera(cam)
    fe.frame_start(0.0)
    f = fe.flat()
    assert np.array_equal(f["instances"], [[ids[e["uuid"]], 0, 0, 0, 2880]]) and np.array_equal(f["transforms"][0], np.eye(4, dtype=np.float32).reshape(16))
    lights = np.frombuffer(f["lights"], np.float32)
    assert np.allclose(lights[4:7], (0.0, 0.0, 1.0)) and np.allclose(lights[0:3], (1, 1, 1))     # spherical (0, 0) = +z
    fe.close()


def test_committed_mesh_arrays_equal_the_shipped_glb_files(snk, assets):
    """tests/golden/shipped_mesh_arrays.npz (what the GPU tests render: the box has no /root/reference) = load_glb of the .glb files"""
    d = np.load(os.path.join(GOLD, "shipped_mesh_arrays.npz"))
    facts = json.load(open(os.path.join(GOLD, "shipped_meshes.json")))
    for name in ("cube", "sphere"):
        assert d[f"{name}_positions"].shape[0] == facts[name]["num_vertices"] and d[f"{name}_indices"].size == 3 * facts[name]["num_triangles"]
        assert np.allclose(d[f"{name}_positions"].min(0), facts[name]["bounds_lo"]) and np.allclose(d[f"{name}_positions"].max(0), facts[name]["bounds_hi"])
        assert np.allclose(np.linalg.norm(d[f"{name}_normals"], axis=1), 1.0, atol=1e-3)
        ref = f"/root/reference/Core/res/meshes/{name}.glb"
        if os.path.exists(ref):
            m = assets.load_glb(ref)
            for k in ("positions", "normals", "tangents", "uvs", "indices", "submeshes"):
                assert np.array_equal(getattr(m, k), d[f"{name}_{k}"]), (name, k)
