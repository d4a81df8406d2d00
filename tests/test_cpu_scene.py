"""Scene front-end (include/snkscene.h, snake-vk_b200/host/snkscene.cpp — SURVEY §8(f) row N2) against (a) the reference's
vendored glm through oracle/_ref when /root/reference is mounted, (b) the committed glm golden vectors, (c) independent
float64 numpy restatements of the hierarchy rules, and (d) the buffer-system semantics (slot allocation, emissive list,
previous-frame transforms, UBO hand-over). CPU only."""
import ctypes as C
import os

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")


@pytest.fixture()
def fe(snk):
    f = snk.frontend.SceneFrontEnd()
    yield f
    f.close()


def fp(a):
    return a.ctypes.data_as(C.c_void_p)


def test_header_symbols_are_exported(snk):
    import re
    hdr = open(os.path.join(ROOT, "include", "snkscene.h")).read()
    names = set(re.findall(r"SNKSCENE_API[^;(]*?\b(snkscene_\w+)\s*\(", hdr))
    L = snk.frontend.lib()
    assert len(names) >= 28 and all(hasattr(L, n) for n in names), [n for n in names if not hasattr(L, n)]


def test_root_matrices_match_glm(snk, fe):
    S = snk.scenes
    g = np.load(os.path.join(GOLD, "glm_vectors.npz"))
    for i in range(len(g["pos"])):
        e = fe.create_entity(pos=g["pos"][i], euler=g["euler"][i], scale=g["scale"][i])
        assert np.allclose(fe.matrix(e).reshape(16), g["trs_out"][i], rtol=2e-6, atol=2e-6)
    # the same through the reference's own glm where it is mounted (authoring container): bit for bit
    so = os.path.join(ROOT, "oracle", "_ref", "libsnkref_glm.so")
    if os.path.exists(so):
        ref = C.CDLL(so)
        rng = np.random.default_rng(5)
        worst = 0
        for _ in range(200):
            pos, eul, scl = (rng.uniform(-20, 20, 3).astype(np.float32), rng.uniform(-360, 360, 3).astype(np.float32), rng.uniform(0.1, 4, 3).astype(np.float32))
            want = np.zeros(16, np.float32)
            ref.ref_trs(fp(pos), fp(eul), fp(scl), fp(want))
            e = fe.create_entity(pos=pos, euler=eul, scale=scl)
            got = fe.matrix(e).reshape(16)
            worst = max(worst, int(np.abs(got.view(np.int32) - want.view(np.int32)).max()))
        assert worst <= 2, f"root TRS differs from glm by {worst} ulp"


def euler_R(euler_deg):
    """float64 rotation of ExtraMath::Rotate3D = mat3_cast(quat(radians(euler))) as a math matrix"""
    e = np.radians(np.asarray(euler_deg, np.float64)) * 0.5
    c, s = np.cos(e), np.sin(e)
    w = c[0] * c[1] * c[2] + s[0] * s[1] * s[2]
    x = s[0] * c[1] * c[2] - c[0] * s[1] * s[2]
    y = c[0] * s[1] * c[2] + s[0] * c[1] * s[2]
    z = c[0] * c[1] * s[2] - s[0] * s[1] * c[2]
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y)],
                     [2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x)],
                     [2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)]])


def math_matrix(m_cm):
    return np.asarray(m_cm, np.float64).T  # [col][row] storage -> math matrix


def test_hierarchy_rules(snk, fe):
    """TransformSystem.cpp:31-114: child = parent * T(pos) R S', S' = total scale with the parent's scale divided out again per row;
    absolute scale / orientation accumulate over the ancestors; `absolute` entities ignore the parent"""
    ppos, peul, pscl = (3.0, 1.0, -2.0), (20.0, 45.0, -30.0), (2.0, 0.5, 1.5)
    cpos, ceul, cscl = (1.0, 2.0, 0.5), (-15.0, 60.0, 10.0), (0.5, 3.0, 1.25)
    p = fe.create_entity(pos=ppos, euler=peul, scale=pscl)
    c = fe.create_entity(parent=p, pos=cpos, euler=ceul, scale=cscl)
    gc = fe.create_entity(parent=c, pos=(0.2, 0.0, 0.1), euler=(5.0, 5.0, 5.0), scale=(2.0, 2.0, 2.0))
    P = np.eye(4); P[:3, :3] = euler_R(peul) * np.asarray(pscl); P[:3, 3] = ppos
    assert np.allclose(math_matrix(fe.matrix(p)), P, atol=1e-5)
    ps = np.asarray(pscl)
    inv = np.array([ps[1] * ps[2], ps[0] * ps[2], ps[1] * ps[0]]) / ps.prod()
    L = np.eye(4); L[:3, :3] = euler_R(ceul) * (np.asarray(cscl) * ps)[None, :] * inv[:, None]; L[:3, 3] = cpos
    assert np.allclose(math_matrix(fe.matrix(c)), P @ L, atol=2e-5)
    a = fe.abs(c)
    assert np.allclose(a["scale"], np.asarray(cscl) * ps) and np.allclose(a["orientation"], np.asarray(ceul) + np.asarray(peul))
    assert np.allclose(a["pos"], (P @ L)[:3, 3], atol=2e-5)
    g = fe.abs(gc)
    assert np.allclose(g["scale"], 2.0 * np.asarray(cscl) * ps) and np.allclose(g["orientation"], np.asarray((5.0,) * 3) + np.asarray(ceul) + np.asarray(peul))
    # moving the parent moves the subtree (UpdateChildren)
    before = fe.matrix(gc).copy()
    fe.set_transform(p, pos=(10.0, 0.0, 0.0))
    assert np.allclose(fe.matrix(gc)[3][:3] - before[3][:3], (7.0, -1.0, 2.0), atol=1e-4)
    # an absolute child is built like a root
    fe.set_absolute(c, True)
    R = np.eye(4); R[:3, :3] = euler_R(ceul) * np.asarray(cscl); R[:3, 3] = cpos
    assert np.allclose(math_matrix(fe.matrix(c)), R, atol=1e-5)
    with pytest.raises(snk.SnkError):
        fe.set_parent(p, gc)  # cycle


def build_small(snk, fe):
    """the instanced test scene of scenes.small_test_scene fed through the front-end, entity by entity"""
    S = snk.scenes
    rng = np.random.default_rng(3)
    mats = S.make_materials([((0.8, 0.8, 0.8), 0.5, 0.0, 0), ((0.9, 0.2, 0.2), 0.3, 0.0, 0), ((0.0, 20.0, 0.0), 0.5, 0.0, S.MAT_FLAG_EMISSIVE),
                             ((20.0, 0.0, 0.0), 0.5, 0.0, S.MAT_FLAG_EMISSIVE)])
    for i in range(mats.shape[0]):
        fe.set_material(i, mats[i:i + 1])
    entry_dt = np.dtype([("data_start_indices_idx", "<u4"), ("data_start_vertex_idx", "<u4"), ("num_submeshes", "<u4"), ("blas_id", "<u4", (64,))])
    meshes = []
    for k, (ni, nv, i0, v0) in enumerate([(6, 4, 0, 0), (960, 162, 6, 4), (36, 24, 966, 166)]):
        e = np.zeros((), entry_dt)
        e["data_start_indices_idx"], e["data_start_vertex_idx"], e["num_submeshes"] = i0, v0, 1
        e["blas_id"][0] = k
        meshes.append(fe.register_mesh(e, np.array([[ni, nv, 0, 0, 0]], np.uint32)))
    ents = [fe.create_entity()]
    fe.set_static_mesh(ents[0], meshes[0], [0])
    for i in range(4):
        e = fe.create_entity(pos=rng.uniform(-5, 5, 3), euler=rng.uniform(0, 360, 3), scale=(float(rng.uniform(0.5, 2)),) * 3)
        fe.set_static_mesh(e, meshes[1], [1])
        ents.append(e)
    for i in range(3):
        e = fe.create_entity(pos=rng.uniform(-5, 5, 3), scale=(0.1, 0.1, 0.1))
        fe.set_static_mesh(e, meshes[2], [2 + i % 2])
        ents.append(e)
    lights = []
    for i in range(3):
        e = fe.create_entity(pos=rng.uniform(-5, 5, 3))
        fe.set_pointlight(e, rng.uniform(5, 10, 3))
        lights.append(e)
    fe.set_directional_light((1.0, 1.0, 1.0), 0.8, 0.9)
    cam = fe.create_entity(pos=(0.0, 2.0, -9.0), euler=(0.0, 180.0, 0.0))
    fe.set_camera(cam, 90.0, 96 / 64)
    return dict(ents=ents, lights=lights, cam=cam, mats=mats)


def test_flat_buffers(snk, fe):
    S = snk.scenes
    b = build_small(snk, fe)
    fe.frame_start(0.0)
    f = fe.flat()
    n_ent = len(b["ents"]) + len(b["lights"]) + 1
    assert f["transforms"].shape == (n_ent, 16) and np.array_equal(f["transforms"], f["prev_transforms"])  # new slots: prev = cur
    inst = f["instances"]
    assert inst.shape == (8, 5) and np.array_equal(inst[:, 0], np.arange(8))           # transform_idx = entity slot, creation order
    assert np.array_equal(inst[0], (0, 0, 0, 0, 6)) and np.array_equal(inst[1], (1, 6, 4, 1, 960)) and np.array_equal(inst[5], (5, 966, 166, 2, 36))
    assert np.array_equal(f["blas_of_instance"], [0, 1, 1, 1, 1, 2, 2, 2])
    assert np.array_equal(f["emissive_idx"], [7, 6, 5])                                 # registry view = reverse insertion order
    assert np.array_equal(f["materials"]["flags"], b["mats"]["flags"]) and np.array_equal(f["materials"]["albedo"], b["mats"]["albedo"])
    # LightBufferSystem layout: 96 B directional, counts @96, 48-B point lights @112 in reverse insertion order
    lights = np.frombuffer(f["lights"], np.float32)
    assert np.frombuffer(f["lights"], np.uint32)[24] == 3 and np.frombuffer(f["lights"], np.uint32)[25] == 0
    assert np.allclose(lights[4:7], S.spherical_to_cartesian(0.8, 0.9), atol=1e-6) and lights[3] == 1.0 and lights[7] == 1.0
    for k, e in enumerate(reversed(b["lights"])):
        assert np.allclose(lights[28 + 12 * k + 4:28 + 12 * k + 7], fe.abs(e)["pos"]) and np.allclose(lights[28 + 12 * k + 8:28 + 12 * k + 11], (0.05, 0.25, 0.5))
    # SceneInfoBufferSystem::UpdateUBO against the glm-pinned numpy producers
    u = f["ubo"]
    fwd = fe.abs(b["cam"])["forward"]
    assert np.allclose(fwd, (0.0, 0.0, 1.0), atol=1e-6)                                 # yaw 180: looking down +z
    assert np.allclose(u["view"], S.look_at((0.0, 2.0, -9.0), np.asarray((0.0, 2.0, -9.0)) + fwd).reshape(16), atol=1e-5)
    assert np.allclose(u["proj"], S.perspective_vk(90.0, 96 / 64, 0.1, 2500.0).reshape(16), rtol=1e-6)
    assert np.allclose(u["proj_view"], S.mat_mul_cm(u["proj"].reshape(4, 4), u["view"].reshape(4, 4)).reshape(16), rtol=1e-5, atol=1e-5)
    assert int(u["frame_idx"]) == 0 and int(u["prev_frame_invalid"]) == 1 and np.allclose(u["cam_pos"], (0.0, 2.0, -9.0, 1.0))


def test_frame_to_frame_semantics(snk, fe):
    b = build_small(snk, fe)
    fe.frame_start(0.0); fe.frame_end()
    f0 = fe.flat()
    mover = b["ents"][2]
    fe.set_transform(mover, pos=(9.0, 9.0, 9.0))
    fe.set_transform(b["cam"], pos=(0.0, 2.0, -8.5))
    fe.frame_start(16.6667); fe.frame_end()
    f1 = fe.flat()
    assert np.array_equal(f1["prev_transforms"], f0["transforms"])                      # prev = last frame's cur
    assert np.allclose(f1["transforms"][mover][12:15], (9.0, 9.0, 9.0)) and not np.array_equal(f1["transforms"][mover], f0["transforms"][mover])
    assert np.array_equal(np.delete(f1["transforms"], [mover, b["cam"]], 0), np.delete(f0["transforms"], [mover, b["cam"]], 0))
    assert f1["prev_ubo"].tobytes() == f0["ubo"].tobytes() and int(f1["ubo"]["frame_idx"]) == 1 and int(f1["ubo"]["prev_frame_invalid"]) == 0
    fe.frame_start(33.3333); fe.frame_end()
    f2 = fe.flat()
    assert np.array_equal(f2["prev_transforms"], f2["transforms"])                      # nothing moved: zero motion
    # updating a StaticMeshComponent moves the entity to fresh instance slots; the old ones stay but leave the TLAS
    fe.set_static_mesh(b["ents"][1], 1, [0])
    fe.invalidate_history()
    fe.frame_start(50.0)
    f3 = fe.flat()
    assert f3["instances"].shape[0] == 9 and f3["blas_of_instance"][1] == 0xFFFFFFFF and f3["blas_of_instance"][8] == 1
    assert np.array_equal(f3["instances"][8], (1, 6, 4, 0, 960)) and int(f3["ubo"]["prev_frame_invalid"]) == 1
    fe.remove_static_mesh(b["ents"][7])                                                  # an emissive cube leaves
    fe.frame_start(66.0)
    assert np.array_equal(fe.flat()["emissive_idx"], [6, 5]) and fe.flat()["blas_of_instance"][7] == 0xFFFFFFFF


def test_cpp_host_program_builds_and_refuses_without_gpu(snk):
    """host/snk_rt.hpp (C++ mirror of SNAKE::RT) + host/example_rt.cpp compile against the installed headers and link both
    libraries; without a CUDA device the program must fail loudly (no CPU fallback), with one it renders (tests/test_gpu_scene.py)"""
    import subprocess
    exe = os.path.join(ROOT, "snake-vk_b200", "example_rt")
    subprocess.run(["make", "-C", os.path.join(ROOT, "snake-vk_b200"), "-s", "example_rt"], check=True)
    assert os.path.exists(exe)
    import torch
    if not torch.cuda.is_available():
        r = subprocess.run([exe], capture_output=True, text=True)
        assert r.returncode == 1 and "no usable CUDA device" in r.stderr and "no CPU path" in r.stderr
