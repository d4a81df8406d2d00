"""Generates the committed fixtures under tests/golden/. Run in the authoring container
(needs /root/reference for the glm-backed part and the shipped meshes):

    python tools/make_golden.py

  glm_vectors.npz     outputs of the reference's vendored glm (oracle/_ref/libsnkref_glm.so, compiled
                      against /root/reference/Core/extern/glm) for the host-side matrix producers:
                      projection, view, TRS, instance 3x4, inverse, normal matrix + determinant, spherical
  shipped_meshes.json facts parsed from Core/res/meshes/{cube,sphere}.glb (counts, bounds) and, for the
                      sphere, analytic-vs-mesh hit statistics of the brute-force intersector
  trace_cornell.npz   seeded rays + brute-force closest / any hits on C1 (oracle)
  reference_frame_vectors.npz  main() of RayGen.rgen (+ closest-hit / miss shaders) EXECUTED from the reference's shader text for
                      40 pixels x 4 frames x both passes (tools/glsl_frame.py): temporal + final reservoirs, colour, ray counts
  frame_small.npz     3 consecutive reference-mode frames of the small test scene (oracle):
                      prim ids, depth, both reservoir halves, direct, radiance, ray counts
"""
import ctypes as C
import importlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle_api as O  # noqa: E402

snk_scenes = importlib.import_module("snake-vk_b200.scenes")
assets = importlib.import_module("snake-vk_b200.assets")
GOLD = os.path.join(ROOT, "tests", "golden")
REF = "/root/reference"


def fp(a):
    return a.ctypes.data_as(C.c_void_p)


def glm_vectors():
    L = C.CDLL(os.path.join(ROOT, "oracle", "_ref", "libsnkref_glm.so"))
    L.ref_projection.argtypes = [C.c_float] * 4 + [C.c_void_p]
    L.ref_spherical.argtypes = [C.c_float, C.c_float, C.c_void_p]
    rng = np.random.default_rng(0x5EED)
    proj_in = np.array([[90, 16 / 9, 0.1, 2500], [90, 1.0, 0.1, 2500], [60, 4 / 3, 0.05, 100], [120, 2.0, 1.0, 10000]], np.float32)
    proj_out = np.zeros((len(proj_in), 16), np.float32)
    for i, p in enumerate(proj_in):
        L.ref_projection(*[C.c_float(float(x)) for x in p], fp(proj_out[i]))
    n = 16
    pos = rng.uniform(-20, 20, (n, 3)).astype(np.float32)
    fwd = rng.normal(size=(n, 3)).astype(np.float32)
    fwd /= np.linalg.norm(fwd, axis=1, keepdims=True)
    fwd[0] = (0, 0, 1)
    fwd[1] = (0.0, -0.05, 1.0) / np.linalg.norm((0.0, -0.05, 1.0))
    view_out = np.zeros((n, 16), np.float32)
    for i in range(n):
        L.ref_view(fp(pos[i]), fp(fwd[i]), fp(view_out[i]))
    euler = rng.uniform(0, 360, (n, 3)).astype(np.float32)
    scale = rng.uniform(0.1, 3, (n, 3)).astype(np.float32)
    euler[0] = 0
    scale[0] = 1
    trs_out = np.zeros((n, 16), np.float32)
    inst_out = np.zeros((n, 12), np.float32)
    inv_out = np.zeros((n, 16), np.float32)
    nm_out = np.zeros((n, 9), np.float32)
    det_out = np.zeros(n, np.float32)
    mul_out = np.zeros((n, 16), np.float32)
    for i in range(n):
        L.ref_trs(fp(pos[i]), fp(euler[i]), fp(scale[i]), fp(trs_out[i]))
        L.ref_instance_transform(fp(trs_out[i]), fp(inst_out[i]))
        L.ref_inverse(fp(trs_out[i]), fp(inv_out[i]))
        L.ref_normal_matrix_det(fp(trs_out[i]), fp(nm_out[i]), fp(det_out[i:i + 1]))
        L.ref_mul(fp(proj_out[i % len(proj_out)]), fp(view_out[i]), fp(mul_out[i]))
    sph_in = np.array([[0.8, 0.9], [0.0, 0.0], [1.0, 2.0], [3.0, 0.3]], np.float32)
    sph_out = np.zeros((len(sph_in), 3), np.float32)
    for i, (t, p) in enumerate(sph_in):
        L.ref_spherical(C.c_float(float(t)), C.c_float(float(p)), fp(sph_out[i]))
    np.savez_compressed(os.path.join(GOLD, "glm_vectors.npz"), proj_in=proj_in, proj_out=proj_out, pos=pos, fwd=fwd, view_out=view_out,
                        euler=euler, scale=scale, trs_out=trs_out, inst_out=inst_out, inv_out=inv_out, nm_out=nm_out, det_out=det_out,
                        mul_out=mul_out, sph_in=sph_in, sph_out=sph_out)


def shipped_scene():
    """facts of the scene file the editor's project template ships (Editor/res/project-template/res/scenes/my_scene.json)"""
    s = assets.load_scene(os.path.join(REF, "Editor", "res", "project-template", "res", "scenes", "my_scene.json"))
    json.dump({"_source": "facts parsed from Editor/res/project-template/res/scenes/my_scene.json by snake-vk_b200.assets.load_scene "
                          "(tools/make_golden.py:shipped_scene)", "scene": s}, open(os.path.join(GOLD, "shipped_scene.json"), "w"), indent=1)


def shipped_meshes():
    out = {}
    for name in ("cube", "sphere"):
        m = assets.load_glb(os.path.join(REF, "Core", "res", "meshes", f"{name}.glb"))
        out[name] = {"num_vertices": int(m.positions.shape[0]), "num_triangles": int(m.num_tris),
                     "bounds_lo": m.positions.min(0).tolist(), "bounds_hi": m.positions.max(0).tolist(),
                     "num_submeshes": int(m.submeshes.shape[0])}
        if name == "sphere":
            # rays from outside towards the origin region: mesh hit vs analytic unit sphere
            rng = np.random.default_rng(3)
            n = 4000
            rays = np.zeros(n, snk_scenes.RAY_DTYPE)
            o = rng.normal(size=(n, 3))
            o = 4 * o / np.linalg.norm(o, axis=1, keepdims=True)
            tgt = rng.uniform(-0.6, 0.6, (n, 3))
            d = tgt - o
            d /= np.linalg.norm(d, axis=1, keepdims=True)
            rays["o"], rays["d"], rays["tmin"], rays["tmax"] = o.astype(np.float32), d.astype(np.float32), 0.0, 100.0
            sc = snk_scenes.Scene("sphere", 8, 8, [m], [snk_scenes.Entity(0, np.eye(4, dtype=np.float32), [0])],
                                  snk_scenes.make_materials([((0.8, 0.8, 0.8), 0.5, 0.0, 0)]), snk_scenes.Camera((0, 0, -4), (0, 0, 1)))
            h = O.OracleScene(sc).trace(rays, brute=True)
            b = np.einsum("ij,ij->i", o, d)
            t_an = -b - np.sqrt(b * b - (np.einsum("ij,ij->i", o, o) - 1.0))
            assert (h["t"] > 0).all()
            out[name]["analytic_max_abs_dt"] = float(np.abs(h["t"] - t_an).max())  # faceting error of a 960-tri sphere
    json.dump(out, open(os.path.join(GOLD, "shipped_meshes.json"), "w"), indent=1)


def shipped_mesh_arrays():
    """vertex / index arrays of Core/res/meshes/{cube,sphere}.glb as snake-vk_b200.assets.load_glb reads them: the GPU box has no
    /root/reference, and tests/test_gpu_assets.py renders exactly these meshes -> tests/golden/shipped_mesh_arrays.npz"""
    out = {}
    for name in ("cube", "sphere"):
        m = assets.load_glb(os.path.join(REF, "Core", "res", "meshes", f"{name}.glb"))
        for k in ("positions", "normals", "tangents", "uvs", "indices", "submeshes"):
            out[f"{name}_{k}"] = np.ascontiguousarray(getattr(m, k))
    np.savez_compressed(os.path.join(GOLD, "shipped_mesh_arrays.npz"), **out)


def trace_cornell():
    sc = snk_scenes.cornell(64, 64)
    rng = np.random.default_rng(11)
    n = 4000
    rays = np.zeros(n, snk_scenes.RAY_DTYPE)
    rays["o"] = rng.uniform((0, 0, -2), (5.5, 5.5, 5.5), (n, 3)).astype(np.float32)
    d = rng.normal(size=(n, 3))
    rays["d"] = (d / np.linalg.norm(d, axis=1, keepdims=True)).astype(np.float32)
    rays["tmin"], rays["tmax"] = 0.001, 1000.0
    osc = O.OracleScene(sc)
    np.savez_compressed(os.path.join(GOLD, "trace_cornell.npz"), rays=rays, closest=osc.trace(rays, brute=True),
                        any=osc.trace(rays, any_hit=True, brute=True)["t"] >= 0)


def frame_small():
    sc = snk_scenes.small_test_scene(48, 32)
    osc = O.OracleScene(sc)
    fr = O.OracleFrame(osc)
    prev = None
    out = {}
    for f in range(3):
        pos = np.asarray(sc.camera.pos, np.float32) + np.asarray((0.05, 0.0, 0.1), np.float32) * f
        ubo = sc.ubo(f, cam_pos=pos)
        fr.set(ubo, ubo if prev is None else prev, O.default_settings())
        fr.gbuffer()
        fr.render()
        prev = ubo
        for name in ("prim", "inst", "depth", "reservoirs_temporal", "reservoirs_final", "direct", "radiance"):
            out[f"f{f}_{name}"] = fr.read(name)
        out[f"f{f}_rays"] = np.array(fr.ray_counts(reset=True), np.uint64)
    np.savez_compressed(os.path.join(GOLD, "frame_small.npz"), **out)


# (name, reference file, regex with ONE capture group, how the value is used here)
SOURCE_FACTS = [
    ("LIGHT_RIS_ITERS", "Core/res/shaders/RayGen.rgen", r"#define\s+LIGHT_RIS_ITERS\s+(\d+)u", "setting:ris_iters"),
    ("M_CAP", "Core/res/shaders/RayGen.rgen", r"#define\s+M_CAP\s+(\d+)", "setting:m_cap"),
    ("spatial neighbour radius (px)", "Core/res/shaders/RayGen.rgen", r"render_size \* FindNeighbourPixelCoord\(pixel_coord, ([0-9.]+),", "setting:neighbour_radius"),
    ("spatial normal threshold", "Core/res/shaders/RayGen.rgen", r"if \(dot\(neighbour_normal, normal\) < ([0-9.]+)\) continue;", "setting:normal_threshold"),
    ("temporal normal threshold", "Core/res/shaders/RayGen.rgen", r"if \(dot\(neighbour_normal, normal\) > ([0-9.]+)\) \{", "setting:normal_threshold"),
    ("num_temporal_resamples", "Core/headers/rendering/Raytracing.h", r"int32_t num_temporal_resamples = (\d+);", "setting:num_temporal"),
    ("num_spatial_resamples", "Core/headers/rendering/Raytracing.h", r"int32_t num_spatial_resamples = (\d+);", "setting:num_spatial"),
    ("PI", "Core/res/shaders/LightBuffers.glsl", r"#define PI ([0-9.]+)", "oracle:snk_raygen.h:PI_GLSL = {}f"),
    ("point-light WRS time factor", "Core/res/shaders/RayGen.rgen", r"fract\(weight\+common_ubo\.time_elapsed\*(0\.0009783) \+ rnd_seed\)", "oracle:snk_raygen.h:time() * {}f"),
    ("emissive WRS time factor", "Core/res/shaders/RayGen.rgen", r"fract\(weight\+common_ubo\.time_elapsed\*(0\.000789) \+ rnd_seed\)", "oracle:snk_raygen.h:time() * {}f"),
    ("light pick time factor", "Core/res/shaders/RayGen.rgen", r"fract\(common_ubo\.time_elapsed \* (0\.001197) \+ world_pos\.x", "oracle:snk_raygen.h:time() * {}f"),
    ("bary seed x", "Core/res/shaders/RayGen.rgen", r"random\(sin\(i\*(27934\.3221)\+gl_LaunchIDEXT\.x\*293\.28\)", "oracle:snk_raygen.h:(float)i * {}f"),
    ("bary seed y", "Core/res/shaders/RayGen.rgen", r"random\(sin\(i\*(12391\.8583)\+gl_LaunchIDEXT\.y\*983\.3399\)", "oracle:snk_raygen.h:(float)i * {}f"),
    ("spatial seed", "Core/res/shaders/RayGen.rgen", r"vec2\(seed, -seed\) \+ sin\(i\*(13477\.3793)\)", "oracle:snk_raygen.h:(float)i * {}f"),
    ("emissive point bias", "Core/res/shaders/RayGen.rgen", r"length\(surface_pos - \(p \+ point_normal \* ([0-9.]+)\)\)", "oracle:snk_raygen.h:* {}f"),
    ("biased_pos rule", "Core/res/shaders/RayGen.rgen", r"vec3 biased_pos = ro \+ normal \* \((0\.01 \+ max\(length\(common_ubo\.cam_pos\.xyz - ro\) \* 0\.001, 0\.01\))\);", "text"),
    ("rgen dir-light shadow tmax", "Core/res/shaders/RayGen.rgen", r"ShadowTest\(biased_pos, ssbo_light_data\.dir_light\.dir\.xyz, (\d+)\);", "oracle:snk_raygen.h:{}.0f"),
    ("rchit dir-light shadow tmax", "Core/res/shaders/RayClosestHit.rchit", r"ShadowTest\(biased_pos, ssbo_light_data\.dir_light\.dir\.xyz, (\d+)\);", "oracle:snk_raygen.h:{}.0f"),
    ("pointlights capacity", "Core/res/shaders/LightBuffers.glsl", r"Pointlight pointlights\[(\d+)\];", "layout:max_pointlights"),
    ("transform capacity", "Core/headers/scene/TransformBufferSystem.h", r"CreateBuffer\(sizeof\(glm::mat4\) \* (\d+),", "layout:max_transforms"),
    ("frames in flight", "Core/headers/core/VkContext.h", r"MAX_FRAMES_IN_FLIGHT = (\d+);", "text"),
    ("z_near", "Core/headers/components/CameraComponent.h", r"float z_near = ([0-9.]+)f;", "camera:z_near"),
    ("z_far", "Core/headers/components/CameraComponent.h", r"float z_far = ([0-9.]+)f;", "camera:z_far"),
    ("projection fov rule", "Core/headers/components/CameraComponent.h", r"glm::perspective\(glm::radians\((fov \* 0\.5f)\), aspect_ratio, z_near, z_far\)", "text"),
    ("particle_restitution", "Core/headers/scene/ParticleSystem.h", r"float particle_restitution = ([0-9.]+)f;", "physics:particle_restitution"),
    ("friction_coefficient", "Core/headers/scene/ParticleSystem.h", r"float friction_coefficient = ([0-9.]+);", "physics:friction_coefficient"),
    ("repulsion_factor", "Core/headers/scene/ParticleSystem.h", r"float repulsion_factor = ([0-9.]+)f;", "physics:repulsion_factor"),
    ("TAA epsilon", "Core/res/shaders/TAA_Resolve.comp", r"#define FLT_EPS ([0-9.]+)f", "oracle:snk_taa.h:eps = {}f"),
    ("TAA variance gamma", "Core/res/shaders/TAA_Resolve.comp", r"const float gamma = ([0-9.]+);", "oracle:snk_taa.h:gamma = {}f"),
    ("TAA sample count", "Core/res/shaders/TAA_Resolve.comp", r"one_divided_by_sample_count = (1\.0 / 9\.0);", "text"),
    ("TAA history weight", "Core/res/shaders/TAA_Resolve.comp", r"mix\(current_col, history_col, ([0-9.]+)\);", "oracle:snk_taa.h:, {}f);"),
    ("TAA disocclusion ramp", "Core/res/shaders/TAA_Resolve.comp", r"clamp\(\(velocity_length - ([0-9.]+)\) \* 10\.0, 0\.0, 1\.0\)", "oracle:snk_taa.h:(velocity_length - {}f) * 10.0f"),
    ("TAA initial extrema", "Core/res/shaders/TAA_Resolve.comp", r"vec3 min_col = vec3\(([0-9.]+)\);", "oracle:snk_taa.h:min_col = v3({}f)"),
    ("particle DT", "Core/res/shaders/Particle.glsl", r"#define DT \((1\.0 / 120\.0)\)", "oracle:snk_particles.h:PTCL_DT = (float)({})"),
    ("particle wall restitution", "Core/res/shaders/ParticleUpdate.rgen", r"const float restitution = ([0-9.]+);", "oracle:snk_particles.h:const float restitution = {}f;"),
    ("particle mass", "Core/res/shaders/ParticleUpdate.rgen", r"const float mass = ([0-9.]+);", "oracle:snk_particles.h:const float mass = {}f;"),
]


def reference_source_facts():
    """constants of the path read straight from the reference's shader / header text (first match, with its line number):
    tests/test_cpu_oracle.py checks that the oracle, the ABI defaults and the layouts carry exactly these values"""
    import re
    facts = []
    for name, rel, rx, use in SOURCE_FACTS:
        text = open(os.path.join(REF, rel), encoding="utf-8", errors="replace").read()
        m = re.search(rx, text)
        assert m, f"reference source fact not found: {name} in {rel}"
        facts.append({"name": name, "file": rel, "line": text.count("\n", 0, m.start()) + 1, "regex": rx, "value": m.group(1), "use": use})
    json.dump(facts, open(os.path.join(GOLD, "reference_source_facts.json"), "w"), indent=1)


GLSL_BLOCKS = [
    ("TriangleData", "Core/res/shaders/RayGen.rgen", r"struct\s+TriangleData\s*\{"),
    ("PixelReservoir", "Core/res/shaders/RayGen.rgen", r"struct\s+PixelReservoir\s*\{"),
    ("Material", "Core/res/shaders/TexMatBuffers.glsl", r"struct\s+Material\s*\{"),
    ("InstanceData", "Core/res/shaders/RaytracingInstances.glsl", r"struct\s+InstanceData\s*\{"),
    ("DirectionalLight", "Core/res/shaders/LightBuffers.glsl", r"struct\s+DirectionalLight\s*\{"),
    ("Pointlight", "Core/res/shaders/LightBuffers.glsl", r"struct\s+Pointlight\s*\{"),
    ("Spotlight", "Core/res/shaders/LightBuffers.glsl", r"struct\s+Spotlight\s*\{"),
    ("LightData", "Core/res/shaders/LightBuffers.glsl", r"readonly\s+buffer\s+LightData\s*\{"),
    ("CommonUBO", "Core/res/shaders/CommonUBO.glsl", r"uniform\s+CommonUBO\s*\{"),
]


def reference_layouts():
    """member lists (type, name, array length) of the GLSL blocks that form the data contract of the path, parsed from the
    reference's shader text; tests/test_cpu_oracle.py lays them out by the std430 / std140 rules and compares every offset with
    the wire formats used here"""
    import re
    out = {}
    for name, rel, rx in GLSL_BLOCKS:
        text = open(os.path.join(REF, rel), encoding="utf-8", errors="replace").read()
        m = re.search(rx, text)
        assert m, f"GLSL block not found: {name}"
        body = text[m.end():text.index("}", m.end())]
        body = re.sub(r"//[^\n]*", "", body)
        members = []
        for decl in body.split(";"):
            decl = decl.strip()
            if not decl:
                continue
            typ, rest = decl.split(None, 1)
            for var in rest.split(","):
                var = var.strip()
                am = re.match(r"(\w+)\[(\d+)\]$", var)
                members.append([typ, am.group(1), int(am.group(2))] if am else [typ, var, 1])
        out[name] = {"file": rel, "line": text.count("\n", 0, m.start()) + 1, "members": members}
    json.dump(out, open(os.path.join(GOLD, "reference_layouts.json"), "w"), indent=1)


def reference_shader_vectors():
    """the reference's own BRDF / light functions (Core/res/shaders/LightBuffers.glsl:65-157), executed from the shader text by
    tools/glsl_exec.py in float32 on 256 random surface / light configurations -> tests/golden/reference_shader_vectors.npz"""
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import glsl_exec as G
    names = ["FresnelSchlick", "DistributionGGX", "GeometrySchlickGGX", "GeometrySmith", "CookTorranceBRDF", "CalcDirectionalLight", "CalcPointlight"]
    env = G.load_functions(os.path.join(REF, "Core/res/shaders/LightBuffers.glsl"), names)
    rng = np.random.default_rng(20261018)
    n_cases = 256
    f32 = np.float32

    def unit(v):
        return (v / np.linalg.norm(v, axis=-1, keepdims=True)).astype(f32)

    n = unit(rng.normal(size=(n_cases, 3)))
    # view and light directions in the upper hemisphere of n (a few grazing ones included)
    v = unit(n * rng.uniform(0.02, 1.0, (n_cases, 1)) + 0.8 * rng.normal(size=(n_cases, 3)))
    l = unit(n * rng.uniform(0.02, 1.0, (n_cases, 1)) + 0.8 * rng.normal(size=(n_cases, 3)))
    rough = rng.uniform(0.05, 1.0, n_cases).astype(f32)
    metal = rng.uniform(0.0, 1.0, n_cases).astype(f32)
    albedo = rng.uniform(0.0, 1.0, (n_cases, 3)).astype(f32)
    f0 = (f32(0.04) * (f32(1.0) - metal[:, None]) + albedo * metal[:, None]).astype(f32)
    world_pos = rng.uniform(-5, 5, (n_cases, 3)).astype(f32)
    pl = np.zeros((n_cases, 12), f32)  # colour4, pos4, exp, linear, constant, pad
    pl[:, 0:3] = rng.uniform(0, 20, (n_cases, 3)); pl[:, 3] = 1
    pl[:, 4:7] = world_pos + l * rng.uniform(0.3, 8.0, (n_cases, 1)).astype(f32); pl[:, 7] = 1
    pl[:, 8:11] = rng.uniform(0.01, 1.0, (n_cases, 3))
    dl = np.zeros((n_cases, 8), f32)  # colour4, dir4 of the directional light
    dl[:, 0:3] = rng.uniform(0, 5, (n_cases, 3)); dl[:, 3] = 1
    dl[:, 4:7] = l; dl[:, 7] = 0
    out = dict(n=n, v=v, l=l, roughness=rough, metallic=metal, albedo=albedo, f0=f0, world_pos=world_pos, pointlight=pl, dirlight=dl)
    res = {k: [] for k in ("fresnel", "ggx", "smith", "cook_torrance", "calc_pointlight", "calc_dirlight")}
    for i in range(n_cases):
        h = env["normalize"](v[i] + l[i])
        res["fresnel"].append(env["FresnelSchlick"](env["max"](env["dot"](h, v[i]), 0.0), f0[i]))
        res["ggx"].append(env["DistributionGGX"](h, n[i], rough[i]))
        res["smith"].append(env["GeometrySmith"](v[i], l[i], n[i], rough[i]))
        res["cook_torrance"].append(env["CookTorranceBRDF"](l[i], np.zeros(3, f32), v[i], f0[i], n[i], rough[i], metal[i]))
        light = G.namespace(colour=pl[i, 0:4], pos=pl[i, 4:8], atten_exp=pl[i, 8], atten_linear=pl[i, 9], atten_constant=pl[i, 10])
        res["calc_pointlight"].append(env["CalcPointlight"](light, v[i], f0[i], world_pos[i], n[i], rough[i], metal[i], albedo[i]))
        env["ssbo_light_data"] = G.namespace(dir_light=G.namespace(colour=dl[i, 0:4], dir=dl[i, 4:8]))
        res["calc_dirlight"].append(env["CalcDirectionalLight"](v[i], f0[i], n[i], rough[i], metal[i], albedo[i]))
    for k, vals in res.items():
        out["ref_" + k] = np.asarray(vals, f32)
    # Util.glsl:23-57: integer hash, float construction, random(float) — bit-exact by nature
    util = G.load_functions(os.path.join(REF, "Core/res/shaders/Util.glsl"), ["hash", "floatConstruct", "random"])
    hx = np.concatenate([np.array([0, 1, 2, 0x7FFFFFFF, 0x80000000, 0xFFFFFFFF], np.uint32), rng.integers(0, 2 ** 32, 250, dtype=np.uint64).astype(np.uint32)])
    out["hash_in"] = hx
    out["ref_hash"] = np.array([int(util["hash"](G.U32(int(x)))) for x in hx], np.uint32)
    rx = np.concatenate([np.array([0.0, -0.0, 1.0, 0.5, 1e-30, 123456.789], f32), rng.normal(scale=100.0, size=250).astype(f32)])
    out["random_in"] = rx
    out["ref_random"] = np.array([util["random"](x) for x in rx], f32)
    # Util.glsl:77-82: gold_noise (tan in float64 rounded once = the numerics contract; chaotic, so compared statistically)
    gn = G.load_functions(os.path.join(REF, "Core/res/shaders/Util.glsl"), ["gold_noise"])
    gxy = np.stack([rng.integers(0, 1920, n_cases), rng.integers(0, 1080, n_cases)], -1).astype(f32)
    gseed = rng.uniform(0, 1, n_cases).astype(f32)
    out.update(gold_xy=gxy, gold_seed=gseed)
    out["ref_gold_noise"] = np.array([gn["gold_noise"](gxy[i], gseed[i]) for i in range(n_cases)], f32)
    # TAA_Resolve.comp:12-35: clip_aabb (history colour clipped towards the neighbourhood box)
    taa = G.load_functions(os.path.join(REF, "Core/res/shaders/TAA_Resolve.comp"), ["clip_aabb"])
    lo = rng.uniform(0, 1, (n_cases, 3)).astype(f32); hi = (lo + rng.uniform(0, 1, (n_cases, 3))).astype(f32)
    p4 = np.concatenate([(lo + (hi - lo) * rng.uniform(0, 1, (n_cases, 3))).astype(f32), np.ones((n_cases, 1), f32)], 1)
    q4 = np.concatenate([rng.uniform(-1, 3, (n_cases, 3)).astype(f32), np.ones((n_cases, 1), f32)], 1)
    out.update(clip_min=lo, clip_max=hi, clip_p=p4, clip_q=q4)
    out["ref_clip_aabb"] = np.asarray([taa["clip_aabb"](lo[i], hi[i], p4[i], q4[i]) for i in range(n_cases)], f32)
    # Particle.glsl:14-16, :30-32: spatial-hash cell of a particle and its key (int32 wrap-around, unsigned modulo)
    ptcl = G.load_functions(os.path.join(REF, "Core/res/shaders/Particle.glsl"), ["GetCellKey", "GetParticleCellPos"])
    ppos = np.concatenate([rng.uniform(-300, 300, (200, 3)), rng.uniform(-2, 2, (56, 3))]).astype(f32)
    cells = np.array([ptcl["GetParticleCellPos"](G.namespace(position_radius=np.append(ppos[i], f32(0.1)))) for i in range(n_cases)], np.int32)
    counts = rng.choice([1, 2, 1000, 100000, 131072, 999983], n_cases).astype(np.uint32)
    out.update(ptcl_pos=ppos, ref_cell_pos=cells, cell_count=counts)
    out["ref_cell_key"] = np.array([int(ptcl["GetCellKey"](tuple(G.I32(int(c)) for c in cells[i]), G.U32(int(counts[i])))) for i in range(n_cases)], np.uint32)
    # GBuffer.frag:41-43: CalcVelocity (motion vector from the current and the previous clip-space position)
    gb = G.load_functions(os.path.join(REF, "Core/res/shaders/GBuffer.frag"), ["CalcVelocity"])
    cur = np.concatenate([rng.uniform(-3, 3, (n_cases, 3)), rng.uniform(0.1, 50, (n_cases, 1))], 1).astype(f32)
    old = (cur + rng.normal(scale=0.05, size=(n_cases, 4))).astype(f32)
    out.update(clip_cur=cur, clip_old=old)
    out["ref_velocity"] = np.asarray([gb["CalcVelocity"](cur[i], old[i]) for i in range(n_cases)], f32)
    # RayGen.rgen:123-137: tangent frame and uniform hemisphere sample of BounceDiffuse (PI comes from LightBuffers.glsl)
    rg = G.load_functions(os.path.join(REF, "Core/res/shaders/RayGen.rgen"), ["CreateCoordinateSystem", "UniformSampleHemisphere"], defines={"PI": env["PI"]})
    r12 = rng.uniform(0, 1, (n_cases, 2)).astype(f32)
    frames = [rg["CreateCoordinateSystem"](n[i]) for i in range(n_cases)]
    out.update(hemi_r=r12, ref_nt=np.asarray([f[0] for f in frames], f32), ref_nb=np.asarray([f[1] for f in frames], f32))
    out["ref_hemisphere"] = np.asarray([rg["UniformSampleHemisphere"](r12[i, 0], r12[i, 1]) for i in range(n_cases)], f32)
    # RayGen.rgen:215-239, :284-294: the streaming reservoir updates with the shader's own random numbers
    wenv = G.load_functions(os.path.join(REF, "Core/res/shaders/Util.glsl"), ["gold_noise"])
    G.load_functions(os.path.join(REF, "Core/res/shaders/RayGen.rgen"), ["InitPixelReservoir", "UpdatePointlightReservoir", "UpdateEmissiveReservoir"],
                     defines={"INVALID_RESERVOIR_IDX": G.U32(0xFFFFFFFF)}, env=wenv)
    n_runs, n_cand = 48, 32
    w_px = np.stack([rng.integers(1, 1920, n_runs), rng.integers(0, 1080, n_runs)], -1).astype(np.uint32)
    w_time = rng.uniform(0, 60000, n_runs).astype(f32)
    w_idx = rng.integers(0, 1024, (n_runs, n_cand)).astype(np.uint32)
    w_tri = rng.integers(0, 5000, (n_runs, n_cand)).astype(np.uint32)
    w_bary = rng.uniform(0, 0.5, (n_runs, n_cand, 2)).astype(f32)
    w_weight = (rng.uniform(0, 4, (n_runs, n_cand)) * (rng.uniform(0, 1, (n_runs, n_cand)) > 0.2)).astype(f32)  # some zero weights
    w_seed = rng.uniform(0, 1, (n_runs, n_cand)).astype(f32)
    res_pl, res_em, sel_pl, sel_em = [], [], [], []
    for r in range(n_runs):
        wenv["gl_LaunchIDEXT"] = np.array([w_px[r, 0], w_px[r, 1], 0], f32)
        wenv["common_ubo"] = G.namespace(time_elapsed=w_time[r])
        a = G.namespace(emissive_triangle=G.namespace()); wenv["InitPixelReservoir"](a)
        b = G.namespace(emissive_triangle=G.namespace()); wenv["InitPixelReservoir"](b)
        sp, se = [], []
        for i in range(n_cand):
            sp.append(bool(wenv["UpdatePointlightReservoir"](a, G.U32(int(w_idx[r, i])), w_weight[r, i], w_seed[r, i])))
            t = G.namespace(object_instance_idx=G.U32(int(w_idx[r, i])), triangle_idx=G.U32(int(w_tri[r, i])), bary_coords=w_bary[r, i])
            se.append(bool(wenv["UpdateEmissiveReservoir"](b, t, w_weight[r, i], w_seed[r, i])))
        res_pl.append([float(a.w_sum_pl), int(a.pl_idx), int(a.pl_m)])
        et = b.emissive_triangle
        res_em.append([float(b.w_sum_em), int(et.object_instance_idx), int(et.triangle_idx), int(b.m_em)] +
                      (list(map(float, et.bary_coords)) if hasattr(et, "bary_coords") else [0.0, 0.0]))
        sel_pl.append(sp); sel_em.append(se)
    out.update(wrs_px=w_px, wrs_time=w_time, wrs_idx=w_idx, wrs_tri=w_tri, wrs_bary=w_bary, wrs_weight=w_weight, wrs_seed=w_seed,
               ref_wrs_pl=np.array(res_pl, np.float64), ref_wrs_em=np.array(res_em, np.float64),
               ref_wrs_sel_pl=np.array(sel_pl, np.uint8), ref_wrs_sel_em=np.array(sel_em, np.uint8))
    # RayGen.rgen:300-309: the point-light loop of DoInitialWRS (32 candidates: light pick, p_hat = |CalcPointlight|,
    # weight = p_hat * N, streaming update seeded with sin(i * 20382997.34741)) on explicit surfaces and 16 random lights
    renv = G.load_functions(os.path.join(REF, "Core/res/shaders/Util.glsl"), ["gold_noise"])
    G.load_functions(os.path.join(REF, "Core/res/shaders/LightBuffers.glsl"), names[:5] + ["CalcPointlight"], env=renv)
    G.load_functions(os.path.join(REF, "Core/res/shaders/RayGen.rgen"), ["InitPixelReservoir", "UpdatePointlightReservoir"],
                     defines={"INVALID_RESERVOIR_IDX": G.U32(0xFFFFFFFF)}, env=renv)
    _, l0, l1 = G.load_block(os.path.join(REF, "Core/res/shaders/RayGen.rgen"), r"for \(uint i = 0u; i < LIGHT_RIS_ITERS; i\+\+\) \{", "RisPointlights",
                             "reservoir, v, f0, world_pos, n, roughness, metallic, albedo", prologue="float p_hat_pl = 0.0;", result="p_hat_pl",
                             env=renv, defines={"LIGHT_RIS_ITERS": G.U32(32)})
    assert (l0, l1) == (300, 309), (l0, l1)
    n_ris = 64
    ris_lights = np.zeros((16, 12), f32)
    ris_lights[:, 0:3] = rng.uniform(0, 30, (16, 3)); ris_lights[:, 3] = 1
    ris_lights[:, 4:7] = rng.uniform(-6, 6, (16, 3)); ris_lights[:, 7] = 1
    ris_lights[:, 8:11] = [0.05, 0.25, 0.5]
    renv["ssbo_light_data"] = G.namespace(num_pointlights=G.U32(16), pointlights=[
        G.namespace(colour=ris_lights[k, 0:4], pos=ris_lights[k, 4:8], atten_exp=ris_lights[k, 8], atten_linear=ris_lights[k, 9], atten_constant=ris_lights[k, 10])
        for k in range(16)])
    ris_px = np.stack([rng.integers(1, 1920, n_ris), rng.integers(0, 1080, n_ris)], -1).astype(np.uint32)
    ris_time = rng.uniform(0, 60000, n_ris).astype(f32)
    ris_res = []
    for r in range(n_ris):
        renv["gl_LaunchIDEXT"] = np.array([ris_px[r, 0], ris_px[r, 1], 0], f32)
        renv["common_ubo"] = G.namespace(time_elapsed=ris_time[r])
        res_ = G.namespace(emissive_triangle=G.namespace()); renv["InitPixelReservoir"](res_)
        p_hat = renv["RisPointlights"](res_, v[r], f0[r], world_pos[r], n[r], rough[r], metal[r], albedo[r])
        ris_res.append([float(res_.w_sum_pl), int(res_.pl_idx), int(res_.pl_m), float(p_hat)])
    out.update(ris_lights=ris_lights, ris_px=ris_px, ris_time=ris_time, ref_ris_pl=np.array(ris_res, np.float64))
    # RayGen.rgen:258-282: ComputeTriangleRadiance (shadow_test = false) on the emissive cubes of small_test_scene, read through
    # the shader's own buffer indexing (instances -> indices -> positions / normals -> model matrix -> material)
    sc = snk_scenes.small_test_scene(64, 48)
    flat = sc.flatten()
    tenv = G.load_functions(os.path.join(REF, "Core/res/shaders/LightBuffers.glsl"), names[:5])
    G.load_functions(os.path.join(REF, "Core/res/shaders/RayGen.rgen"), ["ComputeTriangleRadiance"], env=tenv)
    u32 = G.U32
    tenv["rt_instances"] = G.namespace(i=[G.namespace(transform_idx=u32(int(r_[0])), mesh_buffer_index_offset=u32(int(r_[1])), mesh_buffer_vertex_offset=u32(int(r_[2])),
                                                      material_idx=u32(int(r_[3])), num_mesh_indices=u32(int(r_[4]))) for r_ in flat["instances"]])
    tri_idx = flat["indices"].reshape(-1, 3)
    tenv["indices"] = G.namespace(i=[G.IVec(G.I32(int(c)) for c in t) for t in tri_idx])
    tenv["vert_positions"] = G.namespace(p=[row for row in flat["positions"]])
    tenv["vert_normals"] = G.namespace(n=[row for row in flat["normals"]])
    tenv["transforms"] = G.namespace(m=[G.Mat(np.asarray(t, f32).reshape(4, 4)) for t in flat["transforms"]])
    tenv["material_ubo"] = G.namespace(materials=[G.namespace(albedo=np.asarray(sc.materials["albedo"][k], f32)) for k in range(sc.materials.shape[0])])
    em = [int(e) for e in flat["emissive_idx"]]
    assert em, "small_test_scene has emissive instances"
    n_tr = 128
    tr_obj = np.array([em[k % len(em)] for k in range(n_tr)], np.uint32)
    tr_tri = np.array([rng.integers(0, int(flat["instances"][o][4]) // 3) for o in tr_obj], np.uint32)
    tr_bary = rng.uniform(0, 0.5, (n_tr, 2)).astype(f32)
    tr_pos = rng.uniform((-6, 0.2, -4), (6, 5, 4), (n_tr, 3)).astype(f32)
    # surfaces that roughly face their light (centre of the cube = translation of its model matrix), so that most values are lit
    centres = np.array([np.asarray(flat["transforms"][int(flat["instances"][o][0])], f32).reshape(4, 4)[3, :3] for o in tr_obj], f32)
    tr_n = unit(centres - tr_pos + 0.6 * rng.normal(size=(n_tr, 3)))
    tr_v = unit(tr_n + 0.7 * rng.normal(size=(n_tr, 3)))
    ref_tr = []
    for k in range(n_tr):
        tdat = G.namespace(object_instance_idx=u32(int(tr_obj[k])), triangle_idx=u32(int(tr_tri[k])), bary_coords=tr_bary[k])
        ref_tr.append(tenv["ComputeTriangleRadiance"](tdat, tr_v[k], f0[k], tr_pos[k], tr_n[k], rough[k], metal[k], albedo[k], tr_pos[k], False))
    out.update(tri_obj=tr_obj, tri_tri=tr_tri, tri_bary=tr_bary, tri_surface_pos=tr_pos, tri_n=tr_n, tri_v=tr_v,
               ref_triangle_radiance=np.asarray(ref_tr, f32))
    # RayGen.rgen:311-342: the emissive-triangle loop of DoInitialWRS on the same scene: instance pick, triangle pick, random
    # barycentrics with the flip, area = det(mat3(M)) / 2 * |cross|, pdf, weight, streaming update seeded by sin(i * 38293.28347)
    G.load_functions(os.path.join(REF, "Core/res/shaders/Util.glsl"), ["gold_noise", "hash", "floatConstruct", "random"], env=tenv)
    G.load_functions(os.path.join(REF, "Core/res/shaders/RayGen.rgen"), ["InitPixelReservoir", "UpdateEmissiveReservoir"],
                     defines={"INVALID_RESERVOIR_IDX": G.U32(0xFFFFFFFF)}, env=tenv)
    _, e0, e1 = G.load_block(os.path.join(REF, "Core/res/shaders/RayGen.rgen"), r"for \(uint i = 0u; i < LIGHT_RIS_ITERS; i\+\+\) \{", "RisEmissives",
                             "reservoir, v, f0, world_pos, n, roughness, metallic, albedo, biased_pos", prologue="float p_hat_em = 0.0;", result="p_hat_em",
                             env=tenv, defines={"LIGHT_RIS_ITERS": G.U32(32)}, occurrence=2)
    assert (e0, e1) == (311, 342), (e0, e1)
    tenv["rt_emissive_instance_indices"] = G.namespace(num_emissive_instances=u32(len(em)), i=[u32(e) for e in em])
    n_em_runs = 48
    em_px = np.stack([rng.integers(1, 1920, n_em_runs), rng.integers(0, 1080, n_em_runs)], -1).astype(np.uint32)
    em_time = rng.uniform(0, 60000, n_em_runs).astype(f32)
    em_res = []
    for r in range(n_em_runs):
        tenv["gl_LaunchIDEXT"] = np.array([em_px[r, 0], em_px[r, 1], 0], f32)
        tenv["common_ubo"] = G.namespace(time_elapsed=em_time[r])
        res_ = G.namespace(emissive_triangle=G.namespace(bary_coords=np.zeros(2, f32))); tenv["InitPixelReservoir"](res_)
        p_hat = tenv["RisEmissives"](res_, tr_v[r], f0[r], tr_pos[r], tr_n[r], rough[r], metal[r], albedo[r], tr_pos[r])
        et = res_.emissive_triangle
        em_res.append([float(res_.w_sum_em), int(et.object_instance_idx), int(et.triangle_idx), int(res_.m_em), float(et.bary_coords[0]), float(et.bary_coords[1]), float(p_hat)])
    out.update(em_px=em_px, em_time=em_time, ref_ris_em=np.array(em_res, np.float64))
    # RayGen.rgen:359-386: CombinePixelReservoir with every shadow ray visible (ShadowTest stubbed to 1): the merge arithmetic of
    # temporal and spatial reuse — p_hat at the receiving surface, weight = p_hat * M * W (and p_hat * W * M), the seeds
    G.load_functions(os.path.join(REF, "Core/res/shaders/LightBuffers.glsl"), ["CalcPointlight"], env=tenv)
    tenv["ShadowTest"] = lambda ro, rd, max_dist: f32(1.0)
    G.load_functions(os.path.join(REF, "Core/res/shaders/RayGen.rgen"), ["UpdatePointlightReservoir", "CombinePixelReservoir"], env=tenv)
    spl = np.asarray(sc.pointlights, f32)
    tenv["ssbo_light_data"] = G.namespace(num_pointlights=u32(spl.shape[0]), pointlights=[
        G.namespace(colour=spl[k, 0:4], pos=spl[k, 4:8], atten_exp=spl[k, 8], atten_linear=spl[k, 9], atten_constant=spl[k, 10]) for k in range(spl.shape[0])])
    n_cb = 96
    INV = 0xFFFFFFFF
    cb_px = np.stack([rng.integers(1, 1920, n_cb), rng.integers(0, 1080, n_cb)], -1).astype(np.uint32)
    cb_time = rng.uniform(0, 60000, n_cb).astype(f32)
    cb_seed = rng.uniform(0, 1, n_cb).astype(f32)
    cb_in = np.zeros((n_cb, 2, 11), np.float64)  # per reservoir (current, neighbour): w_sum_pl w_pl pl_idx pl_m obj tri bx by w_sum_em w_em m_em
    cb_phat = rng.uniform(0, 0.5, (n_cb, 2)).astype(f32)
    cb_out, cb_phat_out = [], []
    for k in range(n_cb):
        rs = []
        for which in range(2):
            has_pl, has_em = rng.uniform() < 0.85, rng.uniform() < 0.85
            o = em[int(rng.integers(0, len(em)))]
            vals = [f32(rng.uniform(0, 3)), f32(rng.uniform(0, 2)), int(rng.integers(0, spl.shape[0])) if has_pl else INV, int(rng.integers(1, 40)),
                    o if has_em else INV, int(rng.integers(0, int(flat["instances"][o][4]) // 3)) if has_em else INV, f32(rng.uniform(0, 0.5)), f32(rng.uniform(0, 0.5)),
                    f32(rng.uniform(0, 0.2)), f32(rng.uniform(0, 30)), int(rng.integers(1, 40))]
            cb_in[k, which] = vals
            rs.append(G.namespace(w_sum_pl=f32(vals[0]), w_pl=f32(vals[1]), pl_idx=u32(vals[2]), pl_m=u32(vals[3]),
                                  emissive_triangle=G.namespace(object_instance_idx=u32(vals[4]), triangle_idx=u32(vals[5]), bary_coords=np.array(vals[6:8], f32)),
                                  w_sum_em=f32(vals[8]), w_em=f32(vals[9]), m_em=u32(vals[10])))
        tenv["gl_LaunchIDEXT"] = np.array([cb_px[k, 0], cb_px[k, 1], 0], f32)
        tenv["common_ubo"] = G.namespace(time_elapsed=cb_time[k])
        j = k % n_tr
        ph = tenv["CombinePixelReservoir"](rs[0], rs[1], cb_phat[k, 0], cb_phat[k, 1], tr_v[j], f0[j], tr_pos[j], tr_n[j], rough[j], metal[j], albedo[j], tr_pos[j], cb_seed[k])
        c, et = rs[0], rs[0].emissive_triangle
        cb_out.append([float(c.w_sum_pl), float(c.w_pl), int(c.pl_idx), int(c.pl_m), int(et.object_instance_idx), int(et.triangle_idx),
                       float(et.bary_coords[0]), float(et.bary_coords[1]), float(c.w_sum_em), float(c.w_em), int(c.m_em)])
        cb_phat_out.append([float(ph[0]), float(ph[1])])
    out.update(cb_px=cb_px, cb_time=cb_time, cb_seed=cb_seed, cb_in=cb_in, cb_phat_in=cb_phat, ref_cb_out=np.array(cb_out, np.float64),
               ref_cb_phat=np.array(cb_phat_out, np.float64))
    # RayGen.rgen:508-528: CalculateRadianceFromReservoirs (final shading of the two reservoir samples), visibility stubbed
    tenv["textureSize"] = lambda sampler, lod: G.IVec((G.I32(64), G.I32(48)))
    tenv["sampler_pixel_motion"] = None
    G.load_functions(os.path.join(REF, "Core/res/shaders/RayGen.rgen"), ["CalculateRadianceFromReservoirs"], env=tenv)
    ref_rad = []
    for k in range(n_cb):
        vals = cb_in[k, 1]
        rsv = G.namespace(w_sum_pl=f32(vals[0]), w_pl=f32(vals[1]), pl_idx=u32(int(vals[2])), pl_m=u32(int(vals[3])),
                          emissive_triangle=G.namespace(object_instance_idx=u32(int(vals[4])), triangle_idx=u32(int(vals[5])), bary_coords=np.array(vals[6:8], f32)),
                          w_sum_em=f32(vals[8]), w_em=f32(vals[9]), m_em=u32(int(vals[10])))
        j = k % n_tr
        rma = np.array([rough[j], metal[j], 1.0], f32)
        ref_rad.append(tenv["CalculateRadianceFromReservoirs"](rsv, tr_v[j], f0[j], tr_pos[j], tr_n[j], rma, albedo[j], tr_pos[j]))
    out["ref_reservoir_radiance"] = np.asarray(ref_rad, f32)
    # RayGen.rgen:390-409: FindNeighbourPixelCoord (random offset within the radius, re-drawn until it lies on the screen)
    G.load_functions(os.path.join(REF, "Core/res/shaders/RayGen.rgen"), ["FindNeighbourPixelCoord"], env=tenv)
    n_fn = 128
    fn_px = np.stack([rng.integers(1, 64, n_fn), rng.integers(0, 48, n_fn)], -1).astype(np.uint32)
    fn_time = rng.uniform(0, 60000, n_fn).astype(f32)
    fn_seed = rng.uniform(-1, 1, (n_fn, 2)).astype(f32)
    fn_start = ((fn_px.astype(f32) + f32(0.5)) / np.array([64, 48], f32)).astype(f32)
    ref_fn = []
    for k in range(n_fn):
        tenv["gl_LaunchIDEXT"] = np.array([fn_px[k, 0], fn_px[k, 1], 0], f32)
        tenv["common_ubo"] = G.namespace(time_elapsed=fn_time[k])
        ref_fn.append(tenv["FindNeighbourPixelCoord"](fn_start[k], f32(30.0), fn_seed[k]))
    out.update(fn_px=fn_px, fn_time=fn_time, fn_seed=fn_seed, fn_start=fn_start, ref_find_neighbour=np.asarray(ref_fn, f32))
    # RayGen.rgen:344-356: the tail of DoInitialWRS — W of the two selected samples and the shadow rays it asks for, including the
    # point-light direction `pos - biased_pos / pl_dist` as written (operator precedence: not a unit vector)
    G.load_functions(os.path.join(REF, "Core/res/shaders/RayGen.rgen"), ["TriangleShadowTest"], env=tenv)
    G.load_lines(os.path.join(REF, "Core/res/shaders/RayGen.rgen"), 344, 356, "if (reservoir.pl_idx != INVALID_RESERVOIR_IDX)", "WrsTail",
                 "reservoir, biased_pos, p_hat_pl, p_hat_em", env=tenv)
    asked = []
    tenv["ShadowTest"] = lambda ro, rd, max_dist: (asked.append(np.concatenate([ro, rd, [max_dist]]).astype(f32)), f32(1.0))[1]
    ref_tail, ref_rays, ref_nrays = [], [], []
    for k in range(n_cb):
        vals = cb_in[k, 0]
        rsv = G.namespace(w_sum_pl=f32(vals[0]), w_pl=f32(vals[1]), pl_idx=u32(int(vals[2])), pl_m=u32(int(vals[3])),
                          emissive_triangle=G.namespace(object_instance_idx=u32(int(vals[4])), triangle_idx=u32(int(vals[5])), bary_coords=np.array(vals[6:8], f32)),
                          w_sum_em=f32(vals[8]), w_em=f32(vals[9]), m_em=u32(int(vals[10])))
        del asked[:]
        tenv["WrsTail"](rsv, tr_pos[k % n_tr], cb_phat[k, 0] + f32(0.01), cb_phat[k, 1] + f32(0.01))
        ref_tail.append([float(rsv.w_pl), float(rsv.w_em)])
        ref_nrays.append(len(asked))
        ref_rays.append(np.concatenate(asked + [np.zeros(7, f32)] * (2 - len(asked))))
    tenv["ShadowTest"] = lambda ro, rd, max_dist: f32(1.0)
    out.update(ref_wrs_tail=np.array(ref_tail, f32), ref_wrs_tail_rays=np.array(ref_rays, f32), ref_wrs_tail_nrays=np.array(ref_nrays, np.uint32))
    # RayClosestHit.rchit:87-142: main() of the closest-hit shader for given hits (every instance of the scene, emissive ones included;
    # no albedo textures in this scene), the directional-light shadow ray recorded by the ShadowTest stub
    asked2 = []
    henv = dict(tenv)
    henv["ShadowTest"] = lambda ro, rd, max_dist: (asked2.append(np.concatenate([ro, rd, [max_dist]]).astype(f32)), f32(1.0))[1]
    G.load_functions(os.path.join(REF, "Core/res/shaders/LightBuffers.glsl"), names[:5] + ["CalcDirectionalLight"], env=henv)
    G.load_functions(os.path.join(REF, "Core/res/shaders/RayClosestHit.rchit"), ["main"], env=henv,
                     defines={"INVALID_GLOBAL_INDEX": u32(0xFFFFFFFF), "MAT_FLAG_EMISSIVE": u32(1)})
    henv["vert_tex_coords"] = G.namespace(t=[row for row in flat["uvs"]])
    henv["material_ubo"] = G.namespace(materials=[G.namespace(albedo=np.asarray(sc.materials["albedo"][k], f32), albedo_tex_idx=u32(int(sc.materials["tex"][k][0])),
                                                              flags=u32(int(sc.materials["flags"][k])), roughness=f32(sc.materials["roughness"][k]),
                                                              metallic=f32(sc.materials["metallic"][k])) for k in range(sc.materials.shape[0])])
    lights_buf = sc.flatten()["lights"]
    dl_col = np.frombuffer(lights_buf, f32, 4, 0).copy(); dl_dir = np.frombuffer(lights_buf, f32, 4, 16).copy()
    henv["ssbo_light_data"] = G.namespace(dir_light=G.namespace(colour=dl_col, dir=dl_dir))
    n_hit = 160
    n_inst = flat["instances"].shape[0]
    hit_inst = rng.integers(0, n_inst, n_hit).astype(np.uint32)
    hit_prim = np.array([rng.integers(0, int(flat["instances"][i][4]) // 3) for i in hit_inst], np.uint32)
    hit_uv = rng.uniform(0, 0.5, (n_hit, 2)).astype(f32)
    hit_t = rng.uniform(0.5, 30, n_hit).astype(f32)
    hit_rd = unit(rng.normal(size=(n_hit, 3)))
    ref_hit, ref_hit_ray, ref_hit_nray = [], [], []
    for k in range(n_hit):
        del asked2[:]
        henv["payload"] = G.namespace(colour=np.zeros(3, f32), normal=np.zeros(3, f32), distance=f32(-1.0), roughness=f32(0.0), mat_flags=u32(0))
        henv["attribs"] = hit_uv[k]
        henv["gl_InstanceCustomIndexEXT"] = u32(int(hit_inst[k])); henv["gl_PrimitiveID"] = u32(int(hit_prim[k]))
        henv["gl_RayTmaxEXT"] = hit_t[k]; henv["gl_WorldRayDirectionEXT"] = hit_rd[k]
        henv["main"]()
        pl_ = henv["payload"]
        ref_hit.append(np.concatenate([pl_.colour, pl_.normal, [pl_.distance, pl_.roughness, float(int(pl_.mat_flags))]]).astype(f32))
        ref_hit_nray.append(len(asked2))
        ref_hit_ray.append(asked2[0] if asked2 else np.zeros(7, f32))
    out.update(hit_inst=hit_inst, hit_prim=hit_prim, hit_uv=hit_uv, hit_t=hit_t, hit_rd=hit_rd, ref_hit_payload=np.array(ref_hit, f32),
               ref_hit_ray=np.array(ref_hit_ray, f32), ref_hit_nray=np.array(ref_hit_nray, np.uint32))
    np.savez_compressed(os.path.join(GOLD, "reference_shader_vectors.npz"), **out)


FRAME_SCENE = dict(width=40, height=28)   # small_test_scene: spheres + emissive cubes + point lights + directional light
FRAME_COUNT, FRAME_PIXELS = 5, 96


def frame_sequence(sc, n_frames=FRAME_COUNT):
    """the UBO sequence of the executed-frame pin (camera dolly; frame 0 has no history)"""
    prev = None
    for f in range(n_frames):
        pos = np.asarray(sc.camera.pos, np.float32) + np.asarray((0.05, 0.0, 0.1), np.float32) * np.float32(f)
        ubo = sc.ubo(f, cam_pos=pos)
        yield f, ubo, (ubo if prev is None else prev)
        prev = ubo


def reference_frame_vectors(n_pixels=FRAME_PIXELS, out_path=None):
    """main() of RayGen.rgen executed from the reference's shader text (tools/glsl_frame.py) for `n_pixels` pixels of each of
    FRAME_COUNT consecutive frames, both passes: the temporal reservoir WriteInitialReservoirs stores, the final reservoir
    ComputeAndWriteSpatialReservoirReuse stores and the colour main() stores (directional light + reservoir radiance +
    BounceDiffuse through the executed closest-hit / miss shaders) -> tests/golden/reference_frame_vectors.npz"""
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import glsl_frame as GF
    sc = snk_scenes.small_test_scene(**FRAME_SCENE)
    osc = O.OracleScene(sc)
    fr = O.OracleFrame(osc)
    ex = GF.FrameExec(sc, osc, shaders=os.path.join(REF, "Core/res/shaders"))
    st = O.default_settings(mode=0)
    rng = np.random.default_rng(20261019)
    px_all, temporal, final, colour, nrays = [], [], [], [], []
    for f, ubo, prev in frame_sequence(sc):
        fr.set(ubo, prev, st)
        fr.gbuffer()
        ex.set_frame(O, fr, ubo, st)
        ys, xs = np.nonzero(fr.read("prim") != 0xFFFFFFFF)
        sel = rng.choice(len(ys), n_pixels, replace=False)
        px = np.stack([xs[sel], ys[sel]], -1).astype(np.uint32)
        ex.load_reservoirs()          # history = the oracle's final reservoirs of the previous frame
        rows1, rays1 = [], []
        for x, y in px:
            row, _, rays = ex.run_pixel(int(x), int(y), 1)
            rows1.append(row if row is not None else [np.nan] * 11)   # emissive pixels return before the pass writes
            rays1.append([sum(r[0] == "closest" for r in rays), sum(r[0] == "any" for r in rays)])
        fr.render_pass(1)
        ex.load_reservoirs()          # spatial reuse reads the oracle's temporal half of THIS frame
        rows0, cols0, rays0 = [], [], []
        for x, y in px:
            row, col, rays = ex.run_pixel(int(x), int(y), 0)
            rows0.append(row if row is not None else [np.nan] * 11)
            cols0.append(col)
            rays0.append([sum(r[0] == "closest" for r in rays), sum(r[0] == "any" for r in rays)])
        fr.render_pass(0)
        px_all.append(px); temporal.append(rows1); final.append(rows0); colour.append(cols0); nrays.append(np.concatenate([rays1, rays0], 1))
    out = dict(px=np.array(px_all, np.uint32), ref_temporal=np.array(temporal, np.float64), ref_final=np.array(final, np.float64),
               ref_colour=np.array(colour, np.float32), ref_rays=np.array(nrays, np.uint32))
    if out_path is not False:
        np.savez_compressed(out_path or os.path.join(GOLD, "reference_frame_vectors.npz"), **out)
    return out


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "source_facts":
        reference_source_facts()
        reference_layouts()
        reference_shader_vectors()
        reference_frame_vectors()
        raise SystemExit(0)
    O.build_oracle()
    reference_source_facts()
    reference_layouts()
    reference_shader_vectors()
    reference_frame_vectors()
    glm_vectors()
    shipped_meshes()
    shipped_mesh_arrays()
    shipped_scene()
    trace_cornell()
    frame_small()
    for f in sorted(os.listdir(GOLD)):
        print(f, os.path.getsize(os.path.join(GOLD, f)))
