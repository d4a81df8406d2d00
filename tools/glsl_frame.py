"""Executes main() of the reference's ray-generation shader for single pixels of a frame — both ReSTIR passes
(RayGen.rgen:411-468 WriteInitialReservoirs, :470-506 ComputeAndWriteSpatialReservoirReuse), BounceDiffuse (:139-213) and
main() itself (:530-570) — together with main() of RayClosestHit.rchit (:87-142) and of the two miss shaders for every ray
the pixel traces, all translated from the shader text where it lies under /root/reference by tools/glsl_exec.py.

What is NOT executed (the driver's part of the frame) is stubbed from the CPU oracle's state of the same frame:
  texelFetch / textureSize      the oracle's G-buffer planes, decoded from their storage formats
  texture(prev normals, uv)     bilinear, repeat addressing, float32 weights (filter precision is implementation-defined)
  reservoirs.r[]                the oracle's reservoir buffer (both halves) as it is BEFORE the pass that is executed
  traceRayEXT                   the oracle's brute-force intersector; a closest hit runs the executed closest-hit shader,
                                a miss the executed miss shader
  WorldPosFromDepth             the repository's contract: inverse(view) * inverse(proj) hoisted, float64, rounded once
                                (GLSL leaves the precision of inverse() to the implementation)
So each pass of each frame is pinned one step at a time: same inputs, the reference's own per-pixel program on one side,
oracle/snk_raygen.h on the other. Only numeric vectors are committed (tools/make_golden.py: reference_frame_vectors).
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import glsl_exec as G  # noqa: E402

F = np.float32
U32 = G.U32
REF_SHADERS = "/root/reference/Core/res/shaders"
INVALID = 0xFFFFFFFF


class FlexList(list):
    """buffer array indexed by whatever integer-valued number the shader arithmetic produced"""
    def __getitem__(self, i): return list.__getitem__(self, int(i))
    def __setitem__(self, i, v): list.__setitem__(self, int(i), v)


class Sampler:
    def __init__(self, texels, integer=False):
        self.t, self.integer = texels, integer  # (H, W, C) float32, or uint for the material-flags plane

    def fetch(self, x, y):
        v = self.t[int(y), int(x)]
        return [U32(int(c)) for c in v] if self.integer else np.asarray(v, F).copy()


def reservoir_to_ns(rec):
    return G.namespace(w_sum_pl=F(rec["w_sum_pl"]), w_pl=F(rec["w_pl"]), pl_idx=U32(int(rec["pl_idx"])), pl_m=U32(int(rec["pl_m"])),
                       emissive_triangle=G.namespace(object_instance_idx=U32(int(rec["object_instance_idx"])), triangle_idx=U32(int(rec["triangle_idx"])),
                                                     bary_coords=np.asarray(rec["bary"], F).copy()),
                       w_sum_em=F(rec["w_sum_em"]), w_em=F(rec["w_em"]), m_em=U32(int(rec["m_em"])))


def ns_to_row(r):
    """11 numbers of a reservoir (float64 holds every field exactly): w_sum_pl w_pl pl_idx pl_m obj tri bx by w_sum_em w_em m_em"""
    et = r.emissive_triangle
    bary = et.bary_coords if getattr(et, "bary_coords", None) is not None else (0.0, 0.0)
    def num(v, default=0.0):
        return default if v is None else float(v)
    return [num(r.w_sum_pl), num(getattr(r, "w_pl", None)), float(int(r.pl_idx)), float(int(r.pl_m)), float(int(et.object_instance_idx)),
            float(int(et.triangle_idx)), float(bary[0]), float(bary[1]), num(r.w_sum_em), num(getattr(r, "w_em", None)), float(int(r.m_em))]


class FrameExec:
    """the reference's per-pixel ray-generation program over one scene; set_frame() binds an oracle frame, run_pixel() executes"""

    def __init__(self, scene, oracle_scene, shaders=REF_SHADERS):
        self.sc, self.osc = scene, oracle_scene
        flat = self.flat = scene.flatten()
        lb = os.path.join(shaders, "LightBuffers.glsl")
        brdf = ["FresnelSchlick", "DistributionGGX", "GeometrySchlickGGX", "GeometrySmith", "CookTorranceBRDF", "CalcDirectionalLight", "CalcPointlight"]
        consts = {"INVALID_RESERVOIR_IDX": U32(INVALID), "INVALID_GLOBAL_INDEX": U32(INVALID), "MAT_FLAG_EMISSIVE": U32(1), "LIGHT_RIS_ITERS": U32(32),
                  "gl_RayFlagsTerminateOnFirstHitEXT": 4, "gl_RayFlagsOpaqueEXT": 1, "gl_RayFlagsSkipClosestHitShaderEXT": 8, "accel": None,
                  "img_output": None}
        # ---- scene buffers, shared by every shader stage
        u32 = U32
        bufs = {}
        bufs["rt_instances"] = G.namespace(i=FlexList(G.namespace(transform_idx=u32(int(r[0])), mesh_buffer_index_offset=u32(int(r[1])), mesh_buffer_vertex_offset=u32(int(r[2])),
                                                                  material_idx=u32(int(r[3])), num_mesh_indices=u32(int(r[4]))) for r in flat["instances"]))
        bufs["indices"] = G.namespace(i=FlexList(G.IVec(G.I32(int(c)) for c in t) for t in flat["indices"].reshape(-1, 3)))
        bufs["vert_positions"] = G.namespace(p=FlexList(np.asarray(r, F) for r in flat["positions"]))
        bufs["vert_normals"] = G.namespace(n=FlexList(np.asarray(r, F) for r in flat["normals"]))
        bufs["vert_tex_coords"] = G.namespace(t=FlexList(np.asarray(r, F) for r in flat["uvs"]))
        bufs["transforms"] = G.namespace(m=FlexList(G.Mat(np.asarray(t, F).reshape(4, 4)) for t in flat["transforms"]))
        mats = scene.materials
        bufs["material_ubo"] = G.namespace(materials=FlexList(G.namespace(albedo=np.asarray(mats["albedo"][k], F), albedo_tex_idx=u32(int(mats["tex"][k][0])),
                                                                          flags=u32(int(mats["flags"][k])), roughness=F(mats["roughness"][k]), metallic=F(mats["metallic"][k]))
                                                              for k in range(mats.shape[0])))
        em = [int(e) for e in flat["emissive_idx"]]
        bufs["rt_emissive_instance_indices"] = G.namespace(num_emissive_instances=u32(len(em)), i=FlexList(u32(e) for e in em))
        lights = flat["lights"]
        pl = np.asarray(scene.pointlights, F)
        bufs["ssbo_light_data"] = G.namespace(
            dir_light=G.namespace(colour=np.frombuffer(lights, F, 4, 0).copy(), dir=np.frombuffer(lights, F, 4, 16).copy()),
            num_pointlights=u32(pl.shape[0]),
            pointlights=FlexList(G.namespace(colour=pl[k, 0:4], pos=pl[k, 4:8], atten_exp=pl[k, 8], atten_linear=pl[k, 9], atten_constant=pl[k, 10]) for k in range(pl.shape[0])))
        self.payload = G.namespace(colour=np.zeros(3, F), normal=np.zeros(3, F), distance=F(-1.0), roughness=F(0.0), mat_flags=u32(0))
        self.shadow_payload = G.namespace(hit=False)
        shared = dict(bufs, payload=self.payload, shadow_payload=self.shadow_payload, traceRayEXT=self.trace_ray)
        # ---- ray generation
        env = self.rgen = G.load_functions(os.path.join(shaders, "Util.glsl"), ["gold_noise", "hash", "floatConstruct", "random"])
        G.load_functions(lb, brdf, env=env)
        env.update(consts); env.update(shared)
        env.update(WorldPosFromDepth=self.world_pos_from_depth, texelFetch=lambda s, c, lod: s.fetch(c[0], c[1]), texture=self.texture,
                   textureSize=lambda s, lod: G.IVec((G.I32(self.W), G.I32(self.H))), imageStore=self.image_store)
        G.load_functions(os.path.join(shaders, "RayGen.rgen"),
                         ["ShadowTest", "CreateCoordinateSystem", "UniformSampleHemisphere", "BounceDiffuse", "UpdatePointlightReservoir", "UpdateEmissiveReservoir",
                          "TriangleShadowTest", "ComputeTriangleRadiance", "InitPixelReservoir", "DoInitialWRS", "CombinePixelReservoir", "FindNeighbourPixelCoord",
                          "WriteInitialReservoirs", "ComputeAndWriteSpatialReservoirReuse", "CalculateRadianceFromReservoirs", "main"], env=env, defines=consts)
        # ---- closest hit + miss shaders (own namespaces: each stage has its own main / ShadowTest)
        self.rchit = G.load_functions(lb, brdf)
        self.rchit.update(consts); self.rchit.update(shared)
        G.load_functions(os.path.join(shaders, "RayClosestHit.rchit"), ["ShadowTest", "main"], env=self.rchit, defines=consts)
        self.rmiss = G.load_functions(os.path.join(shaders, "RayMiss.rmiss"), ["main"], env=dict(G.BUILTINS, payload=self.payload))
        self.smiss = G.load_functions(os.path.join(shaders, "Shadow.rmiss"), ["main"], env=dict(G.BUILTINS, shadow_payload=self.shadow_payload))
        self.rays = []  # (kind, o, tmin, d, tmax) of the pixel being executed
        self.W = self.H = 0

    # ---- stubs -----------------------------------------------------------------------------------------------------
    def world_pos_from_depth(self, depth, uv):
        m = self.ivp  # column-major, m[4 * col + row]
        p = [F(F(uv[0]) * F(2.0) - F(1.0)), F(F(uv[1]) * F(2.0) - F(1.0)), F(depth), F(1.0)]
        r = [F(F(F(m[k] * p[0] + m[4 + k] * p[1]) + m[8 + k] * p[2]) + m[12 + k] * p[3]) for k in range(4)]
        w = r[3] if r[3] > F(1e-6) else F(1e-6)
        return np.array([r[0] / w, r[1] / w, r[2] / w], F)

    def texture(self, sampler, uv):
        assert sampler is self.rgen["sampler_prev_frame_normals"], "only the previous-frame normals are sampled with a filter"
        W, H = self.W, self.H
        su, sv = F(F(uv[0]) * F(W) - F(0.5)), F(F(uv[1]) * F(H) - F(0.5))
        fu, fv = np.floor(su), np.floor(sv)
        a, b = F(su - fu), F(sv - fv)
        x0, y0 = int(fu), int(fv)
        x1, y1 = x0 + 1, y0 + 1
        x0, x1, y0, y1 = x0 % W, x1 % W, y0 % H, y1 % H
        t = sampler.t
        top = (t[y0, x0] * F(F(1.0) - a) + t[y0, x1] * a).astype(F)
        bot = (t[y1, x0] * F(F(1.0) - a) + t[y1, x1] * a).astype(F)
        return (top * F(F(1.0) - b) + bot * b).astype(F)

    def image_store(self, image, coord, value):
        self.stored = np.asarray(value, F).copy()

    def trace_ray(self, accel, flags, mask, sbt_offset, sbt_stride, miss_index, ro, tmin, rd, tmax, payload_location):
        O = self.O
        ray = np.zeros(1, O.scenes.RAY_DTYPE)
        ray["o"], ray["tmin"], ray["d"], ray["tmax"] = np.asarray(ro, F), F(tmin), np.asarray(rd, F), F(tmax)
        any_hit = bool(int(flags) & 4)
        self.rays.append(("any" if any_hit else "closest",) + tuple(float(c) for c in ray["o"][0]) + (float(tmin),) + tuple(float(c) for c in ray["d"][0]) + (float(tmax),))
        hit = self.osc.trace(ray, any_hit=any_hit, brute=True, threads=1)[0]
        if any_hit:
            assert int(payload_location) == 1 and int(miss_index) == 1 and int(flags) & 8
            if hit["t"] < 0:
                self.smiss["main"]()  # Shadow.rmiss: hit = false
            return
        assert int(payload_location) == 0 and int(miss_index) == 0
        if hit["t"] < 0:
            self.rmiss["gl_WorldRayDirectionEXT"] = np.asarray(rd, F)
            self.rmiss["main"]()
            return
        e = self.rchit
        e["attribs"] = np.array([hit["u"], hit["v"]], F)
        e["gl_InstanceCustomIndexEXT"] = U32(int(hit["inst"])); e["gl_PrimitiveID"] = U32(int(hit["prim"]))
        e["gl_RayTmaxEXT"] = F(hit["t"]); e["gl_WorldRayDirectionEXT"] = np.asarray(rd, F)
        e["main"]()

    # ---- frame binding ---------------------------------------------------------------------------------------------
    def set_frame(self, O, fr, ubo, settings):
        """binds the oracle frame `fr` (after fr.set + fr.gbuffer): G-buffer planes now, the reservoir buffer at every run_pixel"""
        self.O, self.fr = O, fr
        self.W, self.H = fr.w, fr.h_
        ivp = np.zeros(16, F)
        O.lib().snkor_inv_view_proj(O._p(ubo), O._p(ivp))
        self.ivp = ivp
        env = self.rgen
        alb = fr.read("albedo").astype(F) / F(127.0)
        env["sampler_gbuffer_albedo"] = Sampler(np.maximum(alb, F(-1.0)).astype(F))
        env["sampler_gbuffer_normals"] = Sampler(fr.read("normal").astype(F))
        env["sampler_gbuffer_depth"] = Sampler(fr.read("depth").astype(F)[..., None])
        env["sampler_gbuffer_rma"] = Sampler(fr.read("rma").astype(F))
        env["sampler_gbuffer_mat_flags"] = Sampler(fr.read("flags").astype(np.uint32)[..., None], integer=True)
        env["sampler_pixel_motion"] = Sampler(fr.read("motion").astype(F))
        env["sampler_prev_frame_normals"] = Sampler(fr.read("prev_normal").astype(F))
        env["common_ubo"] = G.namespace(cam_pos=np.asarray(ubo["cam_pos"], F).copy(), time_elapsed=F(ubo["time_elapsed"]), delta_time=F(ubo["delta_time"]),
                                        prev_frame_invalid=U32(int(ubo["prev_frame_invalid"])))
        self.rchit["common_ubo"] = env["common_ubo"]
        self.settings = settings
        env["gl_LaunchSizeEXT"] = np.array([self.W, self.H, 1], F)

    def load_reservoirs(self):
        """the oracle's reservoir buffer as it is now: [final half | temporal half]"""
        fin, tmp = self.fr.read("reservoirs_final").reshape(-1), self.fr.read("reservoirs_temporal").reshape(-1)
        self.res_records = (fin, tmp)
        buf = FlexList([None] * (2 * self.W * self.H))
        self.rgen["reservoirs"] = [LazyReservoirs(buf, fin, tmp)]  # `reservoirs.r[...]`: .r is rewritten to [0] like a swizzle
        return buf

    def run_pixel(self, x, y, pass_):
        """executes main() for pixel (x, y) with push.initial_reservoir_pass = pass_; returns the reservoir it wrote
        (11 numbers), the stored colour (pass 0) and the rays it traced"""
        env = self.rgen
        env["gl_LaunchIDEXT"] = np.array([x, y, 0], F)
        env["push"] = G.namespace(initial_reservoir_pass=int(pass_), num_temporal_resamples=int(self.settings["num_temporal"]),
                                  num_spatial_resamples=int(self.settings["num_spatial"]))
        self.rays = []
        self.stored = None
        lazy = env["reservoirs"][0]
        env["main"]()
        idx = (self.W * self.H if pass_ == 1 else 0) + y * self.W + x
        written = lazy.written.get(idx)
        return (ns_to_row(written) if written is not None else None), self.stored, list(self.rays)


class LazyReservoirs:
    """reservoirs.r[]: records are converted on first read; writes are kept apart (the executed pixel must not disturb its
    neighbours' inputs, exactly like the two halves of the real buffer)"""

    def __init__(self, buf, fin, tmp):
        self.fin, self.tmp, self.n = fin, tmp, fin.shape[0]
        self.written = {}

    def __getitem__(self, i):
        i = int(i)
        if i in self.written:
            return self.written[i]
        rec = self.fin[i] if i < self.n else self.tmp[i - self.n]
        return reservoir_to_ns(rec)

    def __setitem__(self, i, v):
        self.written[int(i)] = v
