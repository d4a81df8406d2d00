"""ctypes bindings of the CPU oracle (oracle/libsnkoracle.so). Test infrastructure only:
importable from tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / reference arm."""
from __future__ import annotations

import ctypes as C
import importlib
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
scenes = importlib.import_module("snake-vk_b200.scenes")

_LIB = None


def build_oracle() -> None:
    subprocess.run(["make", "-C", os.path.join(ROOT, "oracle"), "-s"], check=True)


def lib() -> C.CDLL:
    global _LIB
    if _LIB is None:
        path = os.path.join(ROOT, "oracle", "libsnkoracle.so")
        if not os.path.exists(path):
            build_oracle()
        L = C.CDLL(path)
        L.snkor_scene_create.restype = C.c_void_p
        L.snkor_scene_create.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32, C.c_void_p,
                                         C.c_uint32, C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32, C.c_void_p,
                                         C.c_size_t, C.c_void_p, C.c_uint32]
        L.snkor_scene_destroy.argtypes = [C.c_void_p]
        L.snkor_trace.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint32, C.c_void_p, C.c_int]
        L.snkor_frame_create.restype = C.c_void_p
        L.snkor_frame_create.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32]
        L.snkor_frame_destroy.argtypes = [C.c_void_p]
        L.snkor_frame_set_scene.argtypes = [C.c_void_p, C.c_void_p]
        L.snkor_frame_set.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.snkor_frame_gbuffer.argtypes = [C.c_void_p, C.c_float, C.c_float, C.c_uint32, C.c_uint32, C.c_int]
        L.snkor_frame_render.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_int]
        L.snkor_frame_set_halo.argtypes = [C.c_void_p, C.c_uint32]
        L.snkor_frame_render_pass.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_int]
        L.snkor_frame_read.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_size_t]
        L.snkor_frame_write_reservoirs.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32, C.c_uint32]
        L.snkor_frame_ray_counts.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        L.snkor_inv_view_proj.argtypes = [C.c_void_p, C.c_void_p]
        L.snkor_hash.restype = C.c_uint32
        L.snkor_hash.argtypes = [C.c_uint32]
        L.snkor_random.restype = C.c_float
        L.snkor_random.argtypes = [C.c_float]
        L.snkor_gold_noise.restype = C.c_float
        L.snkor_gold_noise.argtypes = [C.c_float, C.c_float, C.c_float]
        L.snkor_f32_to_f16.restype = C.c_uint16
        L.snkor_f32_to_f16.argtypes = [C.c_float]
        L.snkor_f16_to_f32.restype = C.c_float
        L.snkor_f16_to_f32.argtypes = [C.c_uint16]
        L.snkor_f32_to_snorm8.restype = C.c_int
        L.snkor_f32_to_snorm8.argtypes = [C.c_float]
        L.snkor_brdf_eval.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_float, C.c_float, C.c_void_p]
        L.snkor_calc_pointlight.argtypes = [C.c_void_p] * 5 + [C.c_float, C.c_float, C.c_void_p, C.c_void_p]
        L.snkor_calc_dirlight.argtypes = [C.c_void_p] * 4 + [C.c_float, C.c_float, C.c_void_p, C.c_void_p]
        L.snkor_wrs_pointlight_script.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p]
        L.snkor_pixel_scripted.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_void_p, C.c_uint32]
        L.snkor_sizes.argtypes = [C.c_void_p]
        L.snkor_hardware_threads.restype = C.c_int
        L.snkor_taa_resolve.argtypes = [C.c_uint32, C.c_uint32] + [C.c_void_p] * 5
        L.snkor_taa_resolve.restype = None
        L.snkor_particles_init.argtypes = [C.c_void_p, C.c_uint32]
        L.snkor_particles_init.restype = None
        L.snkor_particles_update.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_int, C.c_int]
        L.snkor_particles_update.restype = C.c_uint64
        L.snkor_particles_grid.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p]
        L.snkor_particles_grid.restype = None
        L.snkor_scene_set_texture.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_void_p, C.c_uint32]
        L.snkor_scene_set_texture.restype = None
        L.snkor_scene_set_tangents.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32]
        L.snkor_scene_set_tangents.restype = None
        L.snkor_sample_texture.argtypes = [C.c_void_p, C.c_uint32, C.c_float, C.c_float, C.c_void_p]
        L.snkor_sample_texture.restype = None
        _LIB = L
    return _LIB


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


AUX = dict(prim=(0, np.uint32, 1), inst=(1, np.uint32, 1), depth=(2, np.float32, 1), normal=(3, np.float16, 4),
           albedo=(4, np.int8, 4), rma=(5, np.float16, 4), motion=(6, np.float16, 2), flags=(7, np.uint16, 1),
           prev_normal=(8, np.float16, 4), prev_depth=(9, np.float32, 1), reservoirs_final=(10, scenes.RESERVOIR_DTYPE, 1),
           reservoirs_temporal=(11, scenes.RESERVOIR_DTYPE, 1), hit_t=(14, np.float32, 1), bary=(15, np.float32, 2),
           radiance=(16, np.float32, 4), direct=(18, np.float32, 4))

SETTINGS_DTYPE = np.dtype([("num_temporal", "<i4"), ("num_spatial", "<i4"), ("mode", "<u4"), ("ris_iters", "<u4"),
                           ("m_cap", "<u4"), ("neighbour_radius", "<f4"), ("normal_threshold", "<f4"),
                           ("max_bounces", "<u4"), ("spp", "<u4"), ("flags", "<u4"), ("reserved", "<u4", (6,))])
assert SETTINGS_DTYPE.itemsize == 64


def default_settings(**kw) -> np.ndarray:
    s = np.zeros((), SETTINGS_DTYPE)
    s["num_temporal"], s["num_spatial"], s["ris_iters"], s["m_cap"] = 1, 3, 32, 20
    s["neighbour_radius"], s["normal_threshold"], s["max_bounces"], s["spp"] = 30.0, 0.95, 4, 1
    for k, v in kw.items():
        s[k] = v
    return s


class OracleScene:
    def __init__(self, scene: "scenes.Scene"):
        self.scene = scene
        self.flat = f = scene.flatten()
        self._keep = f
        lights = np.frombuffer(f["lights"], np.uint8)
        self.h = lib().snkor_scene_create(
            _p(f["positions"]), _p(f["normals"]), _p(f["uvs"]), f["positions"].shape[0], _p(f["indices"]), f["indices"].size,
            _p(f["instances"]), f["instances"].shape[0], _p(f["transforms"]), _p(f["prev_transforms"]), f["transforms"].shape[0],
            _p(f["materials"]), f["materials"].shape[0], _p(lights), lights.size, _p(f["emissive_idx"]), f["emissive_idx"].size)
        for i, (texels, srgb) in enumerate(getattr(scene, "textures", [])):
            t = np.ascontiguousarray(texels, np.uint8)
            lib().snkor_scene_set_texture(self.h, i, t.shape[1], t.shape[0], _p(t), int(bool(srgb)))
        if "tangents" in f:
            lib().snkor_scene_set_tangents(self.h, _p(f["tangents"]), f["tangents"].shape[0])

    def sample_texture(self, idx: int, u: float, v: float) -> np.ndarray:
        out = np.zeros(3, np.float32)
        lib().snkor_sample_texture(self.h, idx, u, v, _p(out))
        return out

    def trace(self, rays: np.ndarray, any_hit=False, brute=False, threads=8) -> np.ndarray:
        rays = np.ascontiguousarray(rays)
        assert rays.dtype == scenes.RAY_DTYPE
        hits = np.zeros(rays.shape[0], scenes.HIT_DTYPE)
        lib().snkor_trace(self.h, _p(rays), rays.shape[0], (1 if any_hit else 0) | (2 if brute else 0), _p(hits), threads)
        return hits

    def __del__(self):
        if getattr(self, "h", None):
            lib().snkor_scene_destroy(self.h)
            self.h = None


class OracleFrame:
    def __init__(self, osc: OracleScene, width=None, height=None):
        self.osc = osc
        self.w = width or osc.scene.width
        self.h_ = height or osc.scene.height
        self.h = lib().snkor_frame_create(osc.h, self.w, self.h_)

    def set(self, cur: np.ndarray, prev: np.ndarray, settings: np.ndarray | None = None):
        self._cur, self._prev, self._st = cur.copy(), prev.copy(), None if settings is None else settings.copy()
        lib().snkor_frame_set(self.h, _p(self._cur), _p(self._prev), _p(self._st) if self._st is not None else None)

    def gbuffer(self, jitter=(0.0, 0.0), rows=(0, 0), threads=8):
        lib().snkor_frame_gbuffer(self.h, jitter[0], jitter[1], rows[0], rows[1], threads)

    def render(self, rows=(0, 0), threads=8):
        lib().snkor_frame_render(self.h, rows[0], rows[1], threads)

    def set_halo(self, halo: int):
        """north-star mode: rows around the rendered band that also run the initial + temporal reservoir stages"""
        lib().snkor_frame_set_halo(self.h, int(halo))

    def render_pass(self, pass_: int, rows=(0, 0), threads=8):
        lib().snkor_frame_render_pass(self.h, pass_, rows[0], rows[1], threads)

    def read(self, name: str) -> np.ndarray:
        what, dt, ch = AUX[name]
        out = np.zeros((self.h_, self.w, ch) if ch > 1 else (self.h_, self.w), dt)
        rc = lib().snkor_frame_read(self.h, what, _p(out), out.nbytes)
        assert rc == 0, name
        return out

    def write_reservoirs(self, half: int, res: np.ndarray, first=0):
        res = np.ascontiguousarray(res.reshape(-1))
        assert lib().snkor_frame_write_reservoirs(self.h, half, _p(res), first, res.size) == 0

    def ray_counts(self, reset=False):
        a, b = C.c_uint64(), C.c_uint64()
        lib().snkor_frame_ray_counts(self.h, C.byref(a), C.byref(b), int(reset))
        return a.value, b.value

    def pixel_scripted(self, x, y, pass_, u=None):
        u = np.zeros(0, np.float32) if u is None else np.ascontiguousarray(u, np.float32)
        lib().snkor_pixel_scripted(self.h, x, y, pass_, _p(u) if u.size else None, u.size)

    def __del__(self):
        if getattr(self, "h", None):
            lib().snkor_frame_destroy(self.h)
            self.h = None


PARTICLE_DTYPE = np.dtype([("position_radius", "<f4", (4,)), ("velocity", "<f4", (4,)), ("colour", "<f4", (4,))])


def particles_init(n: int) -> np.ndarray:
    """Particle.comp Init (oracle/snk_particles.h)"""
    p = np.zeros(n, PARTICLE_DTYPE)
    lib().snkor_particles_init(_p(p), n)
    return p


def particles_update(osc: "OracleScene", particles: np.ndarray, params: np.ndarray | None = None, substeps=3, threads=8):
    """ParticleSystem::UpdateParticles on the CPU; returns (particles after the sub-steps, rays traced)"""
    p = np.ascontiguousarray(particles).copy()
    prm = None if params is None else np.ascontiguousarray(params)
    rays = lib().snkor_particles_update(osc.h, _p(p), p.size, _p(prm), substeps, threads)
    return p, int(rays)


def particles_grid(particles: np.ndarray):
    p = np.ascontiguousarray(particles)
    keys = np.zeros((p.size, 2), np.uint32)
    start = np.zeros(p.size, np.uint32)
    lib().snkor_particles_grid(_p(p), p.size, _p(keys), _p(start))
    return keys, start


def taa_resolve(cur: np.ndarray, velocity: np.ndarray, history: np.ndarray, prev_velocity: np.ndarray) -> np.ndarray:
    """TAA_Resolve.comp (oracle/snk_taa.h) on (H, W, 4) / (H, W, 2) float16 images; returns the resolved frame."""
    h, w = cur.shape[:2]
    cur, history = np.ascontiguousarray(cur, np.float16), np.ascontiguousarray(history, np.float16)
    velocity, prev_velocity = np.ascontiguousarray(velocity, np.float16), np.ascontiguousarray(prev_velocity, np.float16)
    assert cur.shape == history.shape == (h, w, 4) and velocity.shape == prev_velocity.shape == (h, w, 2)
    out = np.zeros_like(cur)
    lib().snkor_taa_resolve(w, h, _p(cur), _p(velocity), _p(history), _p(prev_velocity), _p(out))
    return out


def camera_rays(scene, ubo, width=None, height=None) -> np.ndarray:
    """primary rays as the G-buffer pass shoots them (for ray-level tests); float64 math, rounded."""
    w, h = width or scene.width, height or scene.height
    ivp = np.zeros(16, np.float32)
    lib().snkor_inv_view_proj(_p(ubo), _p(ivp))
    m = ivp.reshape(4, 4).astype(np.float64)  # [col][row]
    xs = (np.arange(w) + 0.5) / w * 2 - 1
    ys = (np.arange(h) + 0.5) / h * 2 - 1
    gx, gy = np.meshgrid(xs, ys)
    clip = np.stack([gx, gy, np.ones_like(gx), np.ones_like(gx)], -1)
    ws = clip @ m  # row-vector times [col][row] = M * v
    far = ws[..., :3] / ws[..., 3:4]
    cam = ubo["cam_pos"][:3].astype(np.float64)
    d = far - cam
    d /= np.linalg.norm(d, axis=-1, keepdims=True)
    rays = np.zeros(w * h, scenes.RAY_DTYPE)
    rays["o"] = cam.astype(np.float32)
    rays["d"] = d.reshape(-1, 3).astype(np.float32)
    rays["tmin"], rays["tmax"] = 0.1, 10000.0
    return rays
