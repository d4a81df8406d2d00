"""ctypes host mirror of the C ABI (include/snkpt.h).

`PathTracer` is a 1:1 wrapper of the entry points; `RT` mirrors the reference's
class (Core/headers/rendering/Raytracing.h:11-31: Init / RecordRenderCmdBuf /
RtSettings) on top of it so the parity tests read like engine code.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import scenes

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


class SnkError(RuntimeError):
    pass


def lib_path() -> str:
    # SNKPT_LIB_VARIANT selects a tuning build (snake-vk_b200/Makefile VARIANT=...); default = the product library
    return os.path.join(_HERE, f"libsnkpt{os.environ.get('SNKPT_LIB_VARIANT', '')}.so")


def load_library() -> C.CDLL:
    """dlopen libsnkpt.so; raises (never falls back) when it is missing."""
    global _LIB
    if _LIB is not None:
        return _LIB
    path = lib_path()
    if not os.path.exists(path):
        raise SnkError(f"{path} not built — run `python -c 'import __graft_entry__ as g; g.build()'` "
                       "(or `make -C snake-vk_b200`); there is no CPU fallback")
    L = C.CDLL(path)
    vp, u32, u64, sz, i32 = C.c_void_p, C.c_uint32, C.c_uint64, C.c_size_t, C.c_int
    sig = {
        "snkpt_create": [vp, vp], "snkpt_default_settings": [vp], "snkpt_resize": [vp, u32, u32],
        "snkpt_set_rows": [vp, u32, u32], "snkpt_set_halo": [vp, u32, u32], "snkpt_reset_history": [vp],
        "snkpt_upload_mesh": [vp, vp, vp, vp, vp, u32, vp, u32, vp, u32, vp, vp],
        "snkpt_update_positions": [vp, vp, vp, u32], "snkpt_refit_blas": [vp, u32, vp], "snkpt_rebuild_blas": [vp, u32, vp],
        "snkpt_set_materials": [vp, vp, u32, u32], "snkpt_set_texture": [vp, u32, u32, u32, vp, u32], "snkpt_set_lights": [vp, vp, sz],
        "snkpt_set_transforms": [vp, vp, vp, u32, u32], "snkpt_set_instances": [vp, vp, vp, u32, vp, u32],
        "snkpt_build_tlas": [vp, vp], "snkpt_set_frame": [vp, vp, vp], "snkpt_trace_primary": [vp, vp, vp],
        "snkpt_render": [vp, vp, vp], "snkpt_accumulate": [vp, i32, vp], "snkpt_taa_resolve": [vp, vp], "snkpt_read_output": [vp, vp, u32], "snkpt_read_output_async": [vp, vp, u32, u32], "snkpt_read_output_wait": [vp, u32],
        "snkpt_particles_create": [vp, u32, i32], "snkpt_particles_write": [vp, vp, u32, u32], "snkpt_particles_read": [vp, vp, u32, u32],
        "snkpt_particles_update": [vp, vp, u32, vp], "snkpt_particles_rays": [vp, vp], "snkpt_sort_uints_by_key": [vp, vp, vp, u32, vp],
        "snkpt_read_aux": [vp, u32, vp, sz], "snkpt_aux_size": [vp, u32, vp],
        "snkpt_write_reservoirs": [vp, u32, vp, u32, u32], "snkpt_device_ptr": [vp, u32, vp, vp],
        "snkpt_get_stats": [vp, vp], "snkpt_synchronize": [vp], "snkpt_trace_rays": [vp, vp, u64, u32, vp, vp],
        "snkpt_count_work": [vp, vp, vp, u64, u32, vp, vp],
        "snkpt_set_profiling": [vp, u32], "snkpt_get_launch_timings": [vp, vp, u32, vp],
        "snkpt_group_unique_id": [vp], "snkpt_group_attach_nccl": [vp, i32, i32, vp, u32, u32], "snkpt_group_attach_local": [vp, i32, u32, u32],
        "snkpt_group_detach": [vp], "snkpt_group_info": [vp, vp, vp, vp], "snkpt_group_set_bands": [vp, i32, vp],
    }
    for name, args in sig.items():
        fn = getattr(L, name)
        fn.argtypes = args
        fn.restype = C.c_int
    L.snkpt_default_settings.restype = None
    L.snkpt_particles_default_params.argtypes = [vp]
    L.snkpt_particles_default_params.restype = None
    L.snkpt_frame_jitter.argtypes = [u32, u32, u32, vp, vp]
    L.snkpt_frame_jitter.restype = None
    L.snkpt_group_band.argtypes = [u32, i32, i32, vp, vp]
    L.snkpt_group_band.restype = None
    L.snkpt_destroy.argtypes = [vp]
    L.snkpt_destroy.restype = None
    L.snkpt_last_error.argtypes = [vp]
    L.snkpt_last_error.restype = C.c_char_p
    _LIB = L
    return L


EXPORTS = ["snkpt_create", "snkpt_destroy", "snkpt_last_error", "snkpt_default_settings", "snkpt_resize", "snkpt_set_rows", "snkpt_set_halo",
           "snkpt_reset_history", "snkpt_upload_mesh", "snkpt_update_positions", "snkpt_refit_blas", "snkpt_rebuild_blas",
           "snkpt_set_materials", "snkpt_set_lights", "snkpt_set_transforms", "snkpt_set_instances", "snkpt_build_tlas",
           "snkpt_set_frame", "snkpt_trace_primary", "snkpt_render", "snkpt_accumulate", "snkpt_read_output", "snkpt_read_aux",
           "snkpt_aux_size", "snkpt_write_reservoirs", "snkpt_device_ptr", "snkpt_get_stats", "snkpt_synchronize",
           "snkpt_trace_rays", "snkpt_count_work", "snkpt_set_profiling", "snkpt_get_launch_timings", "snkpt_taa_resolve", "snkpt_set_texture", "snkpt_read_output_async", "snkpt_read_output_wait", "snkpt_frame_jitter", "snkpt_particles_default_params",
           "snkpt_particles_create", "snkpt_particles_write", "snkpt_particles_read", "snkpt_particles_update", "snkpt_particles_rays", "snkpt_sort_uints_by_key",
           "snkpt_group_band", "snkpt_group_unique_id", "snkpt_group_attach_nccl", "snkpt_group_attach_local", "snkpt_group_detach", "snkpt_group_info", "snkpt_group_set_bands"]

CONFIG_DTYPE = np.dtype([("width", "<u4"), ("height", "<u4"), ("device", "<i4"), ("max_instances", "<u4"),
                         ("max_transforms", "<u4"), ("max_materials", "<u4"), ("row_begin", "<u4"), ("row_end", "<u4"),
                         ("reserved", "<u4", (4,))])
SETTINGS_DTYPE = np.dtype([("num_temporal", "<i4"), ("num_spatial", "<i4"), ("mode", "<u4"), ("ris_iters", "<u4"),
                           ("m_cap", "<u4"), ("neighbour_radius", "<f4"), ("normal_threshold", "<f4"),
                           ("max_bounces", "<u4"), ("spp", "<u4"), ("flags", "<u4"), ("reserved", "<u4", (6,))])
STATS_DTYPE = np.dtype([("rays_primary", "<u8"), ("rays_shadow", "<u8"), ("rays_secondary", "<u8"), ("kernel_launches", "<u8"),
                        ("blas_nodes", "<u8"), ("blas_tris", "<u8"), ("tlas_nodes", "<u8"), ("ms_primary", "<f4"),
                        ("ms_restir1", "<f4"), ("ms_restir2", "<f4"), ("ms_bounce", "<f4"), ("ms_total", "<f4"),
                        ("ms_build_last", "<f4"), ("stack_overflows", "<u4"), ("total_rays_primary", "<u8"),
                        ("total_rays_shadow", "<u8"), ("total_rays_secondary", "<u8")], align=True)
assert STATS_DTYPE.itemsize == 112
SUBMESH_DTYPE = np.dtype([("num_indices", "<u4"), ("num_vertices", "<u4"), ("base_vertex", "<u4"), ("base_index", "<u4"),
                          ("material_index", "<u4")])
MESH_ENTRY_DTYPE = np.dtype([("data_start_indices_idx", "<u4"), ("data_start_vertex_idx", "<u4"), ("num_submeshes", "<u4"),
                             ("blas_id", "<u4", (64,))])
BLAS_INFO_DTYPE = np.dtype([("node_root", "<u4"), ("node_count", "<u4"), ("tri_base", "<u4"), ("tri_count", "<u4"),
                            ("bounds_lo", "<f4", (3,)), ("bounds_hi", "<f4", (3,))])
LAUNCH_TIMING_DTYPE = np.dtype([("name", "S32"), ("ms", "<f4"), ("is_trace", "<u4"), ("units", "<u8"), ("live", "<u8"),
                                ("nodes", "<u8"), ("tris", "<u8"), ("insts", "<u8"), ("alg_bytes", "<u8")])
assert LAUNCH_TIMING_DTYPE.itemsize == 88
assert SETTINGS_DTYPE.itemsize == 64 and CONFIG_DTYPE.itemsize == 48 and SUBMESH_DTYPE.itemsize == 20

MODE_REFERENCE, MODE_NORTHSTAR = 0, 1
RAY_CLOSEST, RAY_ANY, RAY_DEVICE_PTRS = 0, 1, 2
RT_NO_RESTIR, RT_NO_BOUNCE, RT_UNBIASED_REUSE = 1, 2, 4

AUX = dict(prim=(0, np.uint32, 1), inst=(1, np.uint32, 1), depth=(2, np.float32, 1), normal=(3, np.float16, 4),
           albedo=(4, np.int8, 4), rma=(5, np.float16, 4), motion=(6, np.float16, 2), flags=(7, np.uint16, 1),
           prev_normal=(8, np.float16, 4), prev_depth=(9, np.float32, 1), reservoirs_final=(10, scenes.RESERVOIR_DTYPE, 1),
           reservoirs_temporal=(11, scenes.RESERVOIR_DTYPE, 1), hit_t=(14, np.float32, 1), bary=(15, np.float32, 2),
           radiance=(16, np.float32, 4), accum=(17, np.float32, 4), direct=(18, np.float32, 4),
           taa_history=(21, np.float16, 4), taa_prev_velocity=(22, np.float16, 2))


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


PARTICLE_DTYPE = np.dtype([("position_radius", "<f4", (4,)), ("velocity", "<f4", (4,)), ("colour", "<f4", (4,))])
PHYSICS_DTYPE = np.dtype([(k, "<f4") for k in ("particle_restitution", "timestep", "friction_coefficient", "repulsion_factor", "penetration_slop",
                                               "restitution_slop", "sleep_threshold")])
assert PARTICLE_DTYPE.itemsize == 48 and PHYSICS_DTYPE.itemsize == 28


def physics_params(**kw) -> np.ndarray:
    """PhysicsParamsUBO defaults (Core/headers/scene/ParticleSystem.h:16-24)"""
    p = np.zeros((), PHYSICS_DTYPE)
    load_library().snkpt_particles_default_params(_p(p))
    for k, v in kw.items():
        p[k] = v
    return p


def frame_jitter(frame_idx: int, internal_width: int, output_width: int | None = None):
    """GetFrameJitter (RenderCommon.h:50-59) → (push-constant jitter, equivalent camera-ray offset in pixels)."""
    j, px = np.zeros(2, np.float32), np.zeros(2, np.float32)
    load_library().snkpt_frame_jitter(frame_idx, internal_width, output_width or internal_width, _p(j), _p(px))
    return j, px


def default_settings(**kw) -> np.ndarray:
    s = np.zeros((), SETTINGS_DTYPE)
    load_library().snkpt_default_settings(_p(s))
    for k, v in kw.items():
        s[k] = v
    return s


def group_band(height: int, world: int, rank: int) -> tuple[int, int]:
    """rows [b0, b1) of band `rank` of `world` equal bands (snkpt_group_band)"""
    a, b = C.c_uint32(), C.c_uint32()
    load_library().snkpt_group_band(height, world, rank, C.byref(a), C.byref(b))
    return a.value, b.value


def group_unique_id() -> bytes:
    """NCCL unique id for snkpt_group_attach_nccl (made on rank 0, broadcast by the caller)"""
    buf = C.create_string_buffer(128)
    rc = load_library().snkpt_group_unique_id(buf)
    if rc != 0:
        raise SnkError(f"snkpt_group_unique_id failed ({rc}): libnccl.so.2 not found (SNKPT_NCCL_LIB)")
    return buf.raw


def group_attach_local(tracers, radius_rows: int = 30, margin_rows: int = 16):
    """the contexts of ONE process (ordered by rank) become a row-band group; each must then be driven by its own host thread"""
    arr = (C.c_void_p * len(tracers))(*[t.h for t in tracers])
    rc = load_library().snkpt_group_attach_local(arr, len(tracers), radius_rows, margin_rows)
    if rc != 0:
        raise SnkError(f"snkpt_group_attach_local failed ({rc}): {load_library().snkpt_last_error(tracers[0].h).decode()}")


class PathTracer:
    """One context per GPU (include/snkpt.h)."""

    def __init__(self, width: int, height: int, device: int = 0, rows=(0, 0), **caps):
        self.L = load_library()
        cfg = np.zeros((), CONFIG_DTYPE)
        cfg["width"], cfg["height"], cfg["device"] = width, height, device
        cfg["row_begin"], cfg["row_end"] = rows
        for k, v in caps.items():
            cfg[k] = v
        h = C.c_void_p()
        rc = self.L.snkpt_create(_p(cfg), C.byref(h))
        if rc != 0:
            raise SnkError(f"snkpt_create failed ({rc}): {self.L.snkpt_last_error(None).decode()}")
        self.h = h
        self.width, self.height = width, height
        self.rows = (rows[0], rows[1] or height)

    def close(self):
        if getattr(self, "h", None):
            self.L.snkpt_destroy(self.h)
            self.h = None

    __del__ = close

    def _ck(self, rc: int, what: str):
        if rc != 0:
            raise SnkError(f"{what} failed ({rc}): {self.L.snkpt_last_error(self.h).decode()}")

    # ---- scene
    def upload_mesh(self, mesh: "scenes.Mesh", stream=None) -> np.ndarray:
        entry = np.zeros((), MESH_ENTRY_DTYPE)
        sub = np.ascontiguousarray(mesh.submeshes.astype(np.uint32))
        self._ck(self.L.snkpt_upload_mesh(self.h, _p(mesh.positions), _p(mesh.normals), _p(mesh.tangents), _p(mesh.uvs),
                                          mesh.positions.shape[0], _p(mesh.indices), mesh.indices.size, _p(sub),
                                          sub.shape[0], _p(entry), stream), "snkpt_upload_mesh")
        return entry

    def set_materials(self, mats: np.ndarray, first=0):
        mats = np.ascontiguousarray(mats)
        self._ck(self.L.snkpt_set_materials(self.h, _p(mats), first, mats.shape[0]), "snkpt_set_materials")

    def set_lights(self, blob: bytes):
        buf = np.frombuffer(blob, np.uint8).copy()
        self._ck(self.L.snkpt_set_lights(self.h, _p(buf), buf.size), "snkpt_set_lights")

    def set_transforms(self, cur: np.ndarray, prev: np.ndarray | None = None, first=0):
        cur = np.ascontiguousarray(cur, np.float32).reshape(-1, 16)
        prev = None if prev is None else np.ascontiguousarray(prev, np.float32).reshape(-1, 16)
        self._ck(self.L.snkpt_set_transforms(self.h, _p(cur), _p(prev), first, cur.shape[0]), "snkpt_set_transforms")

    def set_instances(self, instances: np.ndarray, blas_of_instance: np.ndarray, emissive_idx: np.ndarray):
        instances = np.ascontiguousarray(instances, np.uint32).reshape(-1, 5)
        blas_of_instance = np.ascontiguousarray(blas_of_instance, np.uint32)
        emissive_idx = np.ascontiguousarray(emissive_idx, np.uint32)
        self._ck(self.L.snkpt_set_instances(self.h, _p(instances), _p(blas_of_instance), instances.shape[0],
                                            _p(emissive_idx) if emissive_idx.size else None, emissive_idx.size),
                 "snkpt_set_instances")

    def build_tlas(self, stream=None):
        self._ck(self.L.snkpt_build_tlas(self.h, stream), "snkpt_build_tlas")

    def upload_scene(self, scene: "scenes.Scene") -> dict:
        """What MeshBufferManager + the scene buffer systems do at load / FrameStart."""
        flat = scene.flatten()
        for m, (i0, v0, b0) in zip(scene.meshes, flat["mesh_entries"]):
            e = self.upload_mesh(m)
            assert (int(e["data_start_indices_idx"]), int(e["data_start_vertex_idx"]), int(e["blas_id"][0])) == (i0, v0, b0)
        for i, (texels, srgb) in enumerate(getattr(scene, "textures", [])):
            self.set_texture(i, texels, srgb)
        self.set_materials(flat["materials"])
        self.set_lights(flat["lights"])
        self.set_transforms(flat["transforms"], flat["prev_transforms"])
        self.set_instances(flat["instances"], flat["blas_of_instance"], flat["emissive_idx"])
        self.build_tlas()
        return flat

    def set_texture(self, idx: int, rgba8: np.ndarray | None, srgb: bool = True):
        """Texture2DAsset upload into slot `idx` of the albedo texture table (None frees the slot)."""
        if rgba8 is None:
            self._ck(self.L.snkpt_set_texture(self.h, idx, 0, 0, None, 0), "snkpt_set_texture")
            return
        t = np.ascontiguousarray(rgba8, np.uint8)
        assert t.ndim == 3 and t.shape[2] == 4
        self._ck(self.L.snkpt_set_texture(self.h, idx, t.shape[1], t.shape[0], _p(t), int(bool(srgb))), "snkpt_set_texture")

    # ---- frame
    def set_frame(self, cur: np.ndarray, prev: np.ndarray | None = None):
        self._cur = np.ascontiguousarray(cur)
        self._prev = self._cur if prev is None else np.ascontiguousarray(prev)
        self._ck(self.L.snkpt_set_frame(self.h, _p(self._cur), _p(self._prev)), "snkpt_set_frame")

    def trace_primary(self, jitter=None, stream=None):
        j = None if jitter is None else np.asarray(jitter, np.float32)
        self._ck(self.L.snkpt_trace_primary(self.h, _p(j), stream), "snkpt_trace_primary")

    def render(self, settings: np.ndarray | None = None, stream=None):
        s = default_settings() if settings is None else np.ascontiguousarray(settings)
        self._ck(self.L.snkpt_render(self.h, _p(s), stream), "snkpt_render")

    def accumulate(self, reset=False, stream=None):
        self._ck(self.L.snkpt_accumulate(self.h, int(reset), stream), "snkpt_accumulate")

    # ---- particle collision pass (ParticleSystem)
    def particles_create(self, count: int, init=True):
        self._ck(self.L.snkpt_particles_create(self.h, count, int(init)), "snkpt_particles_create")
        self.num_particles = count

    def particles_write(self, particles: np.ndarray, first=0):
        a = np.ascontiguousarray(particles)
        assert a.dtype == PARTICLE_DTYPE
        self._ck(self.L.snkpt_particles_write(self.h, _p(a), first, a.size), "snkpt_particles_write")

    def particles_read(self) -> np.ndarray:
        out = np.zeros(self.num_particles, PARTICLE_DTYPE)
        self._ck(self.L.snkpt_particles_read(self.h, _p(out), 0, out.size), "snkpt_particles_read")
        return out

    def particles_update(self, params: np.ndarray | None = None, substeps=3, stream=None):
        p = None if params is None else np.ascontiguousarray(params)
        self._ck(self.L.snkpt_particles_update(self.h, _p(p), substeps, stream), "snkpt_particles_update")

    def sort_uints_by_key(self, device_keys: int, device_values: int, count: int, stream=None):
        """SortUintsByKey (ParticleSystem.cpp:14): stable in-place sort of DEVICE-resident uint32 (key, value) pairs"""
        self._ck(self.L.snkpt_sort_uints_by_key(self.h, device_keys, device_values, count, stream), "snkpt_sort_uints_by_key")

    def particles_rays(self) -> int:
        v = C.c_uint64(0)
        self._ck(self.L.snkpt_particles_rays(self.h, C.byref(v)), "snkpt_particles_rays")
        return v.value

    def taa_resolve(self, stream=None):
        """TAA_ResolvePass::RecordCommandBuffer on the output image of the last render (include/snkpt.h)."""
        self._ck(self.L.snkpt_taa_resolve(self.h, stream), "snkpt_taa_resolve")

    def synchronize(self):
        self._ck(self.L.snkpt_synchronize(self.h), "snkpt_synchronize")

    def set_rows(self, row_begin: int, row_end: int):
        self._ck(self.L.snkpt_set_rows(self.h, row_begin, row_end), "snkpt_set_rows")
        self.rows = (row_begin, row_end or self.height)

    def set_halo(self, pass1_halo: int, gbuffer_halo: int):
        self._ck(self.L.snkpt_set_halo(self.h, pass1_halo, gbuffer_halo), "snkpt_set_halo")

    def reset_history(self):
        self._ck(self.L.snkpt_reset_history(self.h), "snkpt_reset_history")

    # ---- row-band groups (strong scaling; include/snkpt.h "Row-band GROUPS") -------------------------------------
    def group_attach_nccl(self, rank: int, world: int, unique_id: bytes, radius_rows: int = 30, margin_rows: int = 16):
        """one process per GPU: this context renders band `rank` of `world`; reservoir rows travel by ncclSend / ncclRecv"""
        buf = C.create_string_buffer(bytes(unique_id), 128)
        self._ck(self.L.snkpt_group_attach_nccl(self.h, rank, world, buf, radius_rows, margin_rows), "snkpt_group_attach_nccl")

    def group_set_bands(self, row_starts):
        """bands of unequal height for the next group_attach_* (world + 1 row boundaries); None -> equal bands"""
        if row_starts is None:
            self._ck(self.L.snkpt_group_set_bands(self.h, 0, None), "snkpt_group_set_bands")
        else:
            a = np.ascontiguousarray(row_starts, np.uint32)
            self._ck(self.L.snkpt_group_set_bands(self.h, a.size - 1, _p(a)), "snkpt_group_set_bands")

    def group_detach(self):
        self._ck(self.L.snkpt_group_detach(self.h), "snkpt_group_detach")

    def group_info(self):
        r, w, b = C.c_int(), C.c_int(), C.c_uint64()
        self._ck(self.L.snkpt_group_info(self.h, C.byref(r), C.byref(w), C.byref(b)), "snkpt_group_info")
        return r.value, w.value, b.value

    def resize(self, width: int, height: int):
        """window resize (EditorLayer.cpp:351-366, :415-423): every per-pixel resource is recreated, the history is invalid"""
        self._ck(self.L.snkpt_resize(self.h, width, height), "snkpt_resize")
        self.width, self.height = width, height
        self.rows = (0, height)

    # ---- readback
    def read_output(self, fmt="f16", out: np.ndarray | None = None) -> np.ndarray:
        nrows = self.rows[1] - self.rows[0]
        dt = np.float16 if fmt == "f16" else np.float32
        if out is None:
            out = np.empty((nrows, self.width, 4), dt)
        self._ck(self.L.snkpt_read_output(self.h, _p(out), 0 if fmt == "f16" else 1), "snkpt_read_output")
        return out

    def read_output_async(self, out: np.ndarray, slot: int = 0, fmt="f16"):
        """enqueue the read-back of the frame rendered last into `out` (page-locked host memory) without blocking"""
        self._ck(self.L.snkpt_read_output_async(self.h, _p(out), 0 if fmt == "f16" else 1, slot), "snkpt_read_output_async")

    def wait_output(self, slot: int = 0):
        self._ck(self.L.snkpt_read_output_wait(self.h, slot), "snkpt_read_output_wait")

    def read_aux(self, name: str) -> np.ndarray:
        what, dt, ch = AUX[name]
        out = np.zeros((self.height, self.width, ch) if ch > 1 else (self.height, self.width), dt)
        self._ck(self.L.snkpt_read_aux(self.h, what, _p(out), out.nbytes), f"snkpt_read_aux({name})")
        return out

    def read_raw(self, what: int, dtype) -> np.ndarray:
        n = C.c_size_t()
        self._ck(self.L.snkpt_aux_size(self.h, what, C.byref(n)), "snkpt_aux_size")
        out = np.zeros(n.value // np.dtype(dtype).itemsize, dtype)
        if n.value:
            self._ck(self.L.snkpt_read_aux(self.h, what, _p(out), out.nbytes), "snkpt_read_aux")
        return out

    def blas_table(self) -> np.ndarray:
        return self.read_raw(20, BLAS_INFO_DTYPE)

    def write_reservoirs(self, half: int, res: np.ndarray, first=0):
        res = np.ascontiguousarray(res.reshape(-1))
        self._ck(self.L.snkpt_write_reservoirs(self.h, half, _p(res), first, res.size), "snkpt_write_reservoirs")

    def device_ptr(self, what: int):
        p, n = C.c_void_p(), C.c_size_t()
        self._ck(self.L.snkpt_device_ptr(self.h, what, C.byref(p), C.byref(n)), "snkpt_device_ptr")
        return p.value, n.value

    def stats(self) -> np.ndarray:
        s = np.zeros((), STATS_DTYPE)
        self._ck(self.L.snkpt_get_stats(self.h, _p(s)), "snkpt_get_stats")
        return s

    def set_profiling(self, level: int):
        self._ck(self.L.snkpt_set_profiling(self.h, level), "snkpt_set_profiling")

    def launch_timings(self) -> np.ndarray:
        out = np.zeros(2048, LAUNCH_TIMING_DTYPE)
        n = C.c_uint32()
        self._ck(self.L.snkpt_get_launch_timings(self.h, _p(out), out.size, C.byref(n)), "snkpt_get_launch_timings")
        return out[:n.value]

    # ---- ray queries
    def trace_rays(self, rays: np.ndarray, any_hit=False, flags=0) -> np.ndarray:
        rays = np.ascontiguousarray(rays)
        assert rays.dtype == scenes.RAY_DTYPE
        hits = np.zeros(rays.shape[0], scenes.HIT_DTYPE)
        self._ck(self.L.snkpt_trace_rays(self.h, _p(rays), rays.shape[0], (RAY_ANY if any_hit else 0) | flags, _p(hits), None),
                 "snkpt_trace_rays")
        return hits

    def count_work(self, rays: np.ndarray, hits: np.ndarray, any_hit=False) -> np.ndarray:
        rays, hits = np.ascontiguousarray(rays), np.ascontiguousarray(hits)
        out = np.zeros((rays.shape[0], 3), np.uint32)
        self._ck(self.L.snkpt_count_work(self.h, _p(rays), _p(hits), rays.shape[0], RAY_ANY if any_hit else 0, _p(out), None),
                 "snkpt_count_work")
        return out


class RT:
    """Mirror of `SNAKE::RT` (Core/headers/rendering/Raytracing.h:11-31).

    Init(scene, output_size)             ~ RT::Init + EditorLayer::InitGBuffer
    RecordRenderCmdBuf(ubo, prev, st)    ~ GBufferPass + RT::RecordRenderCmdBuf for one frame
    """

    class RtSettings:
        def __init__(self, num_temporal_resamples=1, num_spatial_resamples=3):
            self.num_temporal_resamples = num_temporal_resamples
            self.num_spatial_resamples = num_spatial_resamples

    def __init__(self):
        self.pt: PathTracer | None = None

    def Init(self, scene: "scenes.Scene", device: int = 0, rows=(0, 0)):
        self.pt = PathTracer(scene.width, scene.height, device, rows)
        self.flat = self.pt.upload_scene(scene)
        return self

    def RecordRenderCmdBuf(self, ubo, prev_ubo, settings: "RT.RtSettings | None" = None, mode=MODE_REFERENCE, **extra):
        st = settings or RT.RtSettings()
        s = default_settings(**{**dict(num_temporal=st.num_temporal_resamples, num_spatial=st.num_spatial_resamples, mode=mode), **extra})
        self.pt.set_frame(ubo, prev_ubo)
        self.pt.trace_primary()
        self.pt.render(s)
