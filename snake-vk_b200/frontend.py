"""ctypes mirror of include/snkscene.h — the C++ scene front-end (snake-vk_b200/host/snkscene.cpp):
entity hierarchy → the flat buffers of the path-tracing pass, as the reference's TransformSystem / TransformBufferSystem /
RaytracingInstanceBufferSystem / LightBufferSystem / SceneInfoBufferSystem produce them."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import scenes
from .api import SnkError, load_library

NONE = 0xFFFFFFFF
_LIB = None


def lib() -> C.CDLL:
    global _LIB
    if _LIB is None:
        load_library()  # libsnkscene.so links libsnkpt.so
        path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "libsnkscene.so")
        if not os.path.exists(path):
            raise SnkError(f"{path} is not built (make -C snake-vk_b200)")
        L = C.CDLL(path)
        vp, u32, i32, f32, sz = C.c_void_p, C.c_uint32, C.c_int, C.c_float, C.c_size_t
        sig = {
            "snkscene_create": ([], vp), "snkscene_destroy": ([vp], None), "snkscene_last_error": ([vp], C.c_char_p),
            "snkscene_create_entity": ([vp, u32], u32), "snkscene_set_parent": ([vp, u32, u32], i32),
            "snkscene_set_transform": ([vp, u32, vp, vp, vp], i32), "snkscene_set_absolute": ([vp, u32, i32], i32),
            "snkscene_get_matrix": ([vp, u32, vp], i32), "snkscene_get_abs": ([vp, u32, vp, vp, vp, vp], i32),
            "snkscene_register_mesh": ([vp, vp, vp, u32], u32), "snkscene_set_material": ([vp, u32, vp], i32),
            "snkscene_set_static_mesh": ([vp, u32, u32, vp, u32], i32), "snkscene_remove_static_mesh": ([vp, u32], i32),
            "snkscene_set_directional_light": ([vp, vp, f32, f32], i32), "snkscene_set_pointlight": ([vp, u32, vp, f32, f32, f32], i32),
            "snkscene_remove_pointlight": ([vp, u32], i32), "snkscene_set_camera": ([vp, u32, f32, f32, f32, f32], i32),
            "snkscene_frame_start": ([vp, f32, f32], i32), "snkscene_frame_end": ([vp], i32), "snkscene_invalidate_history": ([vp], i32),
            "snkscene_transforms": ([vp, i32, vp], vp), "snkscene_instances": ([vp, vp], vp), "snkscene_blas_of_instance": ([vp, vp], vp),
            "snkscene_emissive": ([vp, vp], vp), "snkscene_materials": ([vp, vp], vp), "snkscene_lights": ([vp, vp], vp),
            "snkscene_ubo": ([vp, i32], vp), "snkscene_push": ([vp, vp, vp], i32),
        }
        for name, (args, res) in sig.items():
            fn = getattr(L, name)
            fn.argtypes, fn.restype = args, res
        L._names = list(sig)
        _LIB = L
    return _LIB


def _f3(v):
    return None if v is None else np.ascontiguousarray(v, np.float32)


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class SceneFrontEnd:
    def __init__(self):
        self.L = lib()
        self.h = self.L.snkscene_create()

    def close(self):
        if getattr(self, "h", None):
            self.L.snkscene_destroy(self.h)
            self.h = None

    __del__ = close

    def _ck(self, rc, what):
        if rc != 0:
            raise SnkError(f"{what}: {self.L.snkscene_last_error(self.h).decode()}")

    # ---- entities
    def create_entity(self, parent=NONE, pos=None, euler=None, scale=None) -> int:
        e = self.L.snkscene_create_entity(self.h, parent)
        if e == NONE:
            raise SnkError("snkscene_create_entity: unknown parent")
        if pos is not None or euler is not None or scale is not None:
            self.set_transform(e, pos, euler, scale)
        return e

    def set_parent(self, ent, parent):
        self._ck(self.L.snkscene_set_parent(self.h, ent, parent), "snkscene_set_parent")

    def set_transform(self, ent, pos=None, euler=None, scale=None):
        a, b, c = _f3(pos), _f3(euler), _f3(scale)
        self._ck(self.L.snkscene_set_transform(self.h, ent, _p(a), _p(b), _p(c)), "snkscene_set_transform")

    def set_absolute(self, ent, absolute=True):
        self._ck(self.L.snkscene_set_absolute(self.h, ent, int(absolute)), "snkscene_set_absolute")

    def matrix(self, ent) -> np.ndarray:
        m = np.zeros((4, 4), np.float32)
        self._ck(self.L.snkscene_get_matrix(self.h, ent, _p(m)), "snkscene_get_matrix")
        return m  # [col][row]

    def abs(self, ent):
        out = [np.zeros(3, np.float32) for _ in range(4)]
        self._ck(self.L.snkscene_get_abs(self.h, ent, *[_p(o) for o in out]), "snkscene_get_abs")
        return dict(pos=out[0], scale=out[1], orientation=out[2], forward=out[3])

    # ---- components
    def register_mesh(self, entry: np.ndarray, submeshes: np.ndarray) -> int:
        entry = np.ascontiguousarray(entry)
        subs = np.ascontiguousarray(submeshes, np.uint32).reshape(-1, 5)
        m = self.L.snkscene_register_mesh(self.h, _p(entry), _p(subs), subs.shape[0])
        if m == NONE:
            raise SnkError("snkscene_register_mesh: bad arguments")
        return m

    def set_material(self, slot: int, material: np.ndarray):
        m = np.ascontiguousarray(material)
        assert m.dtype == scenes.MATERIAL_DTYPE and m.size == 1
        self._ck(self.L.snkscene_set_material(self.h, slot, _p(m)), "snkscene_set_material")

    def set_static_mesh(self, ent, mesh, material_slots):
        ms = np.ascontiguousarray(material_slots, np.uint32)
        self._ck(self.L.snkscene_set_static_mesh(self.h, ent, mesh, _p(ms), ms.size), "snkscene_set_static_mesh")

    def remove_static_mesh(self, ent):
        self._ck(self.L.snkscene_remove_static_mesh(self.h, ent), "snkscene_remove_static_mesh")

    def set_directional_light(self, colour, theta, phi):
        self._ck(self.L.snkscene_set_directional_light(self.h, _p(_f3(colour)), theta, phi), "snkscene_set_directional_light")

    def set_pointlight(self, ent, colour, exp=0.05, linear=0.25, constant=0.5):
        self._ck(self.L.snkscene_set_pointlight(self.h, ent, _p(_f3(colour)), exp, linear, constant), "snkscene_set_pointlight")

    def remove_pointlight(self, ent):
        self._ck(self.L.snkscene_remove_pointlight(self.h, ent), "snkscene_remove_pointlight")

    def set_camera(self, ent, fov=90.0, aspect=16.0 / 9.0, z_near=0.1, z_far=2500.0):
        self._ck(self.L.snkscene_set_camera(self.h, ent, fov, aspect, z_near, z_far), "snkscene_set_camera")

    # ---- frames
    def frame_start(self, time_elapsed_ms: float, delta_ms: float = 16.6667):
        self._ck(self.L.snkscene_frame_start(self.h, time_elapsed_ms, delta_ms), "snkscene_frame_start")

    def frame_end(self):
        self._ck(self.L.snkscene_frame_end(self.h), "snkscene_frame_end")

    def invalidate_history(self):
        self._ck(self.L.snkscene_invalidate_history(self.h), "snkscene_invalidate_history")

    def _array(self, fn, dtype, per, *pre):
        n = C.c_uint32(0)
        ptr = fn(self.h, *pre, C.byref(n))
        count = n.value
        if not ptr or count == 0:
            return np.zeros((0,) + ((per,) if per > 1 else ()), dtype)
        nbytes = count * per * np.dtype(dtype).itemsize
        buf = (C.c_uint8 * nbytes).from_address(ptr)
        a = np.frombuffer(buf, dtype).copy()
        return a.reshape(count, per) if per > 1 else a

    def flat(self) -> dict:
        """the buffers of the last frame_start in the layout of scenes.Scene.flatten()"""
        nb = C.c_size_t(0)
        lp = self.L.snkscene_lights(self.h, C.byref(nb))
        lights = bytes((C.c_uint8 * nb.value).from_address(lp)) if lp and nb.value else b""
        ubo = [np.frombuffer((C.c_uint8 * 240).from_address(self.L.snkscene_ubo(self.h, k)), scenes.UBO_DTYPE).copy().reshape(()) for k in (0, 1)]
        mats = self._array(self.L.snkscene_materials, np.uint8, 64)
        return dict(transforms=self._array(self.L.snkscene_transforms, np.float32, 16, 0),
                    prev_transforms=self._array(self.L.snkscene_transforms, np.float32, 16, 1),
                    instances=self._array(self.L.snkscene_instances, np.uint32, 5),
                    blas_of_instance=self._array(self.L.snkscene_blas_of_instance, np.uint32, 1),
                    emissive_idx=self._array(self.L.snkscene_emissive, np.uint32, 1),
                    materials=np.frombuffer(mats.tobytes(), scenes.MATERIAL_DTYPE).copy(), lights=lights, ubo=ubo[0], prev_ubo=ubo[1])

    def push(self, pt, stream=None):
        """the frame's upload calls + TLAS build + frame UBOs on a PathTracer (include/snkpt.h)"""
        self._ck(self.L.snkscene_push(self.h, pt.h, stream), "snkscene_push")
