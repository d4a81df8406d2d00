"""Mesh readers for the formats the reference ships.

`load_glb` reads a binary glTF 2.0 file the way AssetLoader::LoadMeshDataFromRawFile
(Core/src/assets/AssetLoader.cpp:337-466) consumes it through assimp for the two core
meshes (Core/res/meshes/cube.glb, sphere.glb; AssetManager.cpp:34-65): one submesh per
primitive, float3 positions / normals, float2 uvs, u32 indices appended into one
MeshData with per-submesh base_vertex / base_index.
"""
from __future__ import annotations

import json
import struct

import numpy as np

from .scenes import Mesh

_COMP = {5120: np.int8, 5121: np.uint8, 5122: np.int16, 5123: np.uint16, 5125: np.uint32, 5126: np.float32}
_NCOMP = {"SCALAR": 1, "VEC2": 2, "VEC3": 3, "VEC4": 4}


def _accessor(j, blob, idx):
    a = j["accessors"][idx]
    bv = j["bufferViews"][a["bufferView"]]
    dt, nc = np.dtype(_COMP[a["componentType"]]), _NCOMP[a["type"]]
    off = bv.get("byteOffset", 0) + a.get("byteOffset", 0)
    stride = bv.get("byteStride", 0) or dt.itemsize * nc
    raw = np.frombuffer(blob, np.uint8, count=stride * (a["count"] - 1) + dt.itemsize * nc, offset=off)
    if stride == dt.itemsize * nc:
        return raw.view(dt).reshape(a["count"], nc).copy()
    rows = np.lib.stride_tricks.as_strided(raw, (a["count"], dt.itemsize * nc), (stride, 1))
    return np.ascontiguousarray(rows).view(dt).reshape(a["count"], nc)


def load_glb(path: str) -> Mesh:
    b = open(path, "rb").read()
    magic, version, _ = struct.unpack_from("<III", b, 0)
    if magic != 0x46546C67 or version != 2:
        raise ValueError(f"{path}: not a glTF 2.0 binary")
    pos = 12
    j, blob = None, b""
    while pos < len(b):
        clen, ctype = struct.unpack_from("<II", b, pos)
        chunk = b[pos + 8:pos + 8 + clen]
        if ctype == 0x4E4F534A:
            j = json.loads(chunk)
        elif ctype == 0x004E4942:
            blob = chunk
        pos += 8 + clen
    P, N, U, I, subs = [], [], [], [], []
    v0 = i0 = 0
    for m in j["meshes"]:
        for prim in m["primitives"]:
            at = prim["attributes"]
            p = _accessor(j, blob, at["POSITION"]).astype(np.float32)
            n = _accessor(j, blob, at["NORMAL"]).astype(np.float32) if "NORMAL" in at else np.zeros_like(p)
            u = _accessor(j, blob, at["TEXCOORD_0"]).astype(np.float32) if "TEXCOORD_0" in at else np.zeros((p.shape[0], 2), np.float32)
            idx = _accessor(j, blob, prim["indices"]).astype(np.uint32).reshape(-1)
            subs.append((idx.size, p.shape[0], v0, i0, len(subs)))
            P.append(p), N.append(n), U.append(u), I.append(idx)
            v0 += p.shape[0]
            i0 += idx.size
    mesh = Mesh.single(np.concatenate(P), np.concatenate(N), np.concatenate(I), np.concatenate(U))
    mesh.submeshes = np.array(subs, np.uint32).reshape(-1, 5)
    return mesh
