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


# ---------------------------------------------------------------------------------------------------------------------
# The engine's own on-disk formats. Binary files are ByteSerializer streams
# (Core/headers/util/ByteSerializer.h:18-111): plain values are raw little-endian bytes, containers and raw arrays are
# prefixed with their BYTE size as a u64 (`size_t`). Text files are nlohmann JSON.
INVALID_UUID = 0  # UUID<uint64_t>::INVALID_UUID
CORE_SPHERE_MESH, CORE_SPHERE_MESH_DATA, CORE_MATERIAL, CORE_TEXTURE, CORE_CUBE_MESH, CORE_CUBE_MESH_DATA = 1, 2, 3, 4, 5, 6  # AssetManager.h:64-72


class _Reader:
    def __init__(self, data: bytes):
        self.b, self.pos = data, 0

    def value(self, fmt: str):
        v = struct.unpack_from("<" + fmt, self.b, self.pos)
        self.pos += struct.calcsize("<" + fmt)
        return v[0] if len(v) == 1 else v

    def blob(self) -> bytes:
        n = self.value("Q")
        if self.pos + n > len(self.b):
            raise ValueError("ByteDeserializer: container runs past the end of the file")
        out = self.b[self.pos:self.pos + n]
        self.pos += n
        return out


class _Writer:
    def __init__(self):
        self.parts = []

    def value(self, fmt: str, *v):
        self.parts.append(struct.pack("<" + fmt, *v))

    def blob(self, data: bytes):
        self.parts.append(struct.pack("<Q", len(data)) + data)

    def bytes(self) -> bytes:
        return b"".join(self.parts)


def save_meshdata(path: str, mesh: Mesh, uuid: int, name: str, material_uuids=(), texture_uuids=()):
    """AssetLoader::SerializeMeshDataBinary (Core/src/assets/AssetLoader.cpp:145-163)"""
    assert path.endswith(".meshdata")
    w = _Writer()
    w.value("Q", uuid)
    w.blob(name.encode())
    w.blob(np.ascontiguousarray(mesh.submeshes, np.uint32).tobytes())            # Submesh: 5 x u32 (MeshData.h:11-17)
    w.value("I", mesh.positions.shape[0])
    w.value("I", mesh.indices.size)
    for arr in (mesh.positions, mesh.normals, mesh.tangents):
        w.blob(np.ascontiguousarray(arr, np.float32).tobytes())                     # aiVector3D = 3 x f32
    w.blob(np.ascontiguousarray(mesh.uvs, np.float32).tobytes())                    # aiVector2D
    w.blob(np.ascontiguousarray(mesh.indices, np.uint32).tobytes())
    w.blob(np.asarray(material_uuids, np.uint64).tobytes())
    w.blob(np.asarray(texture_uuids, np.uint64).tobytes())
    open(path, "wb").write(w.bytes())


def load_meshdata(path: str) -> dict:
    """AssetLoader::DeserializeMeshData (AssetLoader.cpp:166-209) → {uuid, name, mesh, materials, textures}"""
    assert path.endswith(".meshdata")
    r = _Reader(open(path, "rb").read())
    uuid = r.value("Q")
    name = r.blob().decode()
    subs = np.frombuffer(r.blob(), np.uint32).reshape(-1, 5).copy()
    nv, ni = r.value("I"), r.value("I")
    pos, nrm, tan = (np.frombuffer(r.blob(), np.float32).reshape(-1, 3).copy() for _ in range(3))
    uvs = np.frombuffer(r.blob(), np.float32).reshape(-1, 2).copy()
    idx = np.frombuffer(r.blob(), np.uint32).copy()
    mats = np.frombuffer(r.blob(), np.uint64).copy()
    texs = np.frombuffer(r.blob(), np.uint64).copy()
    if pos.shape[0] != nv or idx.size != ni or nrm.shape[0] != nv or uvs.shape[0] != nv:
        raise ValueError(f"{path}: array sizes disagree with num_vertices / num_indices")
    mesh = Mesh(pos, nrm, uvs, tan, idx, subs)
    return dict(uuid=uuid, name=name, mesh=mesh, materials=mats, textures=texs)


_MAT_KEYS = ("albedo_tex", "normal_tex", "roughness_tex", "metallic_tex", "ao_tex")


def save_material(path: str, uuid: int, name: str, albedo=(1.0, 1.0, 1.0), emissive=0.0, roughness=0.5, metallic=0.0, ao=1.0, **tex_uuids):
    """AssetLoader::SerializeMaterialBinary (AssetLoader.cpp:127-143)"""
    assert path.endswith(".mat")
    w = _Writer()
    w.value("Q", uuid)
    w.blob(name.encode())
    w.value("3f", *albedo)
    w.value("4f", emissive, roughness, metallic, ao)
    for k in _MAT_KEYS:
        w.value("Q", tex_uuids.get(k, INVALID_UUID))
    open(path, "wb").write(w.bytes())


def load_material(path: str) -> dict:
    """AssetLoader::DeserializeMaterial (AssetLoader.cpp:212-249). `emissive` is stored but never uploaded: the shaders emit
    `albedo` when the EMISSIVE flag is set — and the flag itself is NOT in the file (it is lost on save)."""
    assert path.endswith(".mat")
    r = _Reader(open(path, "rb").read())
    out = dict(uuid=r.value("Q"), name=r.blob().decode(), albedo=tuple(r.value("3f")))
    out["emissive"], out["roughness"], out["metallic"], out["ao"] = r.value("4f")
    for k in _MAT_KEYS:
        out[k] = r.value("Q")
    return out


def save_smesh(path: str, uuid: int, data_uuid: int, name: str):
    """AssetLoader::SerializeStaticMeshAsset (AssetLoader.cpp:252-262)"""
    assert path.endswith(".smesh")
    json.dump({"UUID": uuid, "DataUUID": data_uuid, "Name": name}, open(path, "w"), indent=4)


def load_smesh(path: str) -> dict:
    """AssetLoader::DeserializeStaticMeshAsset (AssetLoader.cpp:264-292)"""
    assert path.endswith(".smesh")
    j = json.load(open(path))
    return dict(uuid=int(j["UUID"]), data_uuid=int(j["DataUUID"]), name=str(j["Name"]))


def save_scene(path: str, scene: dict):
    """SceneSerializer::SerializeScene / SerializeEntity (Core/src/scene/SceneSerializer.cpp:48-88, :145-160)"""
    ents = []
    for e in scene["entities"]:
        j = {"UUID": e["uuid"], "Name": e.get("name", "New entity"), "ParentUUID": e.get("parent", 0),
             "Transform": {"P": list(e.get("pos", (0, 0, 0))), "S": list(e.get("scale", (1, 1, 1))), "O": list(e.get("orientation", (0, 0, 0))), "Abs": False}}
        if "mesh" in e:
            j["StaticMesh"] = {"AssetUUID": e["mesh"], "Materials": list(e.get("materials", ()))}
        if "pointlight" in e:
            c, l, x = e["pointlight"].get("attenuation", (0.5, 0.25, 0.05))                 # LightAttenuation: constant, linear, exp
            j["Pointlight"] = {"Colour": list(e["pointlight"]["colour"]), "Attenuation": [c, l, x]}
        ents.append(j)
    doc = {"SceneUUID": scene.get("uuid", 1), "Entities": ents,
           "DirLight": {"Colour": list(scene.get("dir_colour", (1, 1, 1))), "SphericalCoords": list(scene.get("dir_spherical", (0, 0)))}}
    json.dump(doc, open(path, "w"), indent=4)


def load_scene(path: str) -> dict:
    """SceneSerializer::DeserializeScene / DeserializeEntity (SceneSerializer.cpp:90-143, :162-181). Files written before the
    `Materials` key existed (the shipped project template) get the default material, as a failed lookup does (:133-137)."""
    j = json.load(open(path))
    ents = []
    for n in j["Entities"]:
        t = n["Transform"]
        e = dict(uuid=int(n["UUID"]), name=n["Name"], parent=int(n["ParentUUID"]), pos=tuple(t["P"]), scale=tuple(t["S"]), orientation=tuple(t["O"]),
                 absolute=bool(t["Abs"]))
        if "StaticMesh" in n:
            e["mesh"] = int(n["StaticMesh"]["AssetUUID"])
            e["materials"] = [int(u) for u in n["StaticMesh"].get("Materials", [])]
        if "Pointlight" in n:
            a = n["Pointlight"]["Attenuation"]
            e["pointlight"] = dict(colour=tuple(n["Pointlight"]["Colour"]), attenuation=(a[0], a[1], a[2]))
        ents.append(e)
    return dict(uuid=int(j["SceneUUID"]), entities=ents, dir_colour=tuple(j["DirLight"]["Colour"]), dir_spherical=tuple(j["DirLight"]["SphericalCoords"]))


def instantiate_scene(scene: dict, fe, meshes: dict, materials: dict, default_material: int = 0) -> dict:
    """Scene JSON → scene front-end (snake-vk_b200/frontend.py): entities are created first, then parented and given their
    components, as DeserializeScene does. `meshes` maps StaticMeshAsset UUIDs to front-end mesh ids (with their material count),
    `materials` maps MaterialAsset UUIDs to global material slots. Returns {entity uuid: front-end entity id}."""
    ids = {e["uuid"]: fe.create_entity() for e in scene["entities"]}
    for e in scene["entities"]:
        ent = ids[e["uuid"]]
        if e["parent"] and e["parent"] in ids:
            fe.set_parent(ent, ids[e["parent"]])
        fe.set_transform(ent, e["pos"], e["orientation"], e["scale"])
        if e.get("absolute"):
            fe.set_absolute(ent, True)
        if "mesh" in e:
            mesh_id, n_mats = meshes[e["mesh"]]
            slots = [materials.get(u, default_material) for u in e.get("materials", [])]
            slots += [default_material] * (n_mats - len(slots))
            fe.set_static_mesh(ent, mesh_id, slots)
        if "pointlight" in e:
            c, l, x = e["pointlight"]["attenuation"]
            fe.set_pointlight(ent, e["pointlight"]["colour"], exp=x, linear=l, constant=c)
    fe.set_directional_light(scene["dir_colour"], *scene["dir_spherical"])
    return ids
