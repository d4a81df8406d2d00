"""Procedural scenes C1..C5 and the host-side flattening that
the reference's buffer systems perform.

Everything here is plain numpy and deterministic (seeded PCG64). The output of
`Scene.flatten()` is exactly what crosses the C ABI (include/snkpt.h): the five
global SoA mesh buffers (MeshBufferManager.cpp:22-88), `InstanceData` records
(RaytracingBufferSystem.cpp:69-103), the emissive-instance list (:105-130),
column-major transforms (TransformSystem.cpp:50-114), 64-byte materials
(ResourceBufferManagers.cpp:64-101), the light SSBO (LightBufferSystem.cpp:65-108)
and the camera UBO (SceneInfoBufferSystem.cpp:66-91, CameraComponent.h:15-19).
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field

import numpy as np

F32 = np.float32
INVALID = 0xFFFFFFFF
MAT_FLAG_EMISSIVE = 1

MATERIAL_DTYPE = np.dtype(
    [
        ("tex", "<u4", (5,)),
        ("roughness", "<f4"),
        ("metallic", "<f4"),
        ("ao", "<f4"),
        ("albedo", "<f4", (4,)),
        ("flags", "<u4"),
        ("pad", "<u4", (3,)),
    ]
)
assert MATERIAL_DTYPE.itemsize == 64

UBO_DTYPE = np.dtype(
    [
        ("view", "<f4", (16,)),
        ("proj", "<f4", (16,)),
        ("proj_view", "<f4", (16,)),
        ("cam_pos", "<f4", (4,)),
        ("cam_forward", "<f4", (4,)),
        ("frame_idx", "<u4"),
        ("time_elapsed", "<f4"),
        ("delta_time", "<f4"),
        ("prev_frame_invalid", "<u4"),
    ]
)
assert UBO_DTYPE.itemsize == 240

RESERVOIR_DTYPE = np.dtype(
    [
        ("w_sum_pl", "<f4"),
        ("w_pl", "<f4"),
        ("pl_idx", "<u4"),
        ("pl_m", "<u4"),
        ("object_instance_idx", "<u4"),
        ("triangle_idx", "<u4"),
        ("bary", "<f4", (2,)),
        ("w_sum_em", "<f4"),
        ("w_em", "<f4"),
        ("m_em", "<u4"),
        ("pad", "<u4"),
    ]
)
assert RESERVOIR_DTYPE.itemsize == 48

RAY_DTYPE = np.dtype([("o", "<f4", (3,)), ("tmin", "<f4"), ("d", "<f4", (3,)), ("tmax", "<f4")])
HIT_DTYPE = np.dtype([("t", "<f4"), ("u", "<f4"), ("v", "<f4"), ("prim", "<u4"), ("inst", "<u4")])
assert RAY_DTYPE.itemsize == 32 and HIT_DTYPE.itemsize == 20

LIGHT_HEADER_BYTES = 112
POINTLIGHT_BYTES = 48
LIGHT_TOTAL_BYTES = 112 + 1024 * 48 + 1024 * 64


# --------------------------------------------------------------------------- math
def perspective_vk(fov_deg: float, aspect: float, z_near: float, z_far: float) -> np.ndarray:
    """CameraComponent::GetProjectionMatrix (CameraComponent.h:15-19):
    glm::perspective(radians(fov*0.5), ...) RH, depth 0..1, then proj[1][1] *= -1.
    Returned column-major as a (4,4) array indexed [col][row]."""
    fovy = F32(math.radians(F32(fov_deg) * F32(0.5)))
    t = F32(math.tan(float(fovy) / 2.0))
    m = np.zeros((4, 4), F32)
    m[0][0] = F32(1.0) / (F32(aspect) * t)
    m[1][1] = -(F32(1.0) / t)
    m[2][2] = F32(z_far) / (F32(z_near) - F32(z_far))
    m[2][3] = F32(-1.0)
    m[3][2] = -(F32(z_far) * F32(z_near)) / (F32(z_far) - F32(z_near))
    return m


def look_at(eye, center, up=(0.0, 1.0, 0.0)) -> np.ndarray:
    """glm::lookAtRH, column-major [col][row]."""
    eye = np.asarray(eye, F32)
    f = np.asarray(center, F32) - eye
    f = f / np.linalg.norm(f).astype(F32)
    s = np.cross(f, np.asarray(up, F32)).astype(F32)
    s = s / np.linalg.norm(s).astype(F32)
    u = np.cross(s, f).astype(F32)
    m = np.eye(4, dtype=F32)
    m[0][0], m[1][0], m[2][0] = s
    m[0][1], m[1][1], m[2][1] = u
    m[0][2], m[1][2], m[2][2] = -f
    m[3][0] = -np.dot(s, eye)
    m[3][1] = -np.dot(u, eye)
    m[3][2] = np.dot(f, eye)
    return m.astype(F32)


def mat_mul_cm(a: np.ndarray, b: np.ndarray) -> np.ndarray:
    """a*b for column-major [col][row] arrays."""
    return (b @ a).astype(F32)  # (a*b)[c][r] = sum_k a[k][r] b[c][k]


def trs(pos=(0, 0, 0), euler_deg=(0, 0, 0), scale=(1, 1, 1)) -> np.ndarray:
    """TransformSystem::RebuildMatrix (TransformSystem.cpp:50-114) for a root
    entity: quat(radians(euler)) → mat3, columns scaled, translation."""
    e = np.radians(np.asarray(euler_deg, np.float64)) * 0.5
    c, s = np.cos(e), np.sin(e)
    w = c[0] * c[1] * c[2] + s[0] * s[1] * s[2]
    x = s[0] * c[1] * c[2] - c[0] * s[1] * s[2]
    y = c[0] * s[1] * c[2] + s[0] * c[1] * s[2]
    z = c[0] * c[1] * s[2] - s[0] * s[1] * c[2]
    m = np.eye(4)
    m[0][0] = 1 - 2 * (y * y + z * z)
    m[0][1] = 2 * (x * y + w * z)
    m[0][2] = 2 * (x * z - w * y)
    m[1][0] = 2 * (x * y - w * z)
    m[1][1] = 1 - 2 * (x * x + z * z)
    m[1][2] = 2 * (y * z + w * x)
    m[2][0] = 2 * (x * z + w * y)
    m[2][1] = 2 * (y * z - w * x)
    m[2][2] = 1 - 2 * (x * x + y * y)
    sc = np.asarray(scale, np.float64)
    for col in range(3):
        m[col][:3] *= sc[col]
    m[3][:3] = np.asarray(pos, np.float64)
    return m.astype(F32)


def transform_points(m: np.ndarray, p: np.ndarray) -> np.ndarray:
    """column-major [col][row] matrix applied to (N,3) points."""
    p = np.asarray(p, np.float64)
    m = m.astype(np.float64)
    return (p[:, 0:1] * m[0][:3] + p[:, 1:2] * m[1][:3] + p[:, 2:3] * m[2][:3] + m[3][:3]).astype(F32)


def spherical_to_cartesian(theta: float, phi: float) -> np.ndarray:
    """ExtraMath::SphericalToCartesian (ExtraMath.cpp:31-33), r = 1."""
    v = np.array(
        [math.sin(phi) * math.cos(theta), math.sin(phi) * math.sin(theta), math.cos(phi)], np.float64
    )
    return (v / np.linalg.norm(v)).astype(F32)


# --------------------------------------------------------------------------- meshes
@dataclass
class Mesh:
    positions: np.ndarray  # (V,3) f32
    normals: np.ndarray  # (V,3) f32
    uvs: np.ndarray  # (V,2) f32
    tangents: np.ndarray  # (V,3) f32
    indices: np.ndarray  # (I,) u32, mesh-local (relative to the submesh base_vertex)
    submeshes: np.ndarray  # (S,5) u32: num_indices, num_vertices, base_vertex, base_index, material_index

    @staticmethod
    def single(positions, normals, indices, uvs=None) -> "Mesh":
        positions = np.ascontiguousarray(positions, F32)
        normals = np.ascontiguousarray(normals, F32)
        indices = np.ascontiguousarray(indices, np.uint32).reshape(-1)
        v = positions.shape[0]
        uvs = np.zeros((v, 2), F32) if uvs is None else np.ascontiguousarray(uvs, F32)
        sub = np.array([[indices.size, v, 0, 0, 0]], np.uint32)
        return Mesh(positions, normals, uvs, np.zeros((v, 3), F32), indices, sub)

    @property
    def num_tris(self) -> int:
        return self.indices.size // 3


def quad_mesh(p0, p1, p2, p3) -> Mesh:
    """two triangles (p0,p1,p2),(p0,p2,p3); normal = cross(p1-p0, p3-p0)."""
    p = np.array([p0, p1, p2, p3], F32)
    n = np.cross(p[1] - p[0], p[3] - p[0])
    n = n / np.linalg.norm(n)
    uv = np.array([[0, 0], [1, 0], [1, 1], [0, 1]], F32)
    return Mesh.single(p, np.tile(n.astype(F32), (4, 1)), [0, 1, 2, 0, 2, 3], uv)


def box_mesh(lo=(-1, -1, -1), hi=(1, 1, 1), bottom=True) -> Mesh:
    """24-vertex cube like Core/res/meshes/cube.glb (12 tris), or 10 tris without the -y face."""
    lo, hi = np.asarray(lo, F32), np.asarray(hi, F32)
    faces = [
        ((1, 0, 0), (0, 1, 0), (0, 0, 1)),
        ((-1, 0, 0), (0, 0, 1), (0, 1, 0)),
        ((0, 1, 0), (0, 0, 1), (1, 0, 0)),
        ((0, -1, 0), (1, 0, 0), (0, 0, 1)),
        ((0, 0, 1), (1, 0, 0), (0, 1, 0)),
        ((0, 0, -1), (0, 1, 0), (1, 0, 0)),
    ]
    pos, nrm, uv, idx = [], [], [], []
    c, h = (lo + hi) * 0.5, (hi - lo) * 0.5
    for n, a, b in faces:
        if not bottom and n == (0, -1, 0):
            continue
        n, a, b = np.array(n, F32), np.array(a, F32), np.array(b, F32)
        base = len(pos)
        for sa, sb in ((-1, -1), (1, -1), (1, 1), (-1, 1)):
            pos.append(c + h * (n + sa * a + sb * b))
            nrm.append(n)
            uv.append(((sa + 1) / 2, (sb + 1) / 2))
        idx += [base, base + 1, base + 2, base, base + 2, base + 3]
    return Mesh.single(np.array(pos, F32), np.array(nrm, F32), idx, np.array(uv, F32))


def icosphere_mesh(subdiv: int) -> Mesh:
    """unit icosphere: 20·4^subdiv triangles (subdiv 3 → 1280, 4 → 5120)."""
    t = (1.0 + math.sqrt(5.0)) / 2.0
    v = [(-1, t, 0), (1, t, 0), (-1, -t, 0), (1, -t, 0), (0, -1, t), (0, 1, t), (0, -1, -t), (0, 1, -t),
         (t, 0, -1), (t, 0, 1), (-t, 0, -1), (-t, 0, 1)]
    verts = [np.array(p, np.float64) / np.linalg.norm(p) for p in v]
    faces = [(0, 11, 5), (0, 5, 1), (0, 1, 7), (0, 7, 10), (0, 10, 11), (1, 5, 9), (5, 11, 4), (11, 10, 2),
             (10, 7, 6), (7, 1, 8), (3, 9, 4), (3, 4, 2), (3, 2, 6), (3, 6, 8), (3, 8, 9), (4, 9, 5),
             (2, 4, 11), (6, 2, 10), (8, 6, 7), (9, 8, 1)]
    for _ in range(subdiv):
        cache: dict[tuple[int, int], int] = {}

        def mid(a: int, b: int) -> int:
            key = (a, b) if a < b else (b, a)
            if key not in cache:
                m = verts[a] + verts[b]
                verts.append(m / np.linalg.norm(m))
                cache[key] = len(verts) - 1
            return cache[key]

        nf = []
        for a, b, c in faces:
            ab, bc, ca = mid(a, b), mid(b, c), mid(c, a)
            nf += [(a, ab, ca), (b, bc, ab), (c, ca, bc), (ab, bc, ca)]
        faces = nf
    p = np.array(verts, F32)
    uv = np.stack([np.arctan2(p[:, 2], p[:, 0]) / (2 * np.pi) + 0.5, np.arccos(np.clip(p[:, 1], -1, 1)) / np.pi], 1)
    return Mesh.single(p, p.copy(), np.array(faces, np.uint32).reshape(-1), uv.astype(F32))


def merge_meshes(meshes: list[Mesh], transforms: list[np.ndarray]) -> Mesh:
    """pre-transform and concatenate into one single-submesh mesh ('flat' variants)."""
    pos, nrm, uv, idx, base = [], [], [], [], 0
    for m, t in zip(meshes, transforms):
        pos.append(transform_points(t, m.positions))
        m3 = np.array([t[0][:3], t[1][:3], t[2][:3]], np.float64)  # rows = columns of M
        nmat = np.linalg.inv(m3)  # (M^-T applied to row vector n) = n @ inv(M3cols^T)^T ...
        n = m.normals.astype(np.float64) @ nmat.T
        n /= np.maximum(np.linalg.norm(n, axis=1, keepdims=True), 1e-30)
        nrm.append(n.astype(F32))
        uv.append(m.uvs)
        idx.append(m.indices.astype(np.uint32) + np.uint32(base))
        base += m.positions.shape[0]
    return Mesh.single(np.concatenate(pos), np.concatenate(nrm), np.concatenate(idx), np.concatenate(uv))


# --------------------------------------------------------------------------- scene
@dataclass
class Entity:
    mesh: int
    transform: np.ndarray  # (4,4) col-major
    materials: list[int]  # global material index per mesh material slot
    prev_transform: np.ndarray | None = None


@dataclass
class Camera:
    pos: tuple
    forward: tuple
    fov: float = 90.0  # reference default; vertical FOV is fov/2 (CameraComponent.h:16)
    z_near: float = 0.1
    z_far: float = 2500.0


@dataclass
class Scene:
    name: str
    width: int
    height: int
    meshes: list[Mesh]
    entities: list[Entity]
    materials: np.ndarray  # MATERIAL_DTYPE
    camera: Camera
    dir_light_colour: tuple = (0.0, 0.0, 0.0)
    dir_light_spherical: tuple = (0.0, 0.0)
    pointlights: np.ndarray = field(default_factory=lambda: np.zeros((0, 12), F32))  # colour4,pos4,exp,lin,const,pad
    textures: list = field(default_factory=list)  # [(rgba8 (H, W, 4) uint8, srgb flag)], indexed by Material.tex[0]

    # ---- what the reference's buffer systems would produce
    def flatten(self) -> dict:
        pos, nrm, tan, uv, idx = [], [], [], [], []
        mesh_entries = []  # (data_start_indices_idx, data_start_vertex_idx, first_blas_id)
        v0 = i0 = blas = 0
        for m in self.meshes:
            mesh_entries.append((i0, v0, blas))
            pos.append(m.positions), nrm.append(m.normals), tan.append(m.tangents), uv.append(m.uvs)
            idx.append(m.indices)
            v0 += m.positions.shape[0]
            i0 += m.indices.size
            blas += m.submeshes.shape[0]
        inst, blas_of, transforms, prev_transforms, emissive = [], [], [], [], []
        for ti, e in enumerate(self.entities):
            m = self.meshes[e.mesh]
            si0, sv0, b0 = mesh_entries[e.mesh]
            transforms.append(e.transform.reshape(16))
            prev_transforms.append((e.prev_transform if e.prev_transform is not None else e.transform).reshape(16))
            for j, sub in enumerate(m.submeshes):
                mat = e.materials[int(sub[4])]
                if int(self.materials["flags"][mat]) & MAT_FLAG_EMISSIVE:
                    emissive.append(len(inst))
                inst.append((ti, si0 + int(sub[3]), sv0 + int(sub[2]), mat, int(sub[0])))
                blas_of.append(b0 + j)
        return dict(
            positions=np.ascontiguousarray(np.concatenate(pos), F32),
            normals=np.ascontiguousarray(np.concatenate(nrm), F32),
            tangents=np.ascontiguousarray(np.concatenate(tan), F32),
            uvs=np.ascontiguousarray(np.concatenate(uv), F32),
            indices=np.ascontiguousarray(np.concatenate(idx), np.uint32),
            mesh_entries=mesh_entries,
            instances=np.array(inst, np.uint32).reshape(-1, 5),
            blas_of_instance=np.array(blas_of, np.uint32),
            transforms=np.ascontiguousarray(np.array(transforms, F32)),
            prev_transforms=np.ascontiguousarray(np.array(prev_transforms, F32)),
            emissive_idx=np.array(emissive, np.uint32),
            materials=self.materials,
            lights=self.light_bytes(),
        )

    def light_bytes(self) -> bytes:
        n = self.pointlights.shape[0]
        buf = np.zeros(LIGHT_HEADER_BYTES // 4 + max(n, 1) * 12, F32)
        buf[0:3] = self.dir_light_colour
        buf[3] = 1.0
        buf[4:7] = spherical_to_cartesian(*self.dir_light_spherical)
        buf[7] = 1.0
        # light_transform (ortho*view) is only used by the raster shadow pass; identity here
        buf[8:24] = np.eye(4, dtype=F32).reshape(16)
        buf.view(np.uint32)[24] = n
        buf.view(np.uint32)[25] = 0
        if n:
            buf[28:28 + n * 12] = self.pointlights.reshape(-1)
        return buf.tobytes()

    def ubo(self, frame_idx: int = 0, cam_pos=None, prev_invalid: bool | None = None) -> np.ndarray:
        """SceneInfoBufferSystem::UpdateUBO; time_elapsed = 1000·frame/60 ms, delta 16.6667 ms."""
        c = self.camera
        pos = np.asarray(c.pos if cam_pos is None else cam_pos, F32)
        fwd = np.asarray(c.forward, F32)
        fwd = fwd / np.linalg.norm(fwd)
        u = np.zeros((), UBO_DTYPE)
        view = look_at(pos, pos + fwd)
        proj = perspective_vk(c.fov, self.width / self.height, c.z_near, c.z_far)
        u["view"] = view.reshape(16)
        u["proj"] = proj.reshape(16)
        u["proj_view"] = mat_mul_cm(proj, view).reshape(16)
        u["cam_pos"] = (*pos, 1.0)
        u["cam_forward"] = (*fwd, 1.0)
        u["frame_idx"] = frame_idx
        u["time_elapsed"] = F32(1000.0 * frame_idx / 60.0)
        u["delta_time"] = F32(16.6667)
        u["prev_frame_invalid"] = int(frame_idx == 0 if prev_invalid is None else prev_invalid)
        return u

    @property
    def num_tris(self) -> int:
        return sum(self.meshes[e.mesh].num_tris for e in self.entities)


def make_materials(specs) -> np.ndarray:
    """specs: list of (albedo rgb, roughness, metallic, flags)."""
    m = np.zeros(len(specs), MATERIAL_DTYPE)
    m["tex"] = INVALID
    m["ao"] = 1.0
    for i, (alb, r, me, fl) in enumerate(specs):
        m["albedo"][i] = (*alb, 1.0)
        m["roughness"][i] = r
        m["metallic"][i] = me
        m["flags"][i] = fl
    return m


def default_pointlights(rng, n, lo, hi) -> np.ndarray:
    """PointlightComponent defaults (Lights.h:6-10): attenuation const .5, linear .25, exp .05."""
    pl = np.zeros((n, 12), F32)
    pl[:, 0:3] = rng.uniform(0.5, 1.0, (n, 3)) * 10.0
    pl[:, 3] = 1.0
    pl[:, 4:7] = rng.uniform(lo, hi, (n, 3))
    pl[:, 7] = 1.0
    pl[:, 8], pl[:, 9], pl[:, 10], pl[:, 11] = 0.05, 0.25, 0.5, 1.0
    return pl


# ---- C1 ------------------------------------------------------------------------------
def cornell(width=512, height=512, with_pointlight=False) -> Scene:
    """C1: 5 walls × 2 + ceiling quad light × 2 + 2 open boxes × 10 = 32 tris, 8 instances."""
    s = 5.5
    meshes = [
        quad_mesh((0, 0, 0), (0, 0, s), (s, 0, s), (s, 0, 0)),  # floor, +y normal
        quad_mesh((0, s, 0), (s, s, 0), (s, s, s), (0, s, s)),  # ceiling, -y
        quad_mesh((0, 0, s), (0, s, s), (s, s, s), (s, 0, s)),  # back, -z
        quad_mesh((0, 0, 0), (0, s, 0), (0, s, s), (0, 0, s)),  # left (red), +x
        quad_mesh((s, 0, 0), (s, 0, s), (s, s, s), (s, s, 0)),  # right (green), -x
        quad_mesh((-0.65, 0, -0.525), (0.65, 0, -0.525), (0.65, 0, 0.525), (-0.65, 0, 0.525)),  # light, -y
        box_mesh((-0.5, 0, -0.5), (0.5, 1, 0.5), bottom=False),
    ]
    mats = make_materials([
        ((0.73, 0.73, 0.73), 0.5, 0.0, 0),
        ((0.65, 0.05, 0.05), 0.5, 0.0, 0),
        ((0.12, 0.45, 0.15), 0.5, 0.0, 0),
        ((17.0, 12.0, 4.0), 0.5, 0.0, MAT_FLAG_EMISSIVE),
        ((0.73, 0.73, 0.73), 0.3, 0.5, 0),
    ])
    ident = np.eye(4, dtype=F32)
    ents = [
        Entity(0, ident, [0]), Entity(1, ident, [0]), Entity(2, ident, [0]), Entity(3, ident, [1]),
        Entity(4, ident, [2]),
        Entity(5, trs((2.78, s - 0.01, 2.795)), [3]),
        Entity(6, trs((3.7, 0.0, 1.7), (0, 17.0, 0), (1.65, 1.65, 1.65)), [0]),
        Entity(6, trs((1.9, 0.0, 3.5), (0, -18.0, 0), (1.65, 3.3, 1.65)), [4]),
    ]
    sc = Scene("C1-cornell", width, height, meshes, ents, mats, Camera((2.75, 2.75, -8.0), (0, 0, 1)))
    if with_pointlight:
        pl = np.zeros((2, 12), F32)
        pl[0] = (6, 6, 5, 1, 1.2, 4.2, 1.5, 1, 0.05, 0.25, 0.5, 1)
        pl[1] = (2, 4, 8, 1, 4.4, 2.0, 0.8, 1, 0.05, 0.25, 0.5, 1)
        sc.pointlights = pl
        sc.dir_light_colour = (0.6, 0.6, 0.5)
        sc.dir_light_spherical = (0.8, 0.9)
    return sc


# ---- C2 / C3 / C4 ---------------------------------------------------------------------------
_PALETTE = [(0.8, 0.8, 0.8), (0.8, 0.3, 0.25), (0.25, 0.7, 0.3), (0.3, 0.35, 0.8), (0.8, 0.75, 0.3),
            (0.6, 0.3, 0.7), (0.3, 0.7, 0.75), (0.9, 0.55, 0.25)]


def sphere_field(n_spheres: int, subdiv: int, box=(60.0, 12.0, 30.0), seed=0x5EED0002, width=1920, height=1080,
                 variant="instanced", n_emissive_cubes=0, n_pointlights=16, name="C2") -> Scene:
    """C2/C4 generator: floor quad + n icospheres, random pose in `box` centred on (0, box.y/2, 0)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    bx, by, bz = box
    sphere = icosphere_mesh(subdiv)
    floor = quad_mesh((-bx, 0, -bz), (-bx, 0, bz), (bx, 0, bz), (bx, 0, -bz))
    specs = [(c, 0.5, 0.0, 0) for c in _PALETTE]
    specs += [((20.0, 0, 0), 0.5, 0.0, MAT_FLAG_EMISSIVE), ((0, 20.0, 0), 0.5, 0.0, MAT_FLAG_EMISSIVE),
              ((0, 0, 20.0), 0.5, 0.0, MAT_FLAG_EMISSIVE)]
    mats = make_materials(specs)
    pos = rng.uniform((-bx / 2, 0.5, -bz / 2), (bx / 2, by, bz / 2), (n_spheres, 3))
    scale = rng.uniform(0.5, 2.0, n_spheres)
    euler = rng.uniform(0, 360.0, (n_spheres, 3))
    mat_id = rng.integers(1, len(_PALETTE), n_spheres)
    xforms = [trs(pos[i], euler[i], (scale[i],) * 3) for i in range(n_spheres)]
    ident = np.eye(4, dtype=F32)
    if variant == "instanced":
        meshes = [floor, sphere]
        ents = [Entity(0, ident, [0])] + [Entity(1, xforms[i], [int(mat_id[i])]) for i in range(n_spheres)]
    elif variant == "flat":
        meshes = [merge_meshes([floor] + [sphere] * n_spheres, [ident] + xforms)]
        ents = [Entity(0, ident, [0])]
    else:
        raise ValueError(variant)
    if n_emissive_cubes:
        meshes.append(box_mesh())
        cube = len(meshes) - 1
        cpos = rng.uniform((-bx / 2, 0.5, -bz / 2), (bx / 2, by, bz / 2), (n_emissive_cubes, 3))
        ccol = rng.integers(0, 3, n_emissive_cubes)
        for i in range(n_emissive_cubes):
            ents.append(Entity(cube, trs(cpos[i], (0, 0, 0), (0.1, 0.1, 0.1)), [len(_PALETTE) + int(ccol[i])]))
    sc = Scene(f"{name}-{variant}", width, height, meshes, ents, mats,
               Camera((0.0, by * 0.45, -bz * 0.5 - 8.0), (0.0, -0.05, 1.0)))
    sc.dir_light_colour = (1.0, 1.0, 1.0)
    sc.dir_light_spherical = (0.8, 0.9)
    if n_pointlights:
        sc.pointlights = default_pointlights(rng, n_pointlights, (-bx / 2, 1.0, -bz / 2), (bx / 2, by, bz / 2))
    return sc


def shipped_field(sphere: Mesh, cube: Mesh, n=600, box=(60.0, 12.0, 30.0), seed=0x5EED0006, width=1920, height=1080) -> Scene:
    """the meshes the engine ships (Core/res/meshes/sphere.glb: 960 triangles, cube.glb: 12) instanced `n` times over a cube floor,
    one emissive cube lamp per 32 instances, 16 point lights (north_star: throughput on any mesh shipped in Core/res)"""
    rng = np.random.Generator(np.random.PCG64(seed))
    bx, by, bz = box
    specs = [(c, 0.5, 0.0, 0) for c in _PALETTE] + [((20.0, 18.0, 14.0), 0.5, 0.0, MAT_FLAG_EMISSIVE)]
    mats = make_materials(specs)
    ents = [Entity(1, trs((0.0, -1.0, 0.0), (0, 0, 0), (bx, 1.0, bz)), [0])]
    pos = rng.uniform((-bx / 2, 0.5, -bz / 2), (bx / 2, by, bz / 2), (n, 3))
    scale = rng.uniform(0.5, 2.0, n)
    euler = rng.uniform(0, 360.0, (n, 3))
    mat_id = rng.integers(1, len(_PALETTE), n)
    for i in range(n):
        if i % 32 == 31:
            ents.append(Entity(1, trs(pos[i], euler[i], (0.15,) * 3), [len(_PALETTE)]))
        else:
            ents.append(Entity(0 if i % 4 else 1, trs(pos[i], euler[i], (scale[i],) * 3 if i % 4 else (0.6 * scale[i],) * 3), [int(mat_id[i])]))
    sc = Scene("shipped-field", width, height, [sphere, cube], ents, mats, Camera((0.0, by * 0.45, -bz * 0.5 - 8.0), (0.0, -0.05, 1.0)))
    sc.dir_light_colour = (1.0, 1.0, 1.0)
    sc.dir_light_spherical = (0.8, 0.9)
    sc.pointlights = default_pointlights(rng, 16, (-bx / 2, 1.0, -bz / 2), (bx / 2, by, bz / 2))
    return sc


def c2(variant="instanced", subdiv=3, width=1920, height=1080) -> Scene:
    """C2 Sponza-scale: 204 icospheres (1280 tris) + floor = 261,122 tris; subdiv 4 → 1.04 M."""
    return sphere_field(204, subdiv, variant=variant, width=width, height=height, name="C2")


def c3(width=1920, height=1080) -> Scene:
    """C3: C2 instanced + 854 emissive cubes × 12 = 10,248 emissive triangles."""
    return sphere_field(204, 3, variant="instanced", n_emissive_cubes=854, width=width, height=height,
                        seed=0x5EED0003, name="C3")


def c4(variant="instanced", width=3840, height=2160) -> Scene:
    """C4 Bistro-scale: 2343 icospheres × 1280 = 2,999,040 tris + floor."""
    return sphere_field(2343, 3, box=(160.0, 20.0, 80.0), variant=variant, width=width, height=height,
                        seed=0x5EED0004, name="C4")


def c5_soup(n_tris: int, seed=0x5EED0005) -> Mesh:
    """C5: flat soup of ~n_tris triangles made of icosphere clusters."""
    rng = np.random.Generator(np.random.PCG64(seed))
    sphere = icosphere_mesh(3)
    n = max(1, n_tris // sphere.num_tris)
    side = max(4.0, float(n) ** (1 / 3) * 3.0)
    pos = rng.uniform(-side, side, (n, 3))
    scale = rng.uniform(0.5, 2.0, n)
    euler = rng.uniform(0, 360.0, (n, 3))
    return merge_meshes([sphere] * n, [trs(pos[i], euler[i], (scale[i],) * 3) for i in range(n)])


def textured_test_scene(width=96, height=64, seed=11) -> Scene:
    """small_test_scene with albedo textures: planar texcoords that leave [0,1] (repeat addressing), one sRGB
    power-of-two texture and one UNORM texture of odd size (RayClosestHit.rchit:100-116, GBuffer.frag:56-58)."""
    sc = small_test_scene(width, height, seed=seed)
    rng = np.random.default_rng(seed)
    for m in sc.meshes:
        p = m.positions
        m.uvs = np.ascontiguousarray(np.stack([p[:, 0] * 0.37 + p[:, 2] * 0.11 + 0.25, p[:, 1] * 0.29 - p[:, 2] * 0.23 + 0.5], -1), F32)
    t0 = rng.integers(0, 256, (8, 16, 4), dtype=np.uint8)
    yy, xx = np.mgrid[0:8, 0:16]
    t0[(xx // 2 + yy // 2) % 2 == 0, :3] //= 4  # checker: strong contrast so that filtering errors show
    t1 = rng.integers(0, 256, (7, 5, 4), dtype=np.uint8)
    # tangent-space normal map (UNORM): perturbations around +z, so that the mapped normal stays in the upper hemisphere
    t2 = np.zeros((8, 8, 4), np.uint8)
    t2[..., 0] = rng.integers(88, 168, (8, 8)); t2[..., 1] = rng.integers(88, 168, (8, 8)); t2[..., 2] = 235; t2[..., 3] = 255
    sc.textures = [(t0, 1), (t1, 0), (t2, 0)]
    for m in sc.meshes:  # tangents: any unit vector perpendicular-ish to the normal (GBuffer.vert orthonormalises it)
        ref = np.where(np.abs(m.normals[:, 1:2]) < 0.9, np.float32([[0, 1, 0]]), np.float32([[1, 0, 0]]))
        t = np.cross(ref, m.normals)
        m.tangents = np.ascontiguousarray(t / np.linalg.norm(t, axis=1, keepdims=True), F32)
    mats = sc.materials.copy()
    for i in range(mats.shape[0]):
        if not (int(mats["flags"][i]) & MAT_FLAG_EMISSIVE or i % 3 == 2):
            mats["tex"][i, 0] = i % 2
        if i % 2 == 0 and not int(mats["flags"][i]) & MAT_FLAG_EMISSIVE:
            mats["tex"][i, 1] = 2  # normal_tex_idx
    mats["tex"][mats.shape[0] - 1, 0] = 1  # also on the last material whatever it is (an emissive one keeps its plain albedo as emission)
    sc.materials = mats
    return sc


def small_test_scene(width=96, height=64, seed=7, n_spheres=6, with_emissive=True) -> Scene:
    """tiny instanced scene for CPU-fast parity tests (few hundred triangles)."""
    sc = sphere_field(n_spheres, 1, box=(12.0, 5.0, 8.0), seed=seed, width=width, height=height,
                      n_emissive_cubes=3 if with_emissive else 0, n_pointlights=3, name="T")
    return sc
