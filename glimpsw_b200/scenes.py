"""Procedural meshlet scenes for the BASELINE.json configs (test harness + bench input).

The reference builds meshlets with meshoptimizer at glTF import (src/SwRast/Scene.cpp:193-289);
neither meshoptimizer nor the full Sponza asset is available, so the benchmark scenes are
generated here, deterministically (splitmix64 seeds), straight into the reference's `Meshlet`
SoA layout (Scene.h:15-30). A different meshletisation changes triangle ids, not correctness.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field

import numpy as np

from . import camera as cam
from .layout import MESHLET_DTYPE, MATERIAL_DTYPE, LIGHT_DTYPE, NO_MATERIAL

f32 = np.float32


# ---------------------------------------------------------------------------------------------
# splitmix64 (the generator the reference's benchmarks use, Benchmarks/GatherThroughput.cpp:75-80)
def splitmix64(seed: int, n: int) -> np.ndarray:
    with np.errstate(over="ignore"):
        x = np.uint64(seed) + np.arange(1, n + 1, dtype=np.uint64) * np.uint64(0x9E3779B97F4A7C15)
        z = x
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        return z ^ (z >> np.uint64(31))


def rand01(seed: int, n: int) -> np.ndarray:
    return (splitmix64(seed, n) >> np.uint64(11)).astype(np.float64) * (1.0 / (1 << 53))


# ---------------------------------------------------------------------------------------------
# vertex attribute packing (Scene.cpp:251-275: fp16 UV pairs, oct-encoded normal+tangent as 4 x unorm8)
def pack_half2(u: np.ndarray, v: np.ndarray) -> np.ndarray:
    lo = np.asarray(u, dtype=np.float16).view(np.uint16).astype(np.uint32)
    hi = np.asarray(v, dtype=np.float16).view(np.uint16).astype(np.uint32)
    return lo | (hi << 16)


def _oct_encode(n: np.ndarray) -> np.ndarray:
    n = n / np.maximum(np.abs(n).sum(axis=-1, keepdims=True), 1e-20)
    xy = n[..., :2].copy()
    neg = n[..., 2] < 0
    folded = (1.0 - np.abs(xy[..., ::-1])) * np.where(xy >= 0, 1.0, -1.0)
    xy[neg] = folded[neg]
    return xy * 0.5 + 0.5


def pack_normal_tangent(normal: np.ndarray, tangent: np.ndarray) -> np.ndarray:
    n = np.clip(np.rint(_oct_encode(normal) * 255.0), 0, 255).astype(np.uint32)
    t = np.clip(np.rint(_oct_encode(tangent) * 255.0), 0, 255).astype(np.uint32)
    return n[..., 0] | (n[..., 1] << 8) | (t[..., 0] << 16) | (t[..., 1] << 24)


# ---------------------------------------------------------------------------------------------
@dataclass
class DrawNode:
    """One glTF node = one DrawMeshlets call (Main.cpp:216-240)."""
    meshlet_offset: int
    meshlet_count: int
    model: np.ndarray  # (4,4)[c,r] float32


@dataclass
class SceneData:
    name: str
    meshlets: np.ndarray                       # MESHLET_DTYPE
    nodes: list                                # list[DrawNode]
    camera: cam.Camera
    width: int
    height: int
    materials: np.ndarray = field(default_factory=lambda: np.zeros(0, dtype=MATERIAL_DTYPE))
    textures: list = field(default_factory=list)   # list[TextureData]
    lights: np.ndarray = field(default_factory=lambda: np.zeros(0, dtype=LIGHT_DTYPE))

    @property
    def num_triangles(self) -> int:
        return int(self.meshlets["NumTriangles"].astype(np.int64).sum())

    def view_proj(self):
        c = self.camera
        c.aspect = self.width / self.height
        return c.proj_matrix(), c.view_matrix(True)

    def object_to_clip(self, node: DrawNode) -> np.ndarray:
        p, v = self.view_proj()
        return cam.object_to_clip(p, v, node.model)


def concat_meshlets(parts) -> np.ndarray:
    """np.concatenate re-packs padded structured dtypes (1728 -> 1723 bytes); copy into a fresh array instead."""
    out = np.zeros(sum(len(p) for p in parts), dtype=MESHLET_DTYPE)
    k = 0
    for p in parts:
        assert p.dtype == MESHLET_DTYPE
        out[k:k + len(p)] = p
        k += len(p)
    return out


def set_bounds(meshlets: np.ndarray) -> None:
    """Bounding sphere per meshlet (Scene.cpp:239-241 stores meshopt's; here: bbox centre + max distance)."""
    nv = meshlets["NumVertices"].astype(np.int64)
    pos = meshlets["Positions"]                       # [M,3,64]
    valid = np.arange(64)[None, :] < nv[:, None]      # [M,64]
    big = np.where(valid[:, None, :], pos, np.inf)
    small = np.where(valid[:, None, :], pos, -np.inf)
    lo, hi = big.min(axis=2), small.max(axis=2)
    centre = ((lo + hi) * 0.5).astype(f32)
    d = pos - centre[:, :, None]
    dist2 = np.where(valid, (d.astype(np.float64) ** 2).sum(axis=1), 0.0)
    meshlets["BoundCenter"] = centre
    meshlets["BoundRadius"] = (np.sqrt(dist2.max(axis=1)) * (1 + 1e-6)).astype(f32)


def meshletize(positions: np.ndarray, tris: np.ndarray, uv=None, normals=None, tangents=None,
               material_id: int = NO_MATERIAL, alpha_cutoff: int = 255,
               max_verts: int = 64, max_tris: int = 128) -> np.ndarray:
    """Greedy linear-scan meshletizer (<=64 unique vertices, <=128 triangles per meshlet).

    Stand-in for meshopt_buildMeshlets(…, 64, 128, 0.25) (Scene.cpp:199-237)."""
    tris = np.asarray(tris, dtype=np.int64)
    groups = []           # (vertex id list, local index array)
    remap: dict = {}
    verts: list = []
    local: list = []
    for a, b, c in tris.tolist():
        new = [v for v in (a, b, c) if v not in remap]
        new = list(dict.fromkeys(new))
        if len(local) >= max_tris or len(verts) + len(new) > max_verts:
            groups.append((verts, local))
            remap, verts, local = {}, [], []
            new = list(dict.fromkeys((a, b, c)))
        for v in new:
            remap[v] = len(verts)
            verts.append(v)
        local.append((remap[a], remap[b], remap[c]))
    if local:
        groups.append((verts, local))

    out = np.zeros(len(groups), dtype=MESHLET_DTYPE)
    out["MaterialId"] = material_id
    out["AlphaCutoff"] = alpha_cutoff
    for i, (vs, ls) in enumerate(groups):
        vs = np.asarray(vs, dtype=np.int64)
        ls = np.asarray(ls, dtype=np.uint8)
        m = out[i]
        m["NumVertices"], m["NumTriangles"] = len(vs), len(ls)
        m["Positions"][:, : len(vs)] = positions[vs].T.astype(f32)
        m["Indices"][:, : len(ls)] = ls.T
        if uv is not None:
            m["TexCoords"][: len(vs)] = pack_half2(uv[vs, 0], uv[vs, 1])
        if normals is not None:
            t = tangents[vs, :3] if tangents is not None else np.roll(normals[vs], 1, axis=1)
            m["NormalTangents"][: len(vs)] = pack_normal_tangent(normals[vs], t)
            if tangents is not None and tangents.shape[1] == 4:
                bits = (tangents[vs, 3] < 0).astype(np.uint64) << np.arange(len(vs), dtype=np.uint64)
                m["TangentHandedness"] = np.bitwise_or.reduce(bits)
    set_bounds(out)
    return out


# ---------------------------------------------------------------------------------------------
def _patch_indices(n: int = 8) -> np.ndarray:
    """Triangle list of an n x n vertex patch: (n-1)^2 * 2 triangles, CCW when seen with +row down-screen."""
    idx = []
    for r in range(n - 1):
        for c in range(n - 1):
            v00, v01, v10, v11 = r * n + c, r * n + c + 1, (r + 1) * n + c, (r + 1) * n + c + 1
            idx.append((v00, v10, v01))
            idx.append((v01, v10, v11))
    return np.asarray(idx, dtype=np.uint8)


def _value_noise(x: np.ndarray, y: np.ndarray, seed: int, cells: int = 64) -> np.ndarray:
    lattice = rand01(seed, (cells + 1) * (cells + 1)).reshape(cells + 1, cells + 1)
    fx, fy = np.clip(x, 0, 1) * (cells - 1e-9), np.clip(y, 0, 1) * (cells - 1e-9)
    ix, iy = fx.astype(np.int64), fy.astype(np.int64)
    tx, ty = fx - ix, fy - iy
    tx, ty = tx * tx * (3 - 2 * tx), ty * ty * (3 - 2 * ty)
    a = lattice[iy, ix] * (1 - tx) + lattice[iy, ix + 1] * tx
    b = lattice[iy + 1, ix] * (1 - tx) + lattice[iy + 1, ix + 1] * tx
    return a * (1 - ty) + b * ty


def grid_scene(patches_x: int = 102, patches_y: int = 100, width: int = 1920, height: int = 1080,
               seed: int = 1, material_id: int = NO_MATERIAL, flip_fraction: float = 0.0) -> SceneData:
    """BASELINE config C2: heightfield grid of patches_x*patches_y meshlets, each an 8x8-vertex patch
    (98 triangles) -> 999,600 triangles at the default size. The grid is laid out as a fan in front of an
    oblique camera so every vertex is inside the frustum and triangle footprints span ~0.1 to ~10 px^2
    (row heights grow geometrically towards the bottom of the screen)."""
    n = 8
    cols, rows = patches_x * (n - 1) + 1, patches_y * (n - 1) + 1
    camera = cam.Camera(position=(3.0, 6.0, 2.0), euler=(0.35, -0.25), fov_deg=90.0, aspect=width / height)
    proj, view = camera.proj_matrix(), camera.view_matrix(True)

    s = np.linspace(0.0, 1.0, cols)
    g = np.exp(np.linspace(0.0, math.log(100.0), rows - 1))
    t = np.concatenate([[0.0], np.cumsum(g)]) / g.sum()
    S, T = np.meshgrid(s, t)                                       # [rows, cols]
    sx = (S - 0.5) * 1.8
    sy = -0.9 + 1.8 * T                                            # top of screen (far) -> bottom (near)
    depth = 60.0 * (1.0 - T) ** 2 + 0.6                            # far rows deeper
    bump = (np.sin(S * 37.0) * np.cos(T * 53.0 + S * 11.0) * 0.5 + _value_noise(S, T, seed) - 0.5)
    depth = depth * (1.0 + 0.08 * bump)
    sy = sy + 0.04 * np.sin(S * 91.0 + T * 29.0) * (0.2 + T)
    f = 1.0 / math.tan(math.radians(camera.fov) / 2.0)
    xv = sx * depth * camera.aspect / f
    yv = -sy * depth / f
    zv = -depth
    pv = np.stack([xv, yv, zv, np.ones_like(xv)], axis=-1)         # view-space points

    model = cam.mat_mul(cam.translate((1.5, -0.75, 0.25)), cam.scale(0.5))
    # object = inverse(view * model) * view-space
    vm = cam.mat_mul(view, model).astype(np.float64).T            # as a standard matrix
    obj = (np.linalg.inv(vm) @ pv.reshape(-1, 4).T).T.reshape(rows, cols, 4)[..., :3]

    meshlets = np.zeros(patches_x * patches_y, dtype=MESHLET_DTYPE)
    meshlets["NumVertices"], meshlets["NumTriangles"] = n * n, (n - 1) * (n - 1) * 2
    meshlets["AlphaCutoff"], meshlets["MaterialId"] = 255, material_id
    py, px = np.meshgrid(np.arange(patches_y), np.arange(patches_x), indexing="ij")
    r0, c0 = (py * (n - 1)).reshape(-1), (px * (n - 1)).reshape(-1)
    lr, lc = np.meshgrid(np.arange(n), np.arange(n), indexing="ij")
    rr = r0[:, None] + lr.reshape(1, -1)
    cc = c0[:, None] + lc.reshape(1, -1)
    meshlets["Positions"] = np.transpose(obj[rr, cc], (0, 2, 1)).astype(f32)
    meshlets["TexCoords"] = pack_half2(S[rr, cc] * 8.0, T[rr, cc] * 8.0)
    up = np.zeros((len(meshlets), n * n, 3)); up[..., 1] = 1.0
    tg = np.zeros_like(up); tg[..., 0] = 1.0
    meshlets["NormalTangents"] = pack_normal_tangent(up, tg)
    idx = _patch_indices(n)
    ind = np.zeros((3, 128), dtype=np.uint8)
    ind[:, : len(idx)] = idx.T
    meshlets["Indices"] = ind[None]
    if flip_fraction > 0:   # reverse the winding of a seeded subset -> back-face culled
        flip = rand01(seed + 7, len(meshlets)) < flip_fraction
        sw = meshlets["Indices"][flip]
        sw[:, [1, 2], :] = sw[:, [2, 1], :]
        meshlets["Indices"][flip] = sw
    set_bounds(meshlets)
    return SceneData(f"grid{patches_x}x{patches_y}", meshlets, [DrawNode(0, len(meshlets), model)], camera, width, height)


# ---------------------------------------------------------------------------------------------
def icosphere(subdivisions: int):
    """Unit icosphere: 20 * 4^s triangles; children of a triangle stay consecutive (good meshlet locality)."""
    p = (1.0 + math.sqrt(5.0)) / 2.0
    v = np.array([[-1, p, 0], [1, p, 0], [-1, -p, 0], [1, -p, 0], [0, -1, p], [0, 1, p], [0, -1, -p], [0, 1, -p],
                  [p, 0, -1], [p, 0, 1], [-p, 0, -1], [-p, 0, 1]], dtype=np.float64)
    v /= np.linalg.norm(v, axis=1, keepdims=True)
    t = np.array([[0, 11, 5], [0, 5, 1], [0, 1, 7], [0, 7, 10], [0, 10, 11], [1, 5, 9], [5, 11, 4], [11, 10, 2],
                  [10, 7, 6], [7, 1, 8], [3, 9, 4], [3, 4, 2], [3, 2, 6], [3, 6, 8], [3, 8, 9], [4, 9, 5],
                  [2, 4, 11], [6, 2, 10], [8, 6, 7], [9, 8, 1]], dtype=np.int64)
    for _ in range(subdivisions):
        e = np.concatenate([t[:, [0, 1]], t[:, [1, 2]], t[:, [2, 0]]], axis=0)
        es = np.sort(e, axis=1)
        uniq, inv = np.unique(es, axis=0, return_inverse=True)
        mid = v[uniq[:, 0]] + v[uniq[:, 1]]
        mid /= np.linalg.norm(mid, axis=1, keepdims=True)
        base = len(v)
        v = np.concatenate([v, mid], axis=0)
        nt = len(t)
        a, b, c = base + inv[:nt].reshape(-1), base + inv[nt:2 * nt].reshape(-1), base + inv[2 * nt:].reshape(-1)
        t = np.stack([np.stack([t[:, 0], a, c], 1), np.stack([t[:, 1], b, a], 1),
                      np.stack([t[:, 2], c, b], 1), np.stack([a, b, c], 1)], axis=1).reshape(-1, 3)
    return v, t


def instanced_scene(subdivisions: int = 6, instances: int = 122, width: int = 1920, height: int = 1080,
                    seed: int = 4, bake_transforms: bool = False) -> SceneData:
    """BASELINE config C4: noise-displaced icosphere (81,920 triangles at 6 subdivisions) instanced on a
    jittered 3-D lattice with per-instance rigid transforms (122 instances ~ 9.99 M triangles). The camera
    sits inside the lattice: most meshlets fail the frustum test and about half of the rest face away.

    bake_transforms: the instance transforms are applied to the meshlets on the host (positions, normals, tangents,
    bounds) and every node's model matrix is the identity. Same triangles, same 122 DrawMeshlets calls, but ONE
    ObjectToClip for the whole frame — which is what ShadingContext::Resolve assumes (it transforms every pixel with the
    context's current matrix, Shading.cpp:509-511; README.md:30), so the resolved image of the multi-node scene is
    meaningful and its parity with the oracle well defined."""
    v, t = icosphere(subdivisions)
    disp = (np.sin(v[:, 0] * 9.0) * np.sin(v[:, 1] * 7.0 + 1.0) * np.sin(v[:, 2] * 8.0 + 2.0)) * 0.06
    disp += (np.sin(v[:, 0] * 31.0 + v[:, 1] * 17.0) * 0.015)
    v = v * (1.0 + disp)[:, None]
    n = v / np.linalg.norm(v, axis=1, keepdims=True)
    uv = np.stack([np.arctan2(n[:, 2], n[:, 0]) / (2 * math.pi) + 0.5, np.arccos(np.clip(n[:, 1], -1, 1)) / math.pi], 1) * 4.0
    # winding: make outward faces CCW in this renderer's screen space (y down): flip to match det>0 for front faces
    base = meshletize(v, t[:, [0, 2, 1]], uv=uv, normals=n)
    nb = len(base)

    side = int(math.ceil(instances ** (1.0 / 3.0)))
    r = rand01(seed, instances * 8).reshape(instances, 8)
    meshlets = np.tile(base, instances)
    nodes = []
    k = 0
    spacing = 3.2
    for i in range(instances):
        gx, gy, gz = i % side, (i // side) % side, i // (side * side)
        centre = (np.array([gx, gy, gz], dtype=np.float64) - (side - 1) / 2.0) * spacing + (r[i, 0:3] - 0.5) * 1.2
        axis = r[i, 3:6] - 0.5 + 1e-3
        model = cam.mat_mul(cam.translate(centre), cam.mat_mul(cam.rotate_axis(axis, r[i, 6] * 2 * math.pi), cam.scale(0.9 + 0.4 * r[i, 7])))
        if bake_transforms:
            A = model.astype(np.float64).T                                   # standard 4x4 (model is stored [c, r])
            part = meshlets[k:k + nb]
            P = base["Positions"].astype(np.float64)                          # [nb, 3, 64]
            part["Positions"] = (np.einsum("rc,mcv->mrv", A[:3, :3], P) + A[:3, 3][None, :, None]).astype(f32)
            # the displacement is radial, so the object-space normal is the normalised position; any unit vector
            # perpendicular to it serves as tangent
            nrm = P / np.maximum(np.linalg.norm(P, axis=1, keepdims=True), 1e-20)
            tan = np.cross(np.broadcast_to(np.array([0.0, 1.0, 0.0]), np.moveaxis(nrm, 1, 2).shape), np.moveaxis(nrm, 1, 2))
            tan = tan / np.maximum(np.linalg.norm(tan, axis=2, keepdims=True), 1e-9)
            Rm = A[:3, :3] / np.linalg.norm(A[:3, 0])
            part["NormalTangents"] = pack_normal_tangent(np.moveaxis(nrm, 1, 2) @ Rm.T, tan @ Rm.T)
            model = cam.identity()
        nodes.append(DrawNode(k, nb, model))
        k += nb
    if bake_transforms:
        set_bounds(meshlets)
    camera = cam.Camera(position=(0.4, 0.3, 1.6), euler=(0.6, 0.15), fov_deg=90.0, aspect=width / height)
    return SceneData(f"icosphere{subdivisions}x{instances}" + ("_baked" if bake_transforms else ""), meshlets, nodes, camera, width, height)


def orbit_cameras(scene: SceneData, count: int, seed: int = 5, radius: float = 6.0):
    """BASELINE config C5: `count` seeded cameras orbiting the scene origin, all looking inwards-ish."""
    r = rand01(seed, count * 3).reshape(count, 3)
    cams = []
    for i in range(count):
        ang = 2 * math.pi * (i / count) + (r[i, 0] - 0.5) * 0.2
        rad = radius * (0.6 + 0.5 * r[i, 1])
        pos = (math.sin(ang) * rad, (r[i, 2] - 0.5) * 3.0, math.cos(ang) * rad)
        cams.append(cam.Camera(position=pos, euler=(-ang, -0.1), fov_deg=90.0, aspect=scene.width / scene.height))
    return cams


# ---------------------------------------------------------------------------------------------
def default_light() -> np.ndarray:
    """The directional light Playground installs when a scene has none (Main.cpp:192-201)."""
    l = np.zeros(1, dtype=LIGHT_DTYPE)
    d = -np.array([0.589494, 0.684509, 0.428906])
    l["Type"] = 0
    l["Direction"] = (d / np.linalg.norm(d)).astype(f32)
    l["Color"] = (1.0, 1.0, 1.0)
    l["Intensity"] = 1500.0
    return l


def make_light(kind: int, position=(0, 0, 0), direction=(0, -1, 0), color=(1, 1, 1), intensity=1500.0, radius=0.0,
               inner=0.3, outer=0.6) -> np.ndarray:
    """Light with the pre-computed fields of Light::SetRadius / SetSpotAngles (Scene.h:66-75)."""
    l = np.zeros(1, dtype=LIGHT_DTYPE)
    d = np.asarray(direction, dtype=np.float64)
    l["Type"], l["Position"], l["Direction"] = kind, position, (d / np.linalg.norm(d)).astype(f32)
    l["Color"], l["Intensity"] = color, intensity
    rad = radius if radius > 0 else 1e6
    l["Radius"] = rad
    l["InvRadiusSq"] = f32(1.0) / (f32(rad) * f32(rad))
    l["SpotInnerAngle"], l["SpotOuterAngle"] = inner, outer
    sc = f32(1.0) / max(f32(math.cos(inner)) - f32(math.cos(outer)), f32(1e-4))
    l["SpotScale"] = sc
    l["SpotOffset"] = -f32(math.cos(outer)) * sc
    return l


def torus_knot_scene(nu: int = 300, nv: int = 120, width: int = 1920, height: int = 1080, tex_size: int = 1024,
                     seed: int = 11, extra_lights: bool = False, alpha_material: bool = False) -> SceneData:
    """BASELINE config C1 stand-in (DamagedHelmet is not among the assets): a (2,3) torus knot of
    2*nu*nv triangles (72,000 by default) with UVs, normals and tangents, split over two materials that
    each bind a two-layer 1024^2 texture, lit by the default directional light."""
    from . import textures as tx

    u = np.linspace(0, 2 * math.pi, nu, endpoint=False)
    v = np.linspace(0, 2 * math.pi, nv, endpoint=False)
    p, q, R, r0, tube = 2, 3, 2.0, 0.8, 0.42
    cu = np.stack([(R + r0 * np.cos(q * u)) * np.cos(p * u), r0 * np.sin(q * u), (R + r0 * np.cos(q * u)) * np.sin(p * u)], 1)
    du = np.roll(cu, -1, 0) - np.roll(cu, 1, 0)
    T = du / np.linalg.norm(du, axis=1, keepdims=True)
    up = np.array([0.0, 1.0, 0.0])
    Nn = np.cross(T, up); Nn /= np.linalg.norm(Nn, axis=1, keepdims=True)
    B = np.cross(T, Nn)
    cv, sv = np.cos(v), np.sin(v)
    normal = Nn[:, None, :] * cv[None, :, None] + B[:, None, :] * sv[None, :, None]           # [nu, nv, 3]
    pos = cu[:, None, :] + tube * normal
    tang = np.broadcast_to(T[:, None, :], normal.shape)
    uvs = np.stack(np.broadcast_arrays((np.arange(nu) / nu * 2.0)[:, None], (np.arange(nv) / nv)[None, :]), -1)
    # closed surface: the UV seam needs duplicated vertices; add one extra ring/column
    def wrap(a):
        a = np.concatenate([a, a[:1]], 0)
        return np.concatenate([a, a[:, :1]], 1)
    pos, normal, tang = wrap(pos), wrap(normal), wrap(tang)
    uvs = wrap(uvs)
    uvs[-1, :, 0] = 2.0
    uvs[:, -1, 1] = 1.0
    rows, cols = nu + 1, nv + 1
    ii, jj = np.meshgrid(np.arange(nu), np.arange(nv), indexing="ij")
    v00, v01, v10, v11 = ii * cols + jj, ii * cols + jj + 1, (ii + 1) * cols + jj, (ii + 1) * cols + jj + 1
    tris = np.stack([np.stack([v00, v10, v01], -1), np.stack([v01, v10, v11], -1)], 2).reshape(-1, 3)
    P, Nf, Tf, UV = pos.reshape(-1, 3), normal.reshape(-1, 3), tang.reshape(-1, 3), uvs.reshape(-1, 2)
    Tf4 = np.concatenate([Tf, np.where((np.arange(len(Tf)) % 7) == 0, -1.0, 1.0)[:, None]], 1)

    half = len(tris) // 2
    materials = np.zeros(2, dtype=MATERIAL_DTYPE)
    materials["TextureId"] = [0, 1]
    materials["AlphaCutoff"] = [255, 128 if alpha_material else 255]
    materials["IsDoubleSided"] = [0, 1 if alpha_material else 0]
    # triangle winding: front faces must have det > 0 in this renderer's screen space
    m0 = meshletize(P, tris[:half][:, [0, 2, 1]], uv=UV, normals=Nf, tangents=Tf4, material_id=0, alpha_cutoff=255)
    m1 = meshletize(P, tris[half:][:, [0, 2, 1]], uv=UV, normals=Nf, tangents=Tf4, material_id=1,
                    alpha_cutoff=128 if alpha_material else 255)
    meshlets = concat_meshlets([m0, m1])
    texs =[tx.procedural_material_texture(tex_size, seed), tx.procedural_material_texture(tex_size, seed + 1, alpha_holes=alpha_material)]
    lights = default_light()
    if extra_lights:
        lights = np.concatenate([lights, make_light(1, position=(0.5, 2.0, 3.0), color=(1.0, 0.6, 0.3), intensity=4000.0, radius=12.0),
                                 make_light(2, position=(-3.0, 3.0, 2.0), direction=(0.6, -0.7, -0.4), color=(0.3, 0.6, 1.0),
                                            intensity=9000.0, radius=20.0, inner=0.25, outer=0.55)])
    model = cam.mat_mul(cam.translate((0.0, 0.2, 0.0)), cam.mat_mul(cam.rotate_axis((0.3, 1.0, 0.1), 0.6), cam.scale(0.9)))
    camera = cam.Camera(position=(0.2, 1.0, 3.2), euler=(0.05, -0.28), fov_deg=90.0, aspect=width / height)
    return SceneData(f"torusknot{nu}x{nv}", meshlets, [DrawNode(0, len(meshlets), model)], camera, width, height,
                     materials=materials, textures=texs, lights=lights)


def room_scene(width: int = 1920, height: int = 1080, subdiv: int = 5, size: float = 10.0, seed: int = 21) -> SceneData:
    """Sponza-like stress for the big-triangle paths: the camera stands inside a box whose faces are coarse
    grids (triangles hundreds of pixels wide, several crossing the near plane or leaving the guard band, which
    the binned path counts as clipped and drops, Rasterizer.cpp:567-569) plus a few pillars of small triangles."""
    faces = []
    s = size
    lin = np.linspace(-s, s, subdiv + 1)
    a, b = np.meshgrid(lin, lin, indexing="ij")
    one = np.ones_like(a)
    for axis, sign in ((0, 1), (0, -1), (1, 1), (1, -1), (2, 1), (2, -1)):
        coords = [a, b]
        coords.insert(axis, one * sign * s)
        P = np.stack(coords, -1).reshape(-1, 3)
        n = subdiv + 1
        ii, jj = np.meshgrid(np.arange(subdiv), np.arange(subdiv), indexing="ij")
        v00, v01, v10, v11 = ii * n + jj, ii * n + jj + 1, (ii + 1) * n + jj, (ii + 1) * n + jj + 1
        t = np.stack([np.stack([v00, v10, v01], -1), np.stack([v01, v10, v11], -1)], 2).reshape(-1, 3)
        # inward-facing: pick the winding whose normal points to the origin
        nrm = np.cross(P[t[0, 1]] - P[t[0, 0]], P[t[0, 2]] - P[t[0, 0]])
        if np.dot(nrm, -P[t[0, 0]]) < 0:   # (this renderer's front faces have det > 0 with y pointing down-screen)
            t = t[:, [0, 2, 1]]
        faces.append(meshletize(P, t))
    # pillars: thin boxes of many small triangles
    r = rand01(seed, 64).reshape(-1, 4)
    for k in range(8):
        cx, cz = (r[k, 0] - 0.5) * 1.4 * s, (r[k, 1] - 0.5) * 1.4 * s
        hgt = np.linspace(-s, s, 40)
        ang = np.linspace(0, 2 * math.pi, 13)
        A, H = np.meshgrid(ang, hgt, indexing="ij")
        P = np.stack([cx + 0.5 * np.cos(A), H, cz + 0.5 * np.sin(A)], -1).reshape(-1, 3)
        n = len(hgt)
        ii, jj = np.meshgrid(np.arange(len(ang) - 1), np.arange(n - 1), indexing="ij")
        v00, v01, v10, v11 = ii * n + jj, ii * n + jj + 1, (ii + 1) * n + jj, (ii + 1) * n + jj + 1
        t = np.stack([np.stack([v00, v01, v10], -1), np.stack([v01, v11, v10], -1)], 2).reshape(-1, 3)
        faces.append(meshletize(P, t))
    meshlets = concat_meshlets(faces)
    camera = cam.Camera(position=(1.0, -7.5, 2.0), euler=(0.7, 0.1), fov_deg=90.0, aspect=width / height)
    return SceneData(f"room{subdiv}", meshlets, [DrawNode(0, len(meshlets), cam.identity())], camera, width, height)


def closeup_alpha_scene(width: int = 960, height: int = 540, nu: int = 40, nv: int = 10, tex_size: int = 128) -> SceneData:
    """Clipper stress: a coarse torus knot (big triangles) with an alpha-tested, double-sided second material,
    seen from a camera that sits just above the tube surface — dozens of triangles of both materials cross the
    camera plane (w < near) or leave the guard band, so the unbinned path has to clip them and, for the
    alpha-tested ones, remap the barycentrics of every piece (Rasterizer.h:312-318)."""
    scene = torus_knot_scene(nu, nv, width, height, tex_size=tex_size, alpha_material=True)
    scene.name = f"closeup_alpha{nu}x{nv}"
    scene.camera = cam.Camera(position=(1.07735, 0.19206, -0.15236), euler=(4.71, 0.0), fov_deg=100.0, aspect=width / height, near_z=0.05)
    return scene


def patchwork_scene(patches_x: int = 40, patches_y: int = 32, width: int = 1280, height: int = 720, num_materials: int = 9,
                    seed: int = 31) -> SceneData:
    """The heightfield grid with a different material on (almost) every neighbouring meshlet: `num_materials` materials with
    textures from 32^2 to 512^2 texels, every other one without a normal / metallic-roughness layer, one without any
    texture-bound material at all (MaterialId = UINT_MAX). Most 4x4 fragments along meshlet borders then hold several
    materials with different mip decisions — the per-fragment material waterfall and filter votes of ResolveSurface
    (Shading.cpp:532-545, Texture.h:432) at their worst."""
    from . import textures as tx
    scene = grid_scene(patches_x, patches_y, width, height, seed=seed, material_id=0)
    pick = (splitmix64(seed, len(scene.meshlets)) % np.uint64(num_materials + 1)).astype(np.int64)
    scene.meshlets["MaterialId"] = np.where(pick == num_materials, NO_MATERIAL, pick).astype(np.uint32)
    scene.materials = np.zeros(num_materials, dtype=MATERIAL_DTYPE)
    scene.materials["AlphaCutoff"] = 255
    scene.materials["TextureId"] = np.arange(num_materials)
    scene.textures = [tx.procedural_material_texture(32 << (k % 5), seed=seed + k, with_nmr=(k % 2 == 0)) for k in range(num_materials)]
    scene.lights = np.concatenate([default_light(), make_light(1, position=(2.5, 4.0, 1.0), color=(1.0, 0.6, 0.3), intensity=60.0, radius=9.0)])
    scene.name = f"patchwork{patches_x}x{patches_y}"
    return scene


def resolve_uniforms(scene: SceneData, node: DrawNode, exposure: float = 1.0) -> dict:
    """The ShadingContext fields Resolve reads (Shading.h:21-33, Shading.cpp:659)."""
    proj, view = scene.view_proj()
    w2c = cam.mat_mul(proj, view)
    return dict(world_to_clip=w2c, object_to_clip=cam.mat_mul(w2c, node.model),
                object_to_world3=np.ascontiguousarray(node.model[0:3, 0:3]),
                inv_screen_proj=cam.inverse_screen_proj(w2c, scene.width, scene.height),
                view_pos=scene.camera.position.astype(f32), exposure=exposure)
