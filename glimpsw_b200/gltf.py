"""Host-side scene import: glTF 2.0 -> the path's input structures.

Mirrors `Scene::ImportGltf` (src/SwRast/Scene.cpp:156-368) and its helpers `LoadTextures` / `CombineNormalMR` /
`InsertEmissiveMask` (:17-102): one material + one layered RGBA8 texture per glTF material, every mesh primitive split
into `Meshlet`s (<= 64 vertices, <= 128 triangles, SoA positions, fp16 UVs, octahedron-packed normal + tangent,
handedness bits), nodes walked depth-first with `GlobalTransform = parent * local`, KHR_lights_punctual lights.

What differs from the reference, by necessity (none of its third-party importers exist in this image):
  * cgltf -> the JSON + buffer parsing below (`.gltf` with external / data-URI buffers, and `.glb`);
  * meshoptimizer's `meshopt_buildMeshlets` -> `build_meshlets_order` (the library's published selection rule: fewest new
    vertices, dangling triangles first, then the distance / normal-cone score with cone_weight 0.25, k-d tree restarts) (the
    default) or the simpler `adjacency_order` (cheapest triangle first, no spatial score), each followed by the
    cutting scan of `scenes.meshletize` (different meshlet boundaries than an upstream import = different surface ids, same
    geometry); `meshopt_computeMeshletBounds` (bounding sphere + normal cone) is restated in `compute_meshlet_bounds`;
  * stb_image -> Pillow for the texture files.
Everything downstream (the meshlet bytes, materials, texture layout, uniforms) is the reference's format, so an imported
scene goes through `api.Rasterizer.upload_scene` / the oracle like the procedural ones.
"""
from __future__ import annotations

import base64
import json
import math
import os
import struct

import numpy as np

from . import camera as cam
from . import textures as tx
from .layout import LIGHT_DTYPE, MATERIAL_DTYPE
from .scenes import DrawNode, SceneData, concat_meshlets, meshletize

f32 = np.float32
NO_MATERIAL = 0xFFFFFFFF

_COMPONENT = {5120: np.int8, 5121: np.uint8, 5122: np.int16, 5123: np.uint16, 5125: np.uint32, 5126: np.float32}
_WIDTH = {"SCALAR": 1, "VEC2": 2, "VEC3": 3, "VEC4": 4, "MAT4": 16}


class GltfFile:
    """The parsed JSON plus its binary buffers (cgltf_parse_file + cgltf_load_buffers, Scene.cpp:157-168)."""

    def __init__(self, path: str):
        self.base = os.path.dirname(os.path.abspath(path))
        raw = open(path, "rb").read()
        glb_bin = None
        if raw[:4] == b"glTF":                                        # binary container: JSON chunk + optional BIN chunk
            _, _, total = struct.unpack_from("<III", raw, 0)
            off, chunks = 12, []
            while off < total:
                n, kind = struct.unpack_from("<II", raw, off)
                chunks.append((kind, raw[off + 8: off + 8 + n]))
                off += 8 + n
            self.json = json.loads(next(c for k, c in chunks if k == 0x4E4F534A))
            glb_bin = next((c for k, c in chunks if k == 0x004E4942), None)
        else:
            self.json = json.loads(raw)
        self.buffers = []
        for b in self.json.get("buffers", []):
            uri = b.get("uri")
            if uri is None:
                data = glb_bin
            elif uri.startswith("data:"):
                data = base64.b64decode(uri.split(",", 1)[1])
            else:
                data = open(os.path.join(self.base, uri), "rb").read()
            if data is None or len(data) < b["byteLength"]:
                raise RuntimeError("Failed to load associated GLTF data")     # Scene.cpp:166
            self.buffers.append(np.frombuffer(data, dtype=np.uint8))

    def accessor(self, index: int) -> np.ndarray:
        """cgltf_accessor_unpack_*: [count, width] array in the accessor's component type (normalized ints -> float)."""
        a = self.json["accessors"][index]
        dt, width, count = _COMPONENT[a["componentType"]], _WIDTH[a["type"]], a["count"]
        if "bufferView" not in a:
            out = np.zeros((count, width), dtype=dt)
        else:
            v = self.json["bufferViews"][a["bufferView"]]
            start = v.get("byteOffset", 0) + a.get("byteOffset", 0)
            item = np.dtype(dt).itemsize * width
            stride = v.get("byteStride", 0) or item
            buf = self.buffers[v["buffer"]]
            idx = start + np.arange(count)[:, None] * stride + np.arange(item)[None, :]
            out = buf[idx].copy().view(dt).reshape(count, width)
        if a.get("normalized") and dt != np.float32:
            out = np.maximum(out.astype(f32) / float(np.iinfo(dt).max), -1.0)
        return out

    def image_rgba(self, texture_info) -> np.ndarray | None:
        """LoadImage (Scene.cpp:51-76): the RGBA8 pixels of a texture view, or None when the view is empty."""
        if not texture_info:
            return None
        tex = self.json["textures"][texture_info["index"]]
        if "source" not in tex:
            return None
        img = self.json["images"][tex["source"]]
        from PIL import Image
        import io
        if "uri" in img and not img["uri"].startswith("data:"):
            src = os.path.join(self.base, img["uri"])
        elif "uri" in img:
            src = io.BytesIO(base64.b64decode(img["uri"].split(",", 1)[1]))
        else:
            v = self.json["bufferViews"][img["bufferView"]]
            o = v.get("byteOffset", 0)
            src = io.BytesIO(self.buffers[v["buffer"]][o: o + v["byteLength"]].tobytes())
        try:
            return np.asarray(Image.open(src).convert("RGBA"), dtype=np.uint8)
        except Exception as exc:
            raise RuntimeError("Failed to load image") from exc                # Scene.cpp:68


def _words(rgba: np.ndarray) -> np.ndarray:
    return rgba.astype(np.uint32) @ np.array([1, 1 << 8, 1 << 16, 1 << 24], dtype=np.uint32)


def combine_normal_mr(normal: np.ndarray, mr: np.ndarray | None) -> np.ndarray:
    """CombineNormalMR (Scene.cpp:17-36): renormalized normal.xy in RG, metallic in B, roughness in A."""
    out = normal.copy()
    n = normal[..., :3].astype(f32) / f32(127.0) - f32(1.0)
    n = n / np.sqrt((n * n).sum(axis=-1, keepdims=True), dtype=f32) * f32(127.0) + f32(127.0)
    r = np.floor(np.abs(n) + f32(0.5)) * np.sign(n)                            # roundf: half away from zero
    out[..., 0] = r[..., 0].astype(np.uint8)
    out[..., 1] = r[..., 1].astype(np.uint8)
    if mr is not None:
        out[..., 2] = mr[..., 2]
        out[..., 3] = mr[..., 1]
    return out


def insert_emissive_mask(base: np.ndarray, emissive: np.ndarray) -> np.ndarray:
    """InsertEmissiveMask (Scene.cpp:38-49): alpha 255 marks emissive texels, everything else is capped at 254."""
    out = base.copy()
    lit = (emissive[..., :3] > 8).any(axis=-1)
    out[..., 3] = np.where(lit, 255, np.minimum(base[..., 3], 254)).astype(np.uint8)
    return out


def load_textures(g: GltfFile, mat: dict) -> tx.TextureData:
    """LoadTextures (Scene.cpp:78-102)."""
    pbr = mat.get("pbrMetallicRoughness", {})
    base = g.image_rgba(pbr.get("baseColorTexture"))
    if base is None:
        return tx.create_texture(4, 4, 1, 1)
    h, w = base.shape[:2]
    normal, mr, emissive = g.image_rgba(mat.get("normalTexture")), g.image_rgba(pbr.get("metallicRoughnessTexture")), g.image_rgba(mat.get("emissiveTexture"))
    has_normals = normal is not None and normal.shape[:2] == (h, w)
    has_emissive = emissive is not None and emissive.shape[:2] == (h, w)
    tex = tx.create_texture(w, h, 8, 3 if has_emissive else (2 if has_normals else 1))
    if has_normals:
        tx.set_pixels(tex, _words(combine_normal_mr(normal, mr if mr is not None and mr.shape[:2] == (h, w) else None)), layer=1)
    if has_emissive:
        base = insert_emissive_mask(base, emissive)
        tx.set_pixels(tex, _words(emissive), layer=2)
    tx.set_pixels(tex, _words(base), layer=0)
    tx.generate_mips(tex)
    return tex


def _node_local(n: dict) -> np.ndarray:
    """cgltf_node_transform_local: the node's matrix, or T * R * S — here as a mathematical (row, column) float64 matrix;
    import_gltf transposes to the package's column-major [c, r] float32 convention when it stores a node."""
    if "matrix" in n:
        return np.asarray(n["matrix"], dtype=np.float64).reshape(4, 4).T        # glTF stores column-major
    t = np.eye(4)
    t[:3, 3] = n.get("translation", (0, 0, 0))
    x, y, z, w = n.get("rotation", (0, 0, 0, 1))
    r = np.eye(4)
    r[:3, :3] = [[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                 [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                 [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]]
    s = np.diag(list(n.get("scale", (1, 1, 1))) + [1.0])
    return t @ r @ s


def adjacency_order(tris: np.ndarray, max_verts: int = 64, max_tris: int = 128) -> np.ndarray:
    """Triangle order for the meshletizer, grown by adjacency like meshopt_buildMeshlets (Scene.cpp:199-237) grows its
    meshlets: the next triangle is one that shares vertices with the meshlet under construction, preferring those that add no
    new vertex, then one, then two; when the meshlet is full the next one is seeded from a neighbour of the old one.
    `scenes.meshletize` (a linear scan that cuts whenever a limit is hit) then finds the same cuts, so an index buffer in
    arbitrary order still yields well-filled meshlets. Returns a permutation of range(len(tris))."""
    tris = np.asarray(tris, dtype=np.int64)
    n = len(tris)
    if n == 0:
        return np.zeros(0, dtype=np.int64)
    T = tris.tolist()
    adj: dict = {}
    for t, (a, b, c) in enumerate(T):
        for v in (a, b, c):
            adj.setdefault(v, []).append(t)
    done = [False] * n
    order = []
    in_meshlet: set = set()
    missing: dict = {}                    # candidate triangle -> number of its vertices not yet in the meshlet
    buckets = (set(), set(), set(), set())
    num_tris = 0
    next_unseen = 0

    def add_vertex(v):
        in_meshlet.add(v)
        for t in adj[v]:
            if done[t]:
                continue
            k = missing.get(t)
            if k is None:
                k = len({x for x in T[t] if x not in in_meshlet})
            else:
                buckets[k].discard(t)
                k -= 1
            missing[t] = k
            buckets[k].add(t)

    def reset(seed_candidates):
        nonlocal num_tris
        in_meshlet.clear()
        missing.clear()
        for bkt in buckets:
            bkt.clear()
        num_tris = 0
        return next((t for t in seed_candidates if not done[t]), None)

    seed = None
    while len(order) < n:
        t = None
        room = max_verts - len(in_meshlet)
        if num_tris < max_tris:
            for k in range(0, min(room, 3) + 1):
                if buckets[k]:
                    t = buckets[k].pop()
                    break
        if t is None:                                   # meshlet full (or nothing adjacent left): start the next one nearby
            leftovers = [x for bkt in buckets for x in bkt]
            seed = reset(leftovers)
            if seed is None:
                while done[next_unseen]:
                    next_unseen += 1
                seed = next_unseen
            t = seed
        missing.pop(t, None)
        done[t] = True
        order.append(t)
        num_tris += 1
        for v in dict.fromkeys(T[t]):
            if v not in in_meshlet:
                add_vertex(v)
    return np.asarray(order, dtype=np.int64)


def build_meshlets_order(positions: np.ndarray, tris: np.ndarray, max_verts: int = 64, max_tris: int = 128,
                         cone_weight: float = 0.25) -> np.ndarray:
    """Triangle order of meshoptimizer's `meshopt_buildMeshlets(…, 64, 128, cone_weight = 0.25)` (Scene.cpp:215-224), restated from
    the library's published algorithm (clusterizer.cpp, v1.1 as pinned by the reference's CMakeLists.txt:21; the library itself is not
    in this image, so ties and floating-point details may differ from an upstream import):
      * every candidate — a live triangle that shares a vertex with the meshlet under construction — is ranked by `extra`: 0 if it
        adds no vertex; 1 if it adds some but one of its vertices has no other live triangle left (a dangling triangle is taken now
        or strands a vertex); else 1 + the number of vertices it adds;
      * within the best rank the lowest score wins: (1 + distance / expected_radius * (1 - cone_weight)) * max(1e-3, 1 - spread *
        cone_weight), distance from the meshlet's mean triangle centroid, spread = triangle normal . meshlet's mean normal,
        expected_radius = sqrt(mean triangle area / 2 * max_tris) / 2;
      * if that triangle does not fit (vertex or triangle limit) the choice is redone with the topological score (live triangles
        around its three vertices) — it seeds the next meshlet;
      * with no candidate left the nearest live triangle to the meshlet's centre (k-d tree over centroids) continues.
    `scenes.meshletize` cuts the returned stream exactly where the library's appendMeshlet would. Returns a permutation."""
    tris = np.asarray(tris, dtype=np.int64)
    n = len(tris)
    if n == 0:
        return np.zeros(0, dtype=np.int64)
    from scipy.spatial import cKDTree
    pos = np.asarray(positions, dtype=np.float64)
    p0, p1, p2 = pos[tris[:, 0]], pos[tris[:, 1]], pos[tris[:, 2]]
    nrm = np.cross(p1 - p0, p2 - p0)
    area = np.linalg.norm(nrm, axis=1)
    normal = np.where(area[:, None] > 0, nrm / np.maximum(area, 1e-300)[:, None], 0.0)
    centroid = (p0 + p1 + p2) / 3.0
    expected_radius = math.sqrt(float(area.sum()) / n * 0.5 * max_tris) * 0.5
    tree = cKDTree(centroid)
    T = tris.tolist()
    adj: list = [[] for _ in range(int(tris.max()) + 1)]
    for t, (a, b, c) in enumerate(T):
        for v in dict.fromkeys((a, b, c)):
            adj[v].append(t)
    live = np.array([len(x) for x in adj], dtype=np.int64)         # live triangles around every vertex
    emitted = np.zeros(n, dtype=bool)
    order: list = []
    used: set = set()                                              # vertices of the meshlet under construction
    missing: dict = {}                                             # candidate triangle -> vertices it would add
    num_tris = 0
    acc_c, acc_n = np.zeros(3), np.zeros(3)

    def reset():
        nonlocal num_tris, acc_c, acc_n
        used.clear()
        missing.clear()
        num_tris = 0
        acc_c, acc_n = np.zeros(3), np.zeros(3)

    def select(use_cone: bool):
        cand = np.fromiter(missing.keys(), dtype=np.int64, count=len(missing))
        miss = np.fromiter(missing.values(), dtype=np.int64, count=len(missing))
        dangling = (live[tris[cand]] == 1).any(axis=1)
        extra = np.where(miss == 0, 0, np.where(dangling, 1, miss + 1))
        best = extra.min()
        pick = np.nonzero(extra == best)[0]
        sel = cand[pick]
        if use_cone and num_tris:
            center = acc_c / num_tris
            ln = np.linalg.norm(acc_n)
            axis = acc_n / ln if ln > 0 else acc_n
            dist = np.linalg.norm(centroid[sel] - center, axis=1)
            spread = normal[sel] @ axis
            score = (1.0 + dist / expected_radius * (1.0 - cone_weight)) * np.maximum(1e-3, 1.0 - spread * cone_weight)
        else:
            score = live[tris[sel]].sum(axis=1) - 3
        k = int(np.argmin(score))
        return int(sel[k]), int(miss[pick[k]])

    while len(order) < n:
        t = None
        if missing:
            t, add = select(True)
            if len(used) + add > max_verts or num_tris >= max_tris:
                t, add = select(False)
        if t is None:                                              # nothing adjacent: continue with the nearest live triangle
            center = acc_c / num_tris if num_tris else np.zeros(3)
            k = 8
            while t is None:
                _, idx = tree.query(center, k=min(k, n))
                for j in np.atleast_1d(idx).tolist():
                    if not emitted[j]:
                        t = j
                        break
                k *= 8
            add = len({v for v in T[t] if v not in used})
        if len(used) + add > max_verts or num_tris >= max_tris:    # appendMeshlet: the meshlet is full, t opens the next one
            reset()
        emitted[t] = True
        order.append(t)
        missing.pop(t, None)
        num_tris += 1
        acc_c = acc_c + centroid[t]
        acc_n = acc_n + normal[t]
        for v in dict.fromkeys(T[t]):
            live[v] -= 1
            adj[v].remove(t)
        for v in dict.fromkeys(T[t]):
            if v in used:
                continue
            used.add(v)
            for u in adj[v]:                                       # its live triangles become candidates / need one vertex less
                k = missing.get(u)
                missing[u] = (len({x for x in T[u] if x not in used}) if k is None else k - 1)
    return np.asarray(order, dtype=np.int64)


def _bounding_sphere(points: np.ndarray):
    """meshoptimizer's computeBoundingSphere (clusterizer.cpp, v1.1 as pinned by the reference's CMakeLists.txt:21; restated from
    the published algorithm, float32 throughout): the extreme points along the three axes give three candidate diameters, the
    longest becomes the initial sphere, then every point outside pulls the sphere towards itself just far enough to be covered."""
    pts = np.asarray(points, dtype=f32)
    pmin, pmax = pts.argmin(axis=0), pts.argmax(axis=0)                      # first index on ties, like the strict compares
    best_d2, best_axis = f32(0), 0
    for axis in range(3):
        d = pts[pmax[axis]] - pts[pmin[axis]]
        d2 = f32(d[0] * d[0] + d[1] * d[1] + d[2] * d[2])
        if d2 > best_d2:
            best_d2, best_axis = d2, axis
    p1, p2 = pts[pmin[best_axis]], pts[pmax[best_axis]]
    center = ((p1 + p2) / f32(2)).astype(f32)
    radius = f32(np.sqrt(best_d2) / f32(2))
    for p in pts:
        d = p - center
        d2 = f32(d[0] * d[0] + d[1] * d[1] + d[2] * d[2])
        if d2 > radius * radius:
            dist = f32(np.sqrt(d2))
            k = f32(0.5) + (radius / dist) / f32(2)
            center = (center * k + p * (f32(1) - k)).astype(f32)
            radius = f32((radius + dist) / f32(2))
    return center, radius


def compute_meshlet_bounds(meshlets: np.ndarray) -> None:
    """meshopt_computeMeshletBounds (Scene.cpp:236-245 stores its result): bounding sphere of the corners of the non-degenerate
    triangles, and the normal cone — axis = centre of the bounding sphere of the unit triangle normals, cutoff =
    sqrt(1 - mindp^2) with mindp the smallest normal . axis, apex = centre - axis * max over triangles of
    dot(centre - corner, n) / dot(axis, n); a cone wider than ~168 degrees (mindp <= 0.1) is stored as cutoff 1 without apex /
    axis. Writes BoundCenter, BoundRadius, ConeApex, ConeAxis, ConeCutoff in place."""
    with np.errstate(all="ignore"):
        for m in meshlets:
            nt = int(m["NumTriangles"])
            pos = m["Positions"].T.astype(f32)                                   # [64, 3]
            idx = m["Indices"][:, :nt].T.astype(np.int64)                        # [nt, 3]
            corners = pos[idx]                                                   # [nt, 3, 3]
            normal = np.cross(corners[:, 1] - corners[:, 0], corners[:, 2] - corners[:, 0]).astype(f32)
            area = np.sqrt((normal * normal).sum(axis=1, dtype=f32)).astype(f32)
            keep = area > 0
            m["BoundCenter"], m["BoundRadius"] = 0, 0
            m["ConeApex"], m["ConeAxis"], m["ConeCutoff"] = 0, 0, 0
            if not keep.any():
                continue
            corners, normal = corners[keep], (normal[keep] / area[keep, None]).astype(f32)
            center, radius = _bounding_sphere(corners.reshape(-1, 3))
            ncenter, _ = _bounding_sphere(normal)
            alen = f32(np.sqrt((ncenter * ncenter).sum(dtype=f32)))
            axis = (ncenter * (f32(0) if alen == 0 else f32(1) / alen)).astype(f32)
            mindp = min(f32(1), f32((normal @ axis).min()))
            m["BoundCenter"], m["BoundRadius"] = center, radius
            if mindp <= f32(0.1):
                m["ConeCutoff"] = 1.0
                continue
            t = ((center[None, :] - corners[:, 0]) * normal).sum(axis=1, dtype=f32) / (normal @ axis)
            maxt = max(f32(0), f32(t.max()))
            m["ConeApex"] = center - axis * maxt
            m["ConeAxis"] = axis
            m["ConeCutoff"] = f32(np.sqrt(f32(1) - mindp * mindp))


def import_gltf(path: str, width: int = 1920, height: int = 1080, camera: cam.Camera | None = None,
                flip_winding: bool = False, builder: str = "meshopt") -> SceneData:
    """Scene::ImportGltf. Returns a SceneData (meshlets, one DrawNode per glTF node with a mesh, materials, textures,
    lights). `flip_winding` swaps two indices of every triangle (for assets authored with the other front face).
    `builder`: "meshopt" = `build_meshlets_order`, the library's scored selection (cone_weight 0.25 like Scene.cpp:215; what the
    committed Sponza fixture is built with); "adjacency" = `adjacency_order`, the quick topological stand-in (on Sponza_LowPoly its
    meshlets' bounding spheres are six times larger on average)."""
    if builder not in ("adjacency", "meshopt"):
        raise ValueError(f"unknown meshlet builder {builder!r}")
    g = GltfFile(path)
    js = g.json
    materials = np.zeros(len(js.get("materials", [])), dtype=MATERIAL_DTYPE)
    textures = []
    for i, mat in enumerate(js.get("materials", [])):                          # Scene.cpp:174-184
        textures.append(load_textures(g, mat))
        materials[i]["TextureId"] = i
        materials[i]["IsDoubleSided"] = 1 if mat.get("doubleSided") else 0
        cutoff = mat.get("alphaCutoff", 0.5)
        materials[i]["AlphaCutoff"] = int(cutoff * 255.0 + 0.5) & 255 if mat.get("alphaMode") == "MASK" else 255

    parts, ranges, cursor = [], [], 0
    for mesh in js.get("meshes", []):                                          # Scene.cpp:193-283
        start = cursor
        for prim in mesh["primitives"]:
            if prim.get("mode", 4) != 4:
                continue
            attrs = prim["attributes"]
            pos = g.accessor(attrs["POSITION"]).astype(f32)
            idx = g.accessor(prim["indices"]).reshape(-1).astype(np.int64) if "indices" in prim else np.arange(len(pos), dtype=np.int64)
            tris = idx[: len(idx) // 3 * 3].reshape(-1, 3)
            if flip_winding:
                tris = tris[:, [0, 2, 1]]
            uv = g.accessor(attrs["TEXCOORD_0"]).astype(f32) if "TEXCOORD_0" in attrs else None
            nrm = g.accessor(attrs["NORMAL"]).astype(f32) if "NORMAL" in attrs else None
            tan = g.accessor(attrs["TANGENT"]).astype(f32) if "TANGENT" in attrs and nrm is not None else None
            if nrm is not None and tan is None:
                tan = np.zeros((len(pos), 4), dtype=f32)                        # glm::vec4 tangent = 0 (Scene.cpp:258)
            mat_id = prim.get("material", None)
            tris = tris[adjacency_order(tris) if builder == "adjacency" else build_meshlets_order(pos, tris)]   # meshopt_buildMeshlets (Scene.cpp:218-224)
            m = meshletize(pos, tris, uv=uv, normals=nrm, tangents=tan,
                           material_id=NO_MATERIAL if mat_id is None else mat_id,
                           alpha_cutoff=255 if mat_id is None else int(materials[mat_id]["AlphaCutoff"]))
            parts.append(m)
            cursor += len(m)
        ranges.append((start, cursor))
    meshlets = concat_meshlets(parts) if parts else meshletize(np.zeros((0, 3), dtype=f32), np.zeros((0, 3), dtype=np.int64))
    compute_meshlet_bounds(meshlets)                                           # Scene.cpp:236-245

    nodes, lights = [], []
    punctual = js.get("extensions", {}).get("KHR_lights_punctual", {}).get("lights", [])

    def recurse(index: int, parent: np.ndarray):                               # Scene.cpp:286-352 (pre-order DFS)
        n = js["nodes"][index]
        glob = parent @ _node_local(n)
        if "mesh" in n:
            a, b = ranges[n["mesh"]]
            if b > a:
                nodes.append(DrawNode(a, b - a, np.ascontiguousarray(glob.T.astype(f32))))
        li = n.get("extensions", {}).get("KHR_lights_punctual", {}).get("light")
        if li is not None:
            src = punctual[li]
            l = np.zeros(1, dtype=LIGHT_DTYPE)
            l["Type"] = {"directional": 0, "point": 1, "spot": 2}[src["type"]]
            l["Position"] = glob[:3, 3]
            d = -glob[:3, 2]
            l["Direction"] = d / np.linalg.norm(d)
            l["Color"] = src.get("color", (1, 1, 1))
            l["Intensity"] = src.get("intensity", 1.0)
            rng = float(src.get("range", 0.0))
            rng = rng if rng > 0 else 1e6                                       # Light::SetRadius (Scene.h:65-68)
            l["Radius"] = rng
            l["InvRadiusSq"] = f32(1.0) / (f32(rng) * f32(rng))
            spot = src.get("spot", {})
            inner, outer = float(spot.get("innerConeAngle", 0.0)), float(spot.get("outerConeAngle", math.pi / 4))
            l["SpotInnerAngle"], l["SpotOuterAngle"] = inner, outer             # Light::SetSpotAngles (Scene.h:69-74)
            scale = f32(1.0) / max(f32(math.cos(inner)) - f32(math.cos(outer)), f32(1e-4))
            l["SpotScale"], l["SpotOffset"] = scale, -f32(math.cos(outer)) * scale
            lights.append(l)
        for c in n.get("children", []):
            recurse(c, glob)

    scene_def = js["scenes"][js.get("scene", 0)] if js.get("scenes") else {"nodes": list(range(len(js.get("nodes", []))))}
    for root in scene_def.get("nodes", []):
        recurse(root, np.eye(4))

    if camera is None:                                                         # the pose RasterBench.cpp:68 hard-codes
        camera = cam.Camera(position=(-0.46608772755098471, 8.4925445559659511, -1.7022251220187172),
                            euler=(-3.13643026, -1.4724431), fov_deg=90.0, aspect=width / height)
    scene = SceneData(os.path.splitext(os.path.basename(path))[0], meshlets, nodes, camera, width, height)
    scene.materials = materials
    scene.textures = textures
    scene.lights = np.concatenate(lights) if lights else np.zeros(0, dtype=LIGHT_DTYPE)
    return scene
