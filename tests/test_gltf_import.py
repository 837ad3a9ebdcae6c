"""Scene import: glimpsw_b200.gltf restates Scene::ImportGltf (Scene.cpp:156-368) on the host.

A small glTF is written on the fly (two nodes under a transformed parent, u16/u32 indices, interleaved attributes,
a material with base colour / normal / metallic-roughness PNGs, a punctual light), imported, and checked against the
source arrays; the imported scene then goes through the same oracle-vs-CUDA parity gate as the procedural scenes. When
the reference checkout is present (this container, not the GPU box) its Sponza_LowPoly asset is imported as well."""
import base64
import json
import os
import struct

import numpy as np
import pytest

from glimpsw_b200 import gltf, scenes, textures as tx
from helpers import oracle_render, gpu_render, assert_visbuffer_equal

SPONZA = "/root/reference/assets/models/Sponza/Sponza_LowPoly.gltf"


def _write_test_gltf(tmp_path, glb=False):
    from PIL import Image
    nu, nv = 24, 16                                                    # a bumpy sheet: (nu x nv) vertices
    u, v = np.meshgrid(np.linspace(0, 1, nu), np.linspace(0, 1, nv), indexing="xy")
    pos = np.stack([u * 2 - 1, v * 2 - 1, 0.15 * np.sin(u * 9) * np.cos(v * 7)], -1).reshape(-1, 3).astype(np.float32)
    nrm = np.tile(np.array([0, 0, 1], dtype=np.float32), (len(pos), 1))
    tan = np.tile(np.array([1, 0, 0, -1], dtype=np.float32), (len(pos), 1))
    uv = np.stack([u, v], -1).reshape(-1, 2).astype(np.float32)
    tris = []
    for r in range(nv - 1):
        for c in range(nu - 1):
            a, b, d, e = r * nu + c, r * nu + c + 1, (r + 1) * nu + c, (r + 1) * nu + c + 1
            tris += [(a, d, b), (b, d, e)]
    tris = np.asarray(tris)
    inter = np.concatenate([pos, nrm, tan, uv], axis=1).astype(np.float32)          # interleaved vertex buffer, stride 48
    idx16, idx32 = tris.astype(np.uint16).tobytes(), tris.astype(np.uint32).tobytes()
    pad = (-len(idx16)) % 4
    blob = inter.tobytes() + idx16 + b"\0" * pad + idx32
    o16, o32 = inter.nbytes, inter.nbytes + len(idx16) + pad
    rng = np.random.default_rng(7)
    for name, arr in (("base.png", rng.integers(30, 255, (32, 32, 4), dtype=np.uint8)), ("normal.png", rng.integers(90, 165, (32, 32, 4), dtype=np.uint8)),
                      ("mr.png", rng.integers(0, 255, (32, 32, 4), dtype=np.uint8))):
        if name == "normal.png":
            arr[..., 2] = 250
        Image.fromarray(arr, "RGBA").save(tmp_path / name)
    n = len(pos)
    acc = lambda view, off, ctype, count, kind, **kw: dict(bufferView=view, byteOffset=off, componentType=ctype, count=count, type=kind, **kw)
    doc = {
        "asset": {"version": "2.0"}, "scene": 0, "scenes": [{"nodes": [0]}],
        "nodes": [{"children": [1, 2, 3], "translation": [0.5, 0.0, -3.0], "scale": [2.0, 2.0, 2.0]},
                  {"mesh": 0}, {"mesh": 1, "matrix": [1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0.25, 1.0, -0.5, 1]},
                  {"translation": [0, 2, 1], "extensions": {"KHR_lights_punctual": {"light": 0}}}],
        "meshes": [{"primitives": [{"attributes": {"POSITION": 0, "NORMAL": 1, "TANGENT": 2, "TEXCOORD_0": 3}, "indices": 4, "material": 0}]},
                   {"primitives": [{"attributes": {"POSITION": 0}, "indices": 5}]}],
        "materials": [{"doubleSided": True, "alphaMode": "MASK", "alphaCutoff": 0.25, "normalTexture": {"index": 1},
                       "pbrMetallicRoughness": {"baseColorTexture": {"index": 0}, "metallicRoughnessTexture": {"index": 2}}}],
        "textures": [{"source": 0}, {"source": 1}, {"source": 2}],
        "images": [{"uri": "base.png"}, {"uri": "normal.png"}, {"uri": "mr.png"}],
        "extensions": {"KHR_lights_punctual": {"lights": [{"type": "spot", "color": [1, 0.5, 0.25], "intensity": 40.0, "range": 9.0,
                                                            "spot": {"innerConeAngle": 0.2, "outerConeAngle": 0.6}}]}},
        "buffers": [{"byteLength": len(blob), "uri": "data:application/octet-stream;base64," + base64.b64encode(blob).decode()}],
        "bufferViews": [{"buffer": 0, "byteOffset": 0, "byteLength": inter.nbytes, "byteStride": 48},
                        {"buffer": 0, "byteOffset": o16, "byteLength": len(idx16)}, {"buffer": 0, "byteOffset": o32, "byteLength": len(idx32)}],
        "accessors": [acc(0, 0, 5126, n, "VEC3"), acc(0, 12, 5126, n, "VEC3"), acc(0, 24, 5126, n, "VEC4"), acc(0, 40, 5126, n, "VEC2"),
                      acc(1, 0, 5123, tris.size, "SCALAR"), acc(2, 0, 5125, tris.size, "SCALAR")],
    }
    if glb:
        doc["buffers"] = [{"byteLength": len(blob)}]
        js = json.dumps(doc).encode()
        js += b" " * ((-len(js)) % 4)
        bin_ = blob + b"\0" * ((-len(blob)) % 4)
        path = tmp_path / "sheet.glb"
        path.write_bytes(struct.pack("<III", 0x46546C67, 2, 12 + 8 + len(js) + 8 + len(bin_)) + struct.pack("<II", len(js), 0x4E4F534A) + js +
                         struct.pack("<II", len(bin_), 0x004E4942) + bin_)
    else:
        path = tmp_path / "sheet.gltf"
        path.write_text(json.dumps(doc))
    return str(path), pos, tris, uv


@pytest.mark.parametrize("glb", [False, True], ids=["gltf", "glb"])
def test_import_matches_source_arrays(tmp_path, glb):
    path, pos, tris, uv = _write_test_gltf(tmp_path, glb)
    scene = gltf.import_gltf(path, 640, 360)
    assert scene.num_triangles == 2 * len(tris) and len(scene.nodes) == 2
    m = scene.meshlets
    assert (m["NumVertices"] <= 64).all() and (m["NumTriangles"] <= 128).all()
    # every imported triangle is a source triangle (as a set of positions), first mesh
    n0 = scene.nodes[0]
    src = {tuple(np.sort(pos[t].view(np.uint32).reshape(3, 3), axis=0).reshape(-1).tolist()) for t in tris}
    for ml in m[n0.meshlet_offset: n0.meshlet_offset + n0.meshlet_count]:
        for k in range(int(ml["NumTriangles"])):
            p = np.stack([ml["Positions"][:, ml["Indices"][c, k]] for c in range(3)])
            assert tuple(np.sort(p.view(np.uint32), axis=0).reshape(-1).tolist()) in src
    # materials (Scene.cpp:178-183): MASK cutoff 0.25 -> uint8(0.25 * 255 + 0.5) = 64, double sided; meshlets inherit both
    assert int(scene.materials[0]["AlphaCutoff"]) == 64 and int(scene.materials[0]["IsDoubleSided"]) == 1
    first, second = m[: n0.meshlet_count], m[n0.meshlet_count:]
    assert (first["MaterialId"] == 0).all() and (first["AlphaCutoff"] == 64).all()
    assert (second["MaterialId"] == 0xFFFFFFFF).all() and (second["AlphaCutoff"] == 255).all()
    assert (first["TangentHandedness"] != 0).all() and (second["TangentHandedness"] == 0).all()      # tangent.w < 0 everywhere
    # UVs are fp16 pairs of the source UVs
    ml = first[0]
    got = ml["TexCoords"][: int(ml["NumVertices"])].astype(np.uint32)
    assert np.all(np.isin((got & 0xFFFF).astype(np.uint16).view(np.float16), uv[:, 0].astype(np.float16)))
    # node transforms: parent T(0.5,0,-3) * S(2); child 2 adds its own matrix (translation (0.25,1,-0.5))
    assert np.allclose(scene.nodes[0].model[3], [0.5, 0.0, -3.0, 1.0]) and np.allclose(scene.nodes[0].model[0, 0], 2.0)
    assert np.allclose(scene.nodes[1].model[3], [0.5 + 2 * 0.25, 2 * 1.0, -3.0 + 2 * -0.5, 1.0])
    # texture: 32x32, 2 layers (base + normal/MR), mip chain; layer 1 alpha = roughness (G of the MR image), blue = metallic
    t = scene.textures[0]
    assert (t.width, t.height, t.num_layers) == (32, 32, 2) and t.mip_levels == 4
    from PIL import Image
    mr = np.asarray(Image.open(os.path.join(os.path.dirname(path), "mr.png")).convert("RGBA"))
    l1 = tx.get_pixels(t, 1, 0)
    assert np.array_equal((l1 >> 24) & 255, mr[..., 1]) and np.array_equal((l1 >> 16) & 255, mr[..., 2])
    # light: spot at parent * (0,2,1) = (0.5, 4, -1), pointing down -Z of the node, radius 9
    l = scene.lights[0]
    assert int(l["Type"]) == 2 and np.allclose(l["Position"], [0.5, 4.0, -1.0]) and np.allclose(l["Direction"], [0, 0, -1])
    assert np.isclose(l["InvRadiusSq"], 1 / 81.0) and np.isclose(l["SpotScale"], 1.0 / (np.cos(0.2) - np.cos(0.6)), rtol=1e-5)


def test_adjacency_order_fills_meshlets_from_a_shuffled_index_buffer():
    """The meshopt_buildMeshlets stand-in: a 64 x 64-vertex grid whose 7,938 triangles arrive in random order. Cutting
    that stream linearly gives ~20 triangles per 64-vertex meshlet; grown by adjacency it must give > 80, keep every
    triangle exactly once and respect both limits."""
    n = 64
    idx = np.arange(n * n).reshape(n, n)
    a, b, c, d = idx[:-1, :-1].ravel(), idx[:-1, 1:].ravel(), idx[1:, :-1].ravel(), idx[1:, 1:].ravel()
    tris = np.concatenate([np.stack([a, c, b], 1), np.stack([b, c, d], 1)])
    rng = np.random.default_rng(5)
    tris = tris[rng.permutation(len(tris))]
    y, x = np.meshgrid(np.arange(n, dtype=np.float32), np.arange(n, dtype=np.float32), indexing="ij")
    pos = np.stack([x.ravel(), y.ravel(), np.zeros(n * n, dtype=np.float32)], 1)
    order = gltf.adjacency_order(tris)
    assert sorted(order.tolist()) == list(range(len(tris)))
    naive, grown = scenes.meshletize(pos, tris), scenes.meshletize(pos, tris[order])
    for m in (naive, grown):
        assert int(m["NumTriangles"].astype(np.int64).sum()) == len(tris)
        assert (m["NumVertices"] <= 64).all() and (m["NumTriangles"] <= 128).all()
    assert len(tris) / len(naive) < 30 and len(tris) / len(grown) > 80
    # same triangles (as sets of positions), only regrouped
    def tri_set(ms):
        out = set()
        for ml in ms:
            for k in range(int(ml["NumTriangles"])):
                p = np.stack([ml["Positions"][:, ml["Indices"][c, k]] for c in range(3)])
                out.add(tuple(np.sort(p.view(np.uint32), axis=0).reshape(-1).tolist()))
        return out
    assert tri_set(naive) == tri_set(grown)


def test_meshopt_scored_builder_makes_compact_meshlets():
    """gltf.build_meshlets_order (meshopt_buildMeshlets' selection rule, cone_weight 0.25): on the shuffled grid it keeps every
    triangle once, respects both limits, fills the meshlets at least as well as the topological stand-in and — the point of the
    distance / cone score — makes them round: the mean bounding-sphere radius drops by a third or more."""
    n = 64
    idx = np.arange(n * n).reshape(n, n)
    a, b, c, d = idx[:-1, :-1].ravel(), idx[:-1, 1:].ravel(), idx[1:, :-1].ravel(), idx[1:, 1:].ravel()
    tris = np.concatenate([np.stack([a, c, b], 1), np.stack([b, c, d], 1)])
    tris = tris[np.random.default_rng(5).permutation(len(tris))]
    y, x = np.meshgrid(np.arange(n, dtype=np.float32), np.arange(n, dtype=np.float32), indexing="ij")
    pos = np.stack([x.ravel(), y.ravel(), 0.3 * np.sin(x.ravel() * 0.4)], 1).astype(np.float32)
    order = gltf.build_meshlets_order(pos, tris)
    assert sorted(order.tolist()) == list(range(len(tris)))
    assert np.array_equal(order, gltf.build_meshlets_order(pos, tris))                 # deterministic
    scored, plain = scenes.meshletize(pos, tris[order]), scenes.meshletize(pos, tris[gltf.adjacency_order(tris)])
    for m in (scored, plain):
        gltf.compute_meshlet_bounds(m)
        assert int(m["NumTriangles"].astype(np.int64).sum()) == len(tris)
        assert (m["NumVertices"] <= 64).all() and (m["NumTriangles"] <= 128).all()
    assert len(scored) <= len(plain)
    assert float(scored["BoundRadius"].mean()) < 0.67 * float(plain["BoundRadius"].mean())
    # two separate sheets: when nothing adjacent is left the builder continues with the nearest live triangle (k-d tree)
    pos2 = np.concatenate([pos, pos + np.array([0, 0, 50], dtype=np.float32)])
    tris2 = np.concatenate([tris[:400], tris[:400] + n * n])
    o2 = gltf.build_meshlets_order(pos2, tris2)
    assert sorted(o2.tolist()) == list(range(len(tris2)))
    assert len(gltf.build_meshlets_order(pos, tris[:0])) == 0


def test_meshlet_bounds_sphere_and_cone():
    """gltf.compute_meshlet_bounds (meshopt_computeMeshletBounds, Scene.cpp:236-245): the sphere encloses every corner of every
    non-degenerate triangle and is no larger than the bounding-box sphere; the cone axis / cutoff bound every triangle normal, the
    apex lies behind every triangle plane; a meshlet whose normals span more than a hemisphere gets cutoff 1."""
    rng = np.random.default_rng(3)
    # a bumpy sheet (narrow normal cones) and a closed blob (wide ones)
    u, v = np.meshgrid(np.linspace(0, 1, 33), np.linspace(0, 1, 33), indexing="xy")
    sheet = np.stack([u * 4, v * 4, 0.2 * np.sin(u * 5) * np.cos(v * 4)], -1).reshape(-1, 3).astype(np.float32)
    tris = []
    for r in range(32):
        for c in range(32):
            a, b, d, e = r * 33 + c, r * 33 + c + 1, (r + 1) * 33 + c, (r + 1) * 33 + c + 1
            tris += [(a, b, d), (b, e, d)]
    ms = scenes.meshletize(sheet, np.asarray(tris))
    verts, faces = scenes.icosphere(2)
    ms = scenes.concat_meshlets([ms, scenes.meshletize((verts * (1 + 0.1 * rng.random((len(verts), 1)))).astype(np.float32), faces)])
    box = ms.copy()
    scenes.set_bounds(box)
    gltf.compute_meshlet_bounds(ms)
    usable = 0
    for k in range(len(ms)):
        nt = int(ms["NumTriangles"][k])
        pos = ms["Positions"][k].T.astype(np.float64)
        c = pos[ms["Indices"][k][:, :nt].T.astype(int)]
        n = np.cross(c[:, 1] - c[:, 0], c[:, 2] - c[:, 0])
        area = np.linalg.norm(n, axis=1)
        c, n = c[area > 0], n[area > 0] / area[area > 0, None]
        centre, radius = ms["BoundCenter"][k].astype(np.float64), float(ms["BoundRadius"][k])
        assert np.linalg.norm(c.reshape(-1, 3) - centre, axis=1).max() <= radius * (1 + 1e-5)
        assert radius <= float(box["BoundRadius"][k]) * 1.16        # Ritter-style sphere: within ~15 % of the bbox sphere, usually tighter
        if ms["ConeCutoff"][k] < 1:
            usable += 1
            axis = ms["ConeAxis"][k].astype(np.float64)
            mindp = (n @ axis).min()
            assert mindp > 0.1 and abs(np.sqrt(1 - mindp * mindp) - ms["ConeCutoff"][k]) < 1e-4
            assert (((ms["ConeApex"][k].astype(np.float64) - c[:, 0]) * n).sum(axis=1)).max() < 1e-4
            # the reference's back-face cone test shape: a camera inside the cone behind the apex sees no front face
            cam_pos = ms["ConeApex"][k].astype(np.float64) - axis * 3.0
            assert ((n * (c[:, 0] - cam_pos)).sum(axis=1) > 0).all()
    assert usable > len(ms) // 3 and (ms["ConeCutoff"] == 1).any()


def test_combine_normal_mr_and_emissive_mask():
    n = np.array([[[127, 127, 254, 9], [254, 127, 127, 9]]], dtype=np.uint8)          # +Z and +X normals
    mr = np.array([[[1, 2, 3, 4], [5, 6, 7, 8]]], dtype=np.uint8)
    out = gltf.combine_normal_mr(n, mr)
    assert out[0, 0].tolist() == [127, 127, 3, 2] and out[0, 1].tolist() == [254, 127, 7, 6]
    base = np.array([[[9, 9, 9, 255], [9, 9, 9, 100]]], dtype=np.uint8)
    em = np.array([[[0, 0, 0, 0], [0, 9, 0, 0]]], dtype=np.uint8)
    assert gltf.insert_emissive_mask(base, em)[..., 3].tolist() == [[254, 255]]


@pytest.mark.skipif(not os.path.exists(SPONZA), reason="reference assets are not on this machine")
def test_import_reference_sponza_lowpoly(orc):
    scene = gltf.import_gltf(SPONZA)
    assert scene.num_triangles == 63084 and [n.meshlet_count > 0 for n in scene.nodes] == [True, True]
    assert np.allclose(np.diag(scene.nodes[0].model)[:3], 0.008)
    ofb, counters = oracle_render(orc, scene)
    n = scene.width * scene.height
    assert int(counters[0]) == 63084 and (ofb.data[1, :n].view(np.float32) > 0).mean() > 0.9      # RasterBench.cpp:68 looks down the atrium


@pytest.mark.skipif(not os.path.exists(SPONZA), reason="reference assets are not on this machine")
def test_reference_sponza_lowpoly_matches_committed_golden():
    """tests/golden/sponza_lowpoly_hashes.json (make_golden_sponza.py): importer + oracle on the reference's own asset,
    binned and unbinned-with-clipping (the mode RasterBench.cpp:77 uses), from the RasterBench camera."""
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
    import make_golden_sponza
    want = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "sponza_lowpoly_hashes.json")))
    assert make_golden_sponza.digest() == want
    assert want["unbinned_clipped"]["counters"][1] >= want["binned"]["counters"][1]      # clipped pieces are rasterized too


@pytest.mark.gpu
def test_imported_scene_parity(tmp_path, orc, rast_factory):
    path, *_ = _write_test_gltf(tmp_path)
    scene = gltf.import_gltf(path, 640, 360, camera=scenes.cam.Camera(position=(0.4, 0.3, 1.5), euler=(0.1, -0.05), fov_deg=90.0, aspect=640 / 360))
    scene.lights = np.concatenate([scenes.default_light(), scene.lights])
    rast = rast_factory()
    ofb, oc = oracle_render(orc, scene)
    gfb, gc, gscene = gpu_render(rast, scene)
    assert_visbuffer_equal(ofb, gfb, "imported gltf")
    assert int(oc[1]) > 0 and [gc["TrianglesProcessed"], gc["TrianglesRasterized"]] == [int(oc[0]), int(oc[1])]
    uni = scenes.resolve_uniforms(scene, scene.nodes[0])
    orc.resolve(ofb, scene.meshlets, scene.materials, scene.textures, scene.lights, **uni)
    rast.resolve(gfb, gscene, **uni)
    n = scene.width * scene.height
    a, b = gfb.download_tiled(0).view(np.uint8).astype(np.int32), ofb.data[0, :n].view(np.uint8).astype(np.int32)
    assert np.abs(a - b).max() <= 2
