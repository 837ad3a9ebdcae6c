"""Generates the fixtures that pin the bench workloads (glimpsw_b200/workloads.py) to the CPU oracle:

  tests/golden/sponza_lowpoly_scene.npz   the reference's own asset (assets/models/Sponza/Sponza_LowPoly.gltf) imported with
                                          glimpsw_b200.gltf, so that config C1's glTF case exists where /root/reference does not
                                          (the GPU box); only rewritten where the asset is present
  tests/golden/bench_configs.json         per workload: SHA-256 of the oracle's depth and surface-id layers (binned mode, the
                                          frame loop of Main.cpp:213-240), the three integer counters, covered pixels, and —
                                          where the workload resolves — the name of the colour fixture
  tests/golden/colour_<name>.npz          the oracle's resolved RGB image (u8, zlib; pixels [::2, ::2]) for the colour gate (<= 2/255)

    python tests/golden/make_golden_bench.py [name ...]

bench.py and tests/test_bench_configs_gpu.py compare the CUDA path against these on the GPU box.
"""
import hashlib
import json
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

ASSET = "/root/reference/assets/models/Sponza/Sponza_LowPoly.gltf"
COLOUR_VIEWS = {"c4_views": [0], "c5_views": [0]}      # views whose resolved image is committed
COLOUR_STRIDE = 2          # the colour fixtures keep every second pixel of every second row (a quarter of the bytes)
HASH_VIEWS = {"c4_views": list(range(64)), "c5_views": [0, 21, 42, 63]}


def oracle_frame(orc, wl, view=None):
    """Clear -> CullMeshlets -> DrawMeshlets per node -> [Resolve] on the CPU oracle. Returns a dict for the fixture + colour."""
    from glimpsw_b200 import workloads
    scene = wl.scene
    if view is not None:
        scene.camera = wl.cameras[view]
    fb = orc.Framebuffer(scene.width, scene.height)
    fb.clear(0xFF000000, 0.0)
    counters = np.zeros(4, dtype=np.uint64)
    proj, vm = scene.view_proj()
    visible = 0
    for node in scene.nodes:
        bitmap = None
        if wl.fused_cull:
            planes = orc.frustum_planes(proj, vm, node.model)
            bitmap, nvis = orc.cull_meshlets(scene.meshlets[node.meshlet_offset:node.meshlet_offset + node.meshlet_count], planes)
            visible += int(nvis)
        orc.draw_meshlets(fb, scene.meshlets, node.meshlet_offset, node.meshlet_count, scene.object_to_clip(node), cull_bitmap=bitmap,
                          materials=scene.materials, counters=counters, textures=scene.textures if len(scene.textures) else None)
    n = scene.width * scene.height
    out = {"depth_sha256": hashlib.sha256(fb.data[1, :n].tobytes()).hexdigest(),
           "id_sha256": hashlib.sha256(fb.data[0, :n].tobytes()).hexdigest(),
           "counters": [int(c) for c in counters[:3]],
           "covered_pixels": int((fb.data[1, :n].view(np.float32) > 0).sum()),
           "meshlets_visible": visible if wl.fused_cull else len(scene.meshlets)}
    colour = None
    if wl.resolve:
        orc.resolve(fb, scene.meshlets, scene.materials, scene.textures, scene.lights, **workloads.view_uniforms(wl, view))
        colour = fb.get_pixels(0).view(np.uint8).reshape(scene.height, scene.width, 4)[::COLOUR_STRIDE, ::COLOUR_STRIDE, :3].copy()
    return out, colour


def main(names):
    from glimpsw_b200 import workloads
    from oracle import orc
    orc.build()
    if os.path.exists(ASSET):
        from glimpsw_b200 import gltf
        workloads.save_sponza_fixture(gltf.import_gltf(ASSET))
    path = workloads.GOLDEN
    golden = json.load(open(path)) if os.path.exists(path) else {}
    for name in names:
        t0 = time.time()
        wl = workloads.build(name)
        entry = {"description": wl.description, "triangles": wl.scene.num_triangles, "meshlets": len(wl.scene.meshlets),
                 "size": [wl.scene.width, wl.scene.height],
                 "meshlets_sha256": hashlib.sha256(np.ascontiguousarray(wl.scene.meshlets).tobytes()).hexdigest()}
        if wl.cameras is None:
            frame, colour = oracle_frame(orc, wl)
            entry.update(frame)
            if colour is not None:
                entry["colour"] = f"colour_{name}.npz"
                np.savez_compressed(os.path.join(HERE, entry["colour"]), rgb=colour, stride=COLOUR_STRIDE)
        else:
            entry["views"] = {}
            for v in HASH_VIEWS[name]:
                frame, colour = oracle_frame(orc, wl, v)
                if v in COLOUR_VIEWS[name]:
                    frame["colour"] = f"colour_{name}_v{v}.npz"
                    np.savez_compressed(os.path.join(HERE, frame["colour"]), rgb=colour, stride=COLOUR_STRIDE)
                entry["views"][str(v)] = frame
                print(f"  {name} view {v}: {frame['counters']} ({time.time() - t0:.0f} s)", flush=True)
        golden[name] = entry
        json.dump(golden, open(path, "w"), indent=1, sort_keys=True)
        print(f"{name}: done in {time.time() - t0:.0f} s", flush=True)


if __name__ == "__main__":
    main(sys.argv[1:] or ["c1_knot", "c1_sponza", "c2_grid", "c3_knot", "c4_views", "c5_views"])
