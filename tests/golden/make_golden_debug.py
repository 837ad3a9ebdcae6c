"""Generates tests/golden/debug_program_hashes.json from the CPU oracle: regression pins for the OverdrawShader program
(FS_Overdraw, Shading.cpp:333-342) and the bit-exact layers of ShadingContext::ResolveDebug (Shading.cpp:734-773).
Like visbuffer_hashes.json these pin OUR oracle (the reference ships no vectors). Re-run only when the
oracle changes on purpose:

    python tests/golden/make_golden_debug.py
"""
import hashlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from glimpsw_b200 import scenes  # noqa: E402
from oracle import orc  # noqa: E402
from helpers import oracle_render  # noqa: E402

CASES = {
    "knot_640x360": lambda: scenes.torus_knot_scene(60, 24, 640, 360, tex_size=128),
    "instanced_small_1280x720": lambda: scenes.instanced_scene(subdivisions=3, instances=27, width=1280, height=720),
    "room_big_tris_960x540": lambda: scenes.room_scene(960, 540),
}
EXACT_LAYERS = ["MeshletId", "TriangleId"]
OVERDRAW_LAYERS = ["OverdrawPixel", "OverdrawQuad"]


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def oracle_overdraw(scene):
    fb = orc.Framebuffer(scene.width, scene.height)
    fb.clear(0, 0.0)
    counters = np.zeros(4, dtype=np.uint64)
    for nd in scene.nodes:
        orc.draw_meshlets(fb, scene.meshlets, nd.meshlet_offset, nd.meshlet_count, scene.object_to_clip(nd), materials=scene.materials,
                          counters=counters, overdraw=True)
    return fb, counters


def digest(scene):
    n = scene.width * scene.height
    uni = scenes.resolve_uniforms(scene, scene.nodes[0])
    out = {}
    od, counters = oracle_overdraw(scene)
    col = od.data[0, :n]
    out["overdraw"] = {"counter_sha256": sha(col), "depth_sha256": sha(od.data[1, :n]), "counters": [int(c) for c in counters[:3]],
                       "covered_pixel_visits": int((col >> 16).astype(np.int64).sum()), "helper_lane_visits": int((col & 0xFFFF).astype(np.int64).sum()),
                       "max_overdraw": int((col >> 16).max())}
    for layer in OVERDRAW_LAYERS:
        fb = orc.Framebuffer(scene.width, scene.height)
        fb.data[...] = od.data
        orc.resolve_debug(fb, scene.meshlets, scene.materials, scene.textures, layer, **uni)
        out[layer] = sha(fb.data[0, :n])
    vis, _ = oracle_render(orc, scene)
    for layer in EXACT_LAYERS:
        fb = orc.Framebuffer(scene.width, scene.height)
        fb.data[...] = vis.data
        orc.resolve_debug(fb, scene.meshlets, scene.materials, scene.textures, layer, **uni)
        out[layer] = sha(fb.data[0, :n])
    return out


if __name__ == "__main__":
    orc.build()
    out = {name: digest(make()) for name, make in CASES.items()}
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "debug_program_hashes.json")
    json.dump(out, open(path, "w"), indent=1, sort_keys=True)
    print(json.dumps(out, indent=1))
