"""Generates tests/golden/visbuffer_hashes.json from the CPU oracle.

The reference has no golden vectors of its own, so these are regression pins of OUR
oracle: SHA-256 of the depth and surface-id layers plus the integer perf counters for reduced-size
versions of the BASELINE configs and the full C2 config. Re-run only when the oracle changes on purpose:

    python tests/golden/make_golden.py
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
    "c2_small_640x360": lambda: (scenes.grid_scene(20, 16, 640, 360, seed=3, flip_fraction=0.2), False),
    "c2_full_1920x1080": lambda: (scenes.grid_scene(), False),
    "c4_small_culled_1280x720": lambda: (scenes.instanced_scene(subdivisions=3, instances=27, width=1280, height=720), True),
    "c1_knot_960x540": lambda: (scenes.torus_knot_scene(120, 48, 960, 540, tex_size=64), False),
    "c3_odd_size_1000x564": lambda: (scenes.grid_scene(16, 12, 1000, 564, seed=8), False),
    "room_big_tris_1920x1080": lambda: (scenes.room_scene(), False),
    "room_clipped_1920x1080": lambda: (scenes.room_scene(), False),
    "knot_alpha_closeup_clipped_960x540": lambda: (scenes.closeup_alpha_scene(), False),
}
# Oracle mode per case (default: the binned path, non-trivial triangles counted and dropped). The *_clipped cases
# pin the unbinned path with EnableClipping (Clipper::ClipTriangles, Rasterizer.cpp:398-491).
MODES = {
    "room_clipped_1920x1080": dict(binned=False, clipping=True),
    "knot_alpha_closeup_clipped_960x540": dict(binned=False, clipping=True),
}


def digest(scene, cull, **mode):
    fb, counters = oracle_render(orc, scene, cull=cull, **mode)
    n = scene.width * scene.height
    return {"depth_sha256": hashlib.sha256(fb.data[1, :n].tobytes()).hexdigest(),
            "id_sha256": hashlib.sha256(fb.data[0, :n].tobytes()).hexdigest(),
            "counters": [int(c) for c in counters[:3]], "triangles": scene.num_triangles,
            "covered_pixels": int((fb.data[1, :n].view(np.float32) > 0).sum())}


if __name__ == "__main__":
    out = {name: digest(*make(), **MODES.get(name, {})) for name, make in CASES.items()}
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "visbuffer_hashes.json")
    json.dump(out, open(path, "w"), indent=1, sort_keys=True)
    print(json.dumps(out, indent=1))
