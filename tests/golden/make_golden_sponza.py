"""Generates tests/golden/sponza_lowpoly_hashes.json: the one asset of the reference checkout that loads
(assets/models/Sponza/Sponza_LowPoly.gltf, 63,084 position-only triangles, 2 nodes) imported with glimpsw_b200.gltf and
rendered by the CPU oracle from the camera RasterBench.cpp:68 hard-codes, 1920x1080 (BASELINE config C1's glTF case).

The asset only exists where /root/reference is mounted (this container, not the GPU box), so the fixture pins importer +
oracle here; tests/test_gltf_import.py re-checks it whenever the asset is present.

    python tests/golden/make_golden_sponza.py
"""
import hashlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

ASSET = "/root/reference/assets/models/Sponza/Sponza_LowPoly.gltf"


def digest():
    from glimpsw_b200 import gltf
    from oracle import orc
    from helpers import oracle_render
    orc.build()
    scene = gltf.import_gltf(ASSET)
    out = {"triangles": scene.num_triangles, "meshlets": len(scene.meshlets),
           "nodes": [[n.meshlet_offset, n.meshlet_count] for n in scene.nodes],
           "meshlets_sha256": hashlib.sha256(np.ascontiguousarray(scene.meshlets).tobytes()).hexdigest()}
    n = scene.width * scene.height
    for mode, kw in (("binned", dict(binned=True)), ("unbinned_clipped", dict(binned=False, clipping=True))):
        fb, counters = oracle_render(orc, scene, **kw)
        out[mode] = {"depth_sha256": hashlib.sha256(fb.data[1, :n].tobytes()).hexdigest(),
                     "id_sha256": hashlib.sha256(fb.data[0, :n].tobytes()).hexdigest(),
                     "counters": [int(c) for c in counters[:3]],
                     "covered_pixels": int((fb.data[1, :n].view(np.float32) > 0).sum())}
    return out


if __name__ == "__main__":
    out = digest()
    json.dump(out, open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "sponza_lowpoly_hashes.json"), "w"), indent=1, sort_keys=True)
    print(json.dumps(out, indent=1))
