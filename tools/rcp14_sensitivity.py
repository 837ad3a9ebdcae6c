#!/usr/bin/env python
"""How far would the upstream BINARY be from the canonical arithmetic?

The reference is compiled with -ffast-math; its vector `1.0f / w` (perspective divide, SIMD.h:473-476) and
`16.0f / det` (Rasterizer.cpp:320) are then most likely vrcp14ps + one Newton-Raphson step, not IEEE divisions. Parity
here is defined against the IEEE arithmetic (mode 0). This tool renders the same scenes with the oracle in both modes
on the host CPU and reports how many depth words / surface ids / covered pixels / counters differ — the size of the
gap nobody can close without the upstream compiler. CPU only; needs AVX-512F for the rcp14 instruction.

    python tools/rcp14_sensitivity.py
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from glimpsw_b200 import scenes  # noqa: E402
from oracle import orc  # noqa: E402
from helpers import oracle_render  # noqa: E402


def compare(scene, cull=False, mode=1):
    """mode: bit 0 = rcp14 + Newton reciprocals, bit 1 = contracted cull determinant (oracle.orc.set_reciprocal_mode)."""
    n = scene.width * scene.height
    out = []
    for mode in (0, mode):
        if not orc.set_reciprocal_mode(mode):
            raise SystemExit("this host has no AVX-512F: cannot execute vrcp14ss")
        fb, counters = oracle_render(orc, scene, cull=cull)
        out.append((fb.data[1, :n].copy(), fb.data[0, :n].copy(), [int(c) for c in counters[:3]]))
    orc.set_reciprocal_mode(0)
    (d0, i0, c0), (d1, i1, c1) = out
    cov0, cov1 = d0.view(np.float32) > 0, d1.view(np.float32) > 0
    ulp = np.abs(d0.astype(np.int64) - d1.astype(np.int64))[cov0 & cov1]
    return {"pixels": n, "covered": int(cov0.sum()), "depth_words_differ": int((d0 != d1).sum()),
            "depth_differ_fraction_of_covered": round(float((d0 != d1).sum()) / max(int(cov0.sum()), 1), 5),
            "max_depth_ulp": int(ulp.max()) if len(ulp) else 0, "ids_differ": int((i0 != i1).sum()),
            "coverage_differs": int((cov0 != cov1).sum()), "counters_ieee": c0, "counters_rcp14_nr": c1}


if __name__ == "__main__":
    orc.build()
    cases = {
        "c2_grid_1M_1080p": (scenes.grid_scene, False),
        "c1_knot_72k_1080p": (scenes.torus_knot_scene, False),
        "c4_small_instanced_720p": (lambda: scenes.instanced_scene(subdivisions=4, instances=27, width=1280, height=720), True),
        "room_big_triangles_1080p": (scenes.room_scene, False),
    }
    report = {}
    for name, (make, cull) in cases.items():
        report[name] = {"rcp14_newton": compare(make(), cull, 1), "contracted_determinant": compare(make(), cull, 2),
                        "both": compare(make(), cull, 3)}
    print(json.dumps(report, indent=1))
