"""Per-view cost of the bench batch on one GPU (frames back to back, one context), and what the static deal `v mod N` loses to
the cost differences between views at N = 2, 4, 8 compared with a cost-aware deal (longest processing time first).

    python tools/view_costs.py [c4_views|c5_views]
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from glimpsw_b200 import api, sharding, workloads  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "c4_views"
wl = workloads.build(name)
scene = wl.scene
stream = torch.cuda.Stream()
torch.cuda.set_stream(stream)
rast = api.Rasterizer(0, fused_frustum_cull=True)
rast.set_stream(stream.cuda_stream)
gscene = rast.upload_scene(scene.meshlets, scene.materials, scene.textures, scene.lights)
fb = rast.create_framebuffer(scene.width, scene.height)
costs = []
for v in range(workloads.NUM_VIEWS):
    batch = rast.create_batch(gscene, workloads.view_draws(rast, wl, v))
    frame = rast.make_frame(batch, api.Rasterizer.make_uniforms(**workloads.view_uniforms(wl, v)))
    for _ in range(2):
        rast.submit_frame(fb, frame)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record(stream)
    for _ in range(8):
        rast.submit_frame(fb, frame)
    e1.record(stream)
    torch.cuda.synchronize()
    costs.append(e0.elapsed_time(e1) / 8 * 1e3)
costs = np.array(costs)
out = {"workload": name, "us_per_view": [round(float(c), 1) for c in costs], "mean_us": round(float(costs.mean()), 1),
       "min_us": round(float(costs.min()), 1), "max_us": round(float(costs.max()), 1), "deals": {}}
for n in (2, 4, 8):
    mod = [costs[sharding.views_for_rank(len(costs), r, n)].sum() for r in range(n)]
    lpt = np.zeros(n)
    for c in sorted(costs, reverse=True):
        lpt[lpt.argmin()] += c
    ideal = costs.sum() / n
    out["deals"][str(n)] = {"v_mod_N_efficiency": round(float(ideal / max(mod)), 4), "lpt_efficiency": round(float(ideal / lpt.max()), 4)}
print(json.dumps(out))
rast.destroy()
