"""Minimal frame loop for ncu captures: N frames of a named workload (glimpsw_b200/workloads.py) through swrb_frame_submit.

    python tools/profile_frame.py <workload> <frames> [view] [binned|direct]
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from glimpsw_b200 import api, workloads  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "c4_views"
frames = int(sys.argv[2]) if len(sys.argv) > 2 else 4
view = int(sys.argv[3]) if len(sys.argv) > 3 else 0
mode = sys.argv[4] if len(sys.argv) > 4 else "binned"
wl = workloads.build(name)
scene = wl.scene
rast = api.Rasterizer(0, enable_binning=(mode == "binned"), fused_frustum_cull=wl.fused_cull)
gscene = rast.upload_scene(scene.meshlets, scene.materials if len(scene.materials) else None, scene.textures, scene.lights)
fb = rast.create_framebuffer(scene.width, scene.height)
v = view if wl.cameras is not None else None
batch = rast.create_batch(gscene, workloads.view_draws(rast, wl, v))
uni = api.Rasterizer.make_uniforms(**workloads.view_uniforms(wl, v)) if wl.resolve else None
frame = rast.make_frame(batch, uni)
for _ in range(frames):
    rast.flush_l2()
    rast.submit_frame(fb, frame)
rast.sync()
print("frames", frames, rast.counters(), rast.draw_stats())
