"""Minimal frame loop for ncu captures: N frames of the bench workload (clear + draw + resolve)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from glimpsw_b200 import api, scenes  # noqa: E402

frames = int(sys.argv[1]) if len(sys.argv) > 1 else 5
mode = sys.argv[2] if len(sys.argv) > 2 else "binned"
which = sys.argv[3] if len(sys.argv) > 3 else "c2"
scene = bench.build_workload(0) if which == "c2" else scenes.instanced_scene()
rast = api.Rasterizer(0, enable_binning=(mode == "binned"), fused_frustum_cull=(which != "c2"))
gscene = rast.upload_scene(scene.meshlets, scene.materials, scene.textures, scene.lights)
fb = rast.create_framebuffer(scene.width, scene.height)
proj, view = scene.view_proj()
batch = rast.make_batch([dict(offset=n.meshlet_offset, count=n.meshlet_count, object_to_clip=scene.object_to_clip(n),
                              planes=rast.frustum_planes(proj, view, n.model)) for n in scene.nodes])
uni = scenes.resolve_uniforms(scene, scene.nodes[0])
for _ in range(frames):
    rast.flush_l2()
    fb.clear(0xFF000000, 0.0)
    rast.draw_prebuilt(fb, gscene, batch)
    if which == "c2":
        rast.resolve(fb, gscene, **uni)
rast.sync()
print("frames", frames, rast.counters())
