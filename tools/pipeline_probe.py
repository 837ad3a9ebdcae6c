"""Probe: how much does keeping 2 (or more) frames in flight on separate streams buy on one GPU?"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
from glimpsw_b200 import api, scenes

scene = bench.build_workload(0)
node = scene.nodes[0]
uni = scenes.resolve_uniforms(scene, node)
tris = scene.num_triangles


def run(in_flight, copies_per_rast, frames=600):
    rasts = [api.Rasterizer(0) for _ in range(in_flight)]
    ctx = []
    for r in rasts:
        gs = [r.upload_scene(scene.meshlets, scene.materials, scene.textures, scene.lights) for _ in range(copies_per_rast)]
        fb = r.create_framebuffer(scene.width, scene.height)
        batch = r.make_batch([dict(offset=node.meshlet_offset, count=node.meshlet_count, object_to_clip=scene.object_to_clip(node))])
        ctx.append((r, gs, fb, batch))

    def frame(k):
        r, gs, fb, batch = ctx[k % in_flight]
        g = gs[(k // in_flight) % copies_per_rast]
        fb.clear(0xFF000000, 0.0)
        r.draw_prebuilt(fb, g, batch)
        r.resolve(fb, g, **uni)

    for k in range(20):
        frame(k)
    for r in rasts:
        r.sync()
    t0 = time.perf_counter()
    for k in range(frames):
        frame(k)
    for r in rasts:
        r.sync()
    dt = time.perf_counter() - t0
    print(f"in_flight={in_flight} scene copies={in_flight * copies_per_rast}: {dt / frames * 1e6:.1f} us/frame  {tris * frames / dt / 1e9:.2f} Gtri/s")
    for r in rasts:
        r.destroy()


for nf, cp in ((1, 8), (2, 4), (3, 3), (4, 2)):
    run(nf, cp)
