"""Developer probe for the e2e number: PCIe copy rates on this box and the e2e loop at several pipeline depths."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import bench
from glimpsw_b200 import api, scenes


def copy_rates():
    n = 64 << 20
    h = torch.empty(n, dtype=torch.uint8).pin_memory()
    h2 = torch.empty(n, dtype=torch.uint8).pin_memory()
    d = torch.empty(n, dtype=torch.uint8, device="cuda")
    d2 = torch.empty(n, dtype=torch.uint8, device="cuda")
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()

    def timeit(fn, reps=10):
        fn(); torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(reps):
            fn()
        torch.cuda.synchronize()
        return (time.perf_counter() - t0) / reps

    def h2d():
        with torch.cuda.stream(s1):
            d.copy_(h, non_blocking=True)

    def d2h():
        with torch.cuda.stream(s2):
            h2.copy_(d2, non_blocking=True)

    def both():
        h2d(); d2h()

    t = timeit(h2d); print(f"H2D 64 MiB pinned: {n / t / 1e9:.1f} GB/s")
    t = timeit(d2h); print(f"D2H 64 MiB pinned: {n / t / 1e9:.1f} GB/s")
    t = timeit(both); print(f"H2D + D2H concurrently: {n / t / 1e9:.1f} GB/s each direction")


def e2e(F, steps=150):
    scene = bench.build_workload(0)
    node = scene.nodes[0]
    uni_c = api.Rasterizer.make_uniforms(**scenes.resolve_uniforms(scene, node))
    ctxs = []
    for i in range(F):
        r = api.Rasterizer(0)
        c = type("C", (), {})()
        c.rast, c.fb = r, r.create_framebuffer(scene.width, scene.height)
        c.scene = r.upload_scene(scene.meshlets, scene.materials, scene.textures, scene.lights)
        c.batch = r.make_batch([dict(offset=node.meshlet_offset, count=node.meshlet_count, object_to_clip=scene.object_to_clip(node))])
        c.host_meshlets = r.alloc_pinned(scene.meshlets.shape, scene.meshlets.dtype)
        c.host_meshlets[...] = scene.meshlets
        c.host_image = r.alloc_pinned((scene.height, scene.width), np.uint32)
        c.uses = 0
        ctxs.append(c)

    def frame(k):
        c = ctxs[k % F]
        if c.uses:
            c.rast.sync()
        c.uses += 1
        c.scene.update_meshlets(c.host_meshlets, 0)
        c.fb.clear(0xFF000000, 0.0)
        c.rast.draw_prebuilt(c.fb, c.scene, c.batch)
        c.rast.resolve_prebuilt(c.fb, c.scene, uni_c)
        c.fb.get_pixels_async(0, c.host_image)

    for k in range(3 * F):
        frame(k)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for k in range(steps):
        frame(k)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / steps
    print(f"e2e F={F}: {dt * 1e3:.4f} ms/step  {scene.num_triangles / dt / 1e6:.0f} Mtri/s  "
          f"(H2D {scene.meshlets.nbytes / dt / 1e9:.1f} GB/s, D2H {scene.width * scene.height * 4 / dt / 1e9:.1f} GB/s)")
    for c in ctxs:
        c.rast.destroy()


if __name__ == "__main__":
    copy_rates()
    for F in (1, 2, 3, 4, 6):
        e2e(F)
