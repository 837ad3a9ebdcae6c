"""Developer probe: where the CPU baseline's frame time goes on this host (draw vs resolve, by thread count)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from glimpsw_b200 import scenes
from oracle import orc
orc.build()
scene = bench.build_workload(0); node = scene.nodes[0]; m = scene.object_to_clip(node); uni = scenes.resolve_uniforms(scene, node)
for th in (1, 4, 8, 16, 0):
    base = orc.Baseline(th)
    fb = orc.Framebuffer(scene.width, scene.height)
    td, tr = [], []
    for i in range(6):
        base.clear(fb, 0xFF000000, 0.0)
        t0 = time.perf_counter(); base.draw_meshlets(fb, scene.meshlets, 0, len(scene.meshlets), m, materials=scene.materials); t1 = time.perf_counter()
        base.resolve(fb, scene.meshlets, scene.materials, scene.textures, scene.lights, **uni); t2 = time.perf_counter()
        td.append(t1 - t0); tr.append(t2 - t1)
    print(f"threads {base.threads}: draw {min(td)*1e3:.1f} ms  resolve {min(tr)*1e3:.1f} ms")
    base.close()
