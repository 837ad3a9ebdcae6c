"""Developer timing probe (not the contract bench): per-mode frame and stage times on one scene."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from glimpsw_b200 import api, scenes


def run(scene, binning, steps=20, fused_cull=False, cull=False, all_culled=False):
    rast = api.Rasterizer(0, enable_binning=binning, fused_frustum_cull=fused_cull)
    gscene = rast.upload_scene(scene.meshlets, scene.materials, scene.textures, scene.lights)
    fb = rast.create_framebuffer(scene.width, scene.height)
    proj, view = scene.view_proj()
    draws = []
    for n in scene.nodes:
        d = dict(offset=n.meshlet_offset, count=n.meshlet_count, object_to_clip=scene.object_to_clip(n))
        if all_culled:
            d["cull_bitmap"] = np.zeros((n.meshlet_count + 15) // 16, dtype=np.uint16)
        if fused_cull:
            d["planes"] = rast.frustum_planes(proj, view, n.model)
        draws.append(d)
    batch = rast.make_batch(draws)
    def frame():
        fb.clear(0xFF000000, 0.0)
        rast.draw_prebuilt(fb, gscene, batch)
    for _ in range(3):
        frame()
    rast.sync()
    times = []
    for _ in range(steps):
        rast.flush_l2()
        rast.timer_begin()
        frame()
        times.append(rast.timer_end())
    rast.enable_stage_timing(True)
    frame()
    st = rast.stage_times_us()
    rast.enable_stage_timing(False)
    rast.reset_counters(); frame(); c = rast.counters(); c.update(rast.draw_stats())
    tris = scene.num_triangles
    med = float(np.median(times))
    print(f"{scene.name} binning={binning} fused_cull={fused_cull}: median {med*1000:.1f} us  min {min(times)*1000:.1f} us  "
          f"{tris/med/1e6:.2f} Gtri/s  stages(us)={ {k: round(v[0],1) for k,v in st.items()} }  counters={c}")
    rast.destroy()


if __name__ == "__main__":
    which = sys.argv[1] if len(sys.argv) > 1 else "c2"
    if which in ("c2", "all"):
        s = scenes.grid_scene()
        run(s, True); run(s, False)
    if which == "floor":
        s = scenes.grid_scene()
        run(s, True, all_culled=True); run(s, False, all_culled=True)
    if which in ("c4", "all"):
        s = scenes.instanced_scene()
        print("c4 tris", s.num_triangles, "meshlets", len(s.meshlets))
        run(s, True, fused_cull=True); run(s, False, fused_cull=True); run(s, True)
