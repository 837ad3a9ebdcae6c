"""Shared helpers for the parity tests: run one scene through the oracle and through the CUDA path."""
import numpy as np


def oracle_render(orc, scene, cull=False, clear=(0xFF000000, 0.0), fb=None, binned=True, clipping=False):
    """Frame loop of Main.cpp:213-240 on the CPU oracle. Returns (fb, counters).
    binned=False: DrawMeshletsST's handling of non-trivial triangles (clipped with `clipping`, else dropped uncounted)."""
    if fb is None:
        fb = orc.Framebuffer(scene.width, scene.height)
        fb.clear(*clear)
    counters = np.zeros(4, dtype=np.uint64)
    proj, view = scene.view_proj()
    for node in scene.nodes:
        bitmap = None
        if cull:
            planes = orc.frustum_planes(proj, view, node.model)
            bitmap, _ = orc.cull_meshlets(scene.meshlets[node.meshlet_offset:node.meshlet_offset + node.meshlet_count], planes)
        orc.draw_meshlets(fb, scene.meshlets, node.meshlet_offset, node.meshlet_count, scene.object_to_clip(node),
                          cull_bitmap=bitmap, materials=scene.materials, counters=counters,
                          textures=scene.textures if len(scene.textures) else None, binned=binned, clipping=clipping)
    return fb, counters


def raster_mode(rast):
    """(binned, clipping) keyword arguments for the oracle that match a Rasterizer mirror's flags."""
    from glimpsw_b200 import api
    return dict(binned=bool(rast.flags & api.FLAG_BINNING), clipping=bool(rast.flags & api.FLAG_CLIPPING))


def gpu_render(rast, scene, cull=False, clear=(0xFF000000, 0.0), batch=True, gscene=None, fb=None):
    """Same frame through libswrb.so. Returns (fb, counters dict, gscene)."""
    if gscene is None:
        gscene = rast.upload_scene(scene.meshlets, scene.materials, scene.textures, scene.lights)
    if fb is None:
        fb = rast.create_framebuffer(scene.width, scene.height)
        fb.clear(*clear)
    rast.reset_counters()
    proj, view = scene.view_proj()
    draws = []
    for node in scene.nodes:
        bitmap = None
        if cull:
            bitmap, _ = rast.cull_meshlets(gscene, node.meshlet_offset, node.meshlet_count, proj, view, node.model)
        draws.append(dict(offset=node.meshlet_offset, count=node.meshlet_count,
                          object_to_clip=scene.object_to_clip(node), cull_bitmap=bitmap))
    if batch:
        rast.draw_batch(fb, gscene, draws)
    else:
        for d in draws:
            rast.draw_meshlets(fb, gscene, d["offset"], d["count"], d["object_to_clip"], cull_bitmap=d["cull_bitmap"])
    return fb, rast.counters(), gscene


def assert_visbuffer_equal(ofb, gfb, what=""):
    n = ofb.width * ofb.height
    gd = gfb.download_tiled(1)
    gc = gfb.download_tiled(0)
    od, oc = ofb.data[1, :n], ofb.data[0, :n]
    bad_d = int((gd != od).sum())
    bad_c = int((gc != oc).sum())
    assert bad_d == 0 and bad_c == 0, f"{what}: {bad_d} depth and {bad_c} id pixels differ of {n}"
