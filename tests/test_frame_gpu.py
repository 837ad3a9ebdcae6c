"""Prepared batches / whole-frame submission (swrb_batch_*, swrb_frame_submit), the key order that decides exact depth ties,
and stream ordering of GetPixels on side streams — CUDA path through the C ABI vs the CPU oracle. Needs a GPU."""
import numpy as np
import pytest

from glimpsw_b200 import api, scenes
from glimpsw_b200.layout import MATERIAL_DTYPE
from helpers import oracle_render, gpu_render, assert_visbuffer_equal
from test_oracle_kat import meshlet_from_clip_tris, tri_px, IDENT

pytestmark = pytest.mark.gpu


def _colour_close(a, b, tol=2):
    return int(np.abs(a.view(np.uint8).astype(np.int32) - b.view(np.uint8).astype(np.int32)).max()) <= tol


@pytest.mark.parametrize("binning", [True, False], ids=["binned", "direct"])
def test_frame_submit_equals_the_call_by_call_loop(orc, rast_factory, binning):
    """Clear -> DrawMeshlets -> Resolve -> GetPixels as ONE swrb_frame_submit, repeated: every frame equals the oracle's
    (vis-buffer exact, colour <= 2/255), although from the second frame on the key seeds and the binner's counters come
    from the previous frame's resolve pass instead of a k_frame_begin launch."""
    scene = scenes.torus_knot_scene(120, 48, 960, 540, tex_size=256)
    node = scene.nodes[0]
    ofb, oc = oracle_render(orc, scene)
    n = scene.width * scene.height
    vis_d, vis_i = ofb.data[1, :n].copy(), ofb.data[0, :n].copy()
    uni = scenes.resolve_uniforms(scene, node)
    orc.resolve(ofb, scene.meshlets, scene.materials, scene.textures, scene.lights, **uni)
    want = ofb.get_pixels(0)

    rast = rast_factory(enable_binning=binning)
    gscene = rast.upload_scene(scene.meshlets, scene.materials, scene.textures, scene.lights)
    fb = rast.create_framebuffer(scene.width, scene.height)
    batch = rast.create_batch(gscene, [dict(offset=node.meshlet_offset, count=node.meshlet_count, object_to_clip=scene.object_to_clip(node))])
    # the prepared batch alone == swrb_draw_batch
    fb.clear(0xFF000000, 0.0)
    rast.reset_counters()
    rast.draw_prepared(fb, batch)
    assert_visbuffer_equal(_FakeFb(scene, vis_i, vis_d), fb, "prepared batch")
    c = rast.counters()
    assert [c["TrianglesProcessed"], c["TrianglesRasterized"], c["TrianglesClipped"]] == [int(oc[0]), int(oc[1]), int(oc[2])]

    uni_c = api.Rasterizer.make_uniforms(**uni)
    host = rast.alloc_pinned((scene.height, scene.width), np.uint32)
    frame = rast.make_frame(batch, uni_c, 0xFF000000, 0.0, pixels_host=host)
    launches = []
    for k in range(4):
        host[...] = 0
        l0 = rast.launch_count()
        rast.submit_frame(fb, frame)
        launches.append(rast.launch_count() - l0)
        rast.sync()
        assert _colour_close(host, want), f"frame {k}"
        # the resolve pass retired the keys: the depth layer is current without a key unpack
        assert np.array_equal(fb.download_tiled(1), vis_d), f"depth layer after frame {k}"
    assert launches[1] < launches[0], launches      # no k_frame_begin from the second frame on
    assert launches[1] == launches[2] == launches[3]
    # a different clear depth invalidates the seeds the resolve pass left behind: k_frame_begin runs again, result still right
    ofb2 = orc.Framebuffer(scene.width, scene.height)
    ofb2.clear(0xFF000000, 0.25)
    orc.draw_meshlets(ofb2, scene.meshlets, node.meshlet_offset, node.meshlet_count, scene.object_to_clip(node), materials=scene.materials,
                      textures=scene.textures)
    fb.clear(0xFF000000, 0.25)
    rast.draw_prepared(fb, batch)
    assert_visbuffer_equal(ofb2, fb, "after a different clear depth")


class _FakeFb:
    """Just enough of orc.Framebuffer for assert_visbuffer_equal."""

    def __init__(self, scene, ids, depth):
        self.width, self.height = scene.width, scene.height
        self.data = np.stack([ids, depth])


def test_prepared_batch_of_122_draws_with_fused_cull(orc, rast_factory):
    """BASELINE config C4 through a prepared batch: 122 DrawMeshlets calls, frustum test fused into the mesh kernel
    (one lane per candidate meshlet), bit-exact vis-buffer and counters vs the oracle's bitmap-culled frame loop."""
    scene = scenes.instanced_scene()
    ofb, oc = oracle_render(orc, scene, cull=True)
    rast = rast_factory(fused_frustum_cull=True)
    gscene = rast.upload_scene(scene.meshlets)
    proj, view = scene.view_proj()
    batch = rast.create_batch(gscene, [dict(offset=nd.meshlet_offset, count=nd.meshlet_count, object_to_clip=scene.object_to_clip(nd),
                                            planes=rast.frustum_planes(proj, view, nd.model)) for nd in scene.nodes])
    fb = rast.create_framebuffer(scene.width, scene.height)
    for _ in range(2):
        fb.clear(0xFF000000, 0.0)
        rast.reset_counters()
        rast.draw_prepared(fb, batch)
        assert_visbuffer_equal(ofb, fb, "C4 prepared batch")
        c = rast.counters()
        assert [c["TrianglesProcessed"], c["TrianglesRasterized"], c["TrianglesClipped"]] == [int(oc[0]), int(oc[1]), int(oc[2])]
    stats = rast.draw_stats()
    assert stats["records"] > 0


def test_depth_tie_between_a_clipped_and_an_unclipped_triangle_of_one_packet(orc, rast_factory):
    """The reference draws a 16-packet's accepted lanes first and the pieces of its clipped lanes afterwards
    (Rasterizer.cpp:181-249), so at bit-equal depth the UNCLIPPED triangle wins even when the clipped one has the smaller
    primitive id. Two coplanar triangles (z = 0.5, w = 1 everywhere -> depth exactly 0.5): prim 0 leaves the guard band and is
    clipped, prim 1 lies inside it and overlaps prim 0 on screen."""
    w, h = 256, 256
    big = [(-0.5, -0.5, 0.5), (-0.5, 0.5, 0.5), (40.0, 0.0, 0.5)]            # x = 40 is far outside the 2896/256 guard band
    small = tri_px([(100, 100), (100, 160), (160, 100)], w, h, z=0.5)
    m = meshlet_from_clip_tris([big, small])
    rast = rast_factory(enable_binning=False, enable_clipping=True)
    ofb = orc.Framebuffer(w, h)
    ofb.clear(0xFFFFFFFF, 0.0)
    oc = orc.draw_meshlets(ofb, m, 0, 1, IDENT, binned=False, clipping=True)
    assert int(oc[2]) == 1                                                   # prim 0 really was clipped
    ids = ofb.get_pixels(0)
    assert ids[120, 120] == 1 and ids[128, 200] == 0                          # overlap -> prim 1 (drawn first); elsewhere prim 0
    gscene = rast.upload_scene(m)
    fb = rast.create_framebuffer(w, h)
    fb.clear(0xFFFFFFFF, 0.0)
    rast.draw_meshlets(fb, gscene, 0, 1, IDENT)
    assert_visbuffer_equal(ofb, fb, "clipped vs unclipped tie")
    # and two unclipped coplanar triangles still resolve to the smaller primitive id
    m2 = meshlet_from_clip_tris([tri_px([(90, 90), (90, 170), (170, 90)], w, h, z=0.5), small])
    ofb2 = orc.Framebuffer(w, h)
    ofb2.clear(0xFFFFFFFF, 0.0)
    orc.draw_meshlets(ofb2, m2, 0, 1, IDENT, binned=False, clipping=True)
    fb.clear(0xFFFFFFFF, 0.0)
    rast.draw_meshlets(fb, rast.upload_scene(m2), 0, 1, IDENT)
    assert_visbuffer_equal(ofb2, fb, "unclipped tie")
    assert fb.get_pixels(0)[120, 120] == 0


def test_get_pixels_on_a_side_stream_is_ordered_after_the_key_unpack(orc, rast_factory):
    """GetPixels of the depth layer on a caller's stream right after a draw: the key unpack the call itself enqueues on the
    device stream must finish before the de-tile kernel reads the layer (no event code on the caller's side), and a later
    clear of the framebuffer must wait for the copy."""
    import torch
    scene = scenes.grid_scene(60, 50, 1280, 720, seed=5)
    ofb, _ = oracle_render(orc, scene)
    rast = rast_factory()
    main, side = torch.cuda.Stream(), torch.cuda.Stream()
    rast.set_stream(main.cuda_stream)
    gfb, _, gscene = gpu_render(rast, scene)
    dst = torch.zeros((scene.height, scene.width), dtype=torch.int32, device="cuda")
    gfb.get_pixels_device(1, dst.data_ptr(), cuda_stream=side.cuda_stream)          # depth: needs k_keys_unpack first
    gfb.clear_layer(1, 0x7F800000)                                                  # overwrites the layer the copy reads
    torch.cuda.synchronize()
    assert np.array_equal(dst.cpu().numpy().view(np.uint32), ofb.get_pixels(1))
    assert int(gfb.download_tiled(1)[0]) == 0x7F800000


def test_scene_validation(rast_factory):
    rast = rast_factory()
    m = meshlet_from_clip_tris([tri_px([(1, 1), (1, 9), (9, 1)], 16, 16)], material_id=0)
    mats = np.zeros(1, dtype=MATERIAL_DTYPE)
    mats["AlphaCutoff"], mats["TextureId"] = 128, -1           # alpha-tested, no texture: a null dereference upstream
    with pytest.raises(api.SwrbError):
        rast.upload_scene(m, mats)
    mats["AlphaCutoff"] = 255
    s = rast.upload_scene(m, mats)
    with pytest.raises(api.SwrbError):                         # a batch cannot capture the transient device cull bitmap
        rast.create_batch(s, [dict(offset=0, count=1, object_to_clip=IDENT, use_device_bitmap=True)])
    assert s.meshlets_device_ptr() != 0
    s.touch()
