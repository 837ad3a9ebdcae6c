"""HiZ occlusion culling: texutil::DownsampleDepth + the HiZ half of CullMeshlets."""
import numpy as np
import pytest

from glimpsw_b200 import scenes, textures as tx
from helpers import oracle_render, gpu_render


def make_scene():
    return scenes.instanced_scene(subdivisions=4, instances=64, width=1920, height=1080)


def oracle_pyramid(orc, scene):
    fb, _ = oracle_render(orc, scene)
    hw, hh = orc.hiz_dims(scene.width, scene.height)
    pyr = tx.create_texture(hw, hh, 16, 1)
    orc.downsample_depth(fb, pyr)
    return fb, pyr


def test_oracle_pyramid_is_a_min_pyramid(orc):
    scene = make_scene()
    fb, pyr = oracle_pyramid(orc, scene)
    assert (pyr.width, pyr.height, pyr.mip_levels) == (1024, 1024, 9)            # Main.cpp:54-56 for 1920x1080
    d = fb.get_pixels(1).view(np.float32)
    for m in range(0, 6):
        t = 1 << (m + 1)
        h, w = scene.height // t * t, scene.width // t * t                        # texels fully inside the frame
        ref = d[:h, :w].reshape(h // t, t, w // t, t).min(axis=(1, 3))
        lvl = tx.get_pixels(pyr, 0, m).view(np.float32)
        assert np.array_equal(lvl[: h // t, : w // t], ref), f"level {m}"
    # partially covered texels only see in-frame pixels: level 3 (16 px) row 67 covers rows 1072..1079 of 1080
    lvl3 = tx.get_pixels(pyr, 0, 3).view(np.float32)
    assert np.array_equal(lvl3[67, :120], d[1072:1080, :1920].reshape(8, 120, 16).min(axis=(0, 2)))


def test_oracle_hiz_only_removes_meshlets_and_keeps_the_image(orc):
    """Culling against the depth pyramid of the SAME view is conservative: the frame must not change."""
    scene = make_scene()
    fb, pyr = oracle_pyramid(orc, scene)
    proj, view = scene.view_proj()
    fb2 = orc.Framebuffer(scene.width, scene.height)
    fb2.clear(0xFF000000, 0.0)
    removed = 0
    for node in scene.nodes:
        ms = scene.meshlets[node.meshlet_offset:node.meshlet_offset + node.meshlet_count]
        b0, n0 = orc.cull_meshlets_hiz(ms, proj, view, node.model, view, scene.width, scene.height, None)
        b1, n1 = orc.cull_meshlets_hiz(ms, proj, view, node.model, view, scene.width, scene.height, pyr)
        assert np.all((b1 & ~b0) == 0) and n1 <= n0
        removed += n0 - n1
        orc.draw_meshlets(fb2, scene.meshlets, node.meshlet_offset, node.meshlet_count, scene.object_to_clip(node), cull_bitmap=b1)
    assert removed > 50
    n = scene.width * scene.height
    assert np.array_equal(fb.data[:, :n], fb2.data[:, :n])


@pytest.mark.gpu
def test_gpu_pyramid_and_hiz_cull_match_oracle(orc, rast_factory):
    scene = make_scene()
    ofb, pyr = oracle_pyramid(orc, scene)
    rast = rast_factory()
    gfb, _, gscene = gpu_render(rast, scene)
    hiz = rast.create_hiz(scene.width, scene.height)
    assert (hiz.width, hiz.height, hiz.mip_levels, hiz.row_shift, hiz.layer_stride) == (pyr.width, pyr.height, pyr.mip_levels, pyr.row_shift, pyr.layer_stride)
    hiz.build(gfb)
    got = hiz.download()
    want = pyr.data[: pyr.layer_stride].view(np.float32)
    # compare every texel the reference's recursion writes (whole 8x8 blocks with their origin inside the frame)
    for m in range(pyr.mip_levels):
        t = 1 << (m + 1)
        blk = (4 if m == pyr.mip_levels - 1 else 8) * t
        ty, txx = -(-scene.height // blk) * (blk // t), -(-scene.width // blk) * (blk // t)
        y, x = np.meshgrid(np.arange(ty), np.arange(txx), indexing="ij")
        off = int(pyr.mip_offsets[m]) + tx.texel_offset(x, y, pyr.row_shift - m)
        assert np.array_equal(got[off], want[off]), f"pyramid level {m}"
    proj, view = scene.view_proj()
    total = 0
    for node in scene.nodes:
        ms = scene.meshlets[node.meshlet_offset:node.meshlet_offset + node.meshlet_count]
        ob, on = orc.cull_meshlets_hiz(ms, proj, view, node.model, view, scene.width, scene.height, pyr)
        gb, gn = rast.cull_meshlets_hiz(gscene, node.meshlet_offset, node.meshlet_count, proj, view, node.model, view,
                                        scene.width, scene.height, hiz)
        assert gn == on and np.array_equal(gb, ob)
        fb_, fn = rast.cull_meshlets_hiz(gscene, node.meshlet_offset, node.meshlet_count, proj, view, node.model, view,
                                         scene.width, scene.height, None)
        assert fn >= gn
        total += fn - gn
    assert total > 50
