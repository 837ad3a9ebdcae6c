"""Scissor rows (swrb_fb_set_scissor_rows) — the sort-first split of one view into horizontal bands.
Inside its band a scissored frame must equal the unscissored one bit for bit: depth and ids against the CPU oracle,
colour against the unscissored CUDA frame (and the oracle within 2/255). Needs a GPU."""
import numpy as np
import pytest

from glimpsw_b200 import api, scenes, sharding
from helpers import oracle_render, gpu_render, raster_mode

pytestmark = pytest.mark.gpu

MODES = {"binned": dict(enable_binning=True), "direct": dict(enable_binning=False),
         "direct_clip": dict(enable_binning=False, enable_clipping=True, enable_guardband=True)}


def _scene(name):
    if name == "room":          # big triangles: tile lists, super-tile lists, the clipper (the camera stands inside the box)
        return scenes.room_scene(640, 360, subdiv=3)
    if name == "knot":          # small textured triangles: inline raster, resolve with textures and lights
        return scenes.torus_knot_scene(120, 48, 960, 544, tex_size=128)
    if name == "alpha":         # alpha-tested materials: k_raster_alpha records
        return scenes.closeup_alpha_scene(640, 360)
    return scenes.patchwork_scene(20, 16, 640, 360)


def _rows(tiled, width, y0, y1):
    return tiled[y0 * width:y1 * width]          # rows [y0, y1) of a 4x4-tiled layer are one contiguous run (y0, y1 multiples of 4)


@pytest.mark.parametrize("mode", list(MODES))
@pytest.mark.parametrize("name", ["room", "knot", "alpha", "patchwork"])
def test_every_band_equals_the_full_frame(orc, rast_factory, name, mode):
    scene = _scene(name)
    rast = rast_factory(**MODES[mode])
    ofb, _ = oracle_render(orc, scene, **raster_mode(rast))
    n, w, h = scene.width * scene.height, scene.width, scene.height
    want_d, want_i = ofb.data[1, :n], ofb.data[0, :n]
    uni = api.Rasterizer.make_uniforms(**scenes.resolve_uniforms(scene, scene.nodes[0]))
    fb_full, c_full, gscene = gpu_render(rast, scene)
    full = np.zeros((h, w), dtype=np.uint32)
    if len(scene.textures):
        rast.resolve_prebuilt(fb_full, gscene, uni)
        full = fb_full.get_pixels(0)
    fb = rast.create_framebuffer(w, h)
    for world in (2, 3):
        image = np.full((h, w), 0xDEADBEEF, dtype=np.uint32)
        processed = 0
        for r in range(world):
            y0, y1 = sharding.band_rows(h, r, world, align=32)
            fb.set_scissor_rows(y0, y1)
            assert fb.scissor_rows() == (y0, y1)
            fb.clear(0xFF000000, 0.0)
            _, c, _ = gpu_render(rast, scene, gscene=gscene, fb=fb)
            processed += c["TrianglesProcessed"]
            assert c["TrianglesProcessed"] <= c_full["TrianglesProcessed"]
            assert np.array_equal(_rows(fb.download_tiled(1), w, y0, y1), _rows(want_d, w, y0, y1)), (name, mode, world, r, "depth")
            assert np.array_equal(_rows(fb.download_tiled(0), w, y0, y1), _rows(want_i, w, y0, y1)), (name, mode, world, r, "ids")
            if len(scene.textures):
                fb.clear(0xFF000000, 0.0)
                gpu_render(rast, scene, gscene=gscene, fb=fb)
                rast.resolve_prebuilt(fb, gscene, uni)
                band = _get_into(rast, fb, np.full((h, w), 0xDEADBEEF, dtype=np.uint32))
                assert np.array_equal(band[y0:y1], full[y0:y1]), (name, mode, world, r, "colour")
                assert (band[:y0] == 0xDEADBEEF).all() and (band[y1:] == 0xDEADBEEF).all()      # GetPixels moved the band only
                image[y0:y1] = band[y0:y1]
        if len(scene.textures):
            assert np.array_equal(image, full)
        assert processed >= c_full["TrianglesProcessed"]          # straddlers count in both bands
    fb.set_scissor_rows(0, 0)


def _get_into(rast, fb, template):
    """GetPixels into a pinned host image that already holds a pattern (only the scissor rows may change)."""
    host = rast.alloc_pinned(template.shape, np.uint32)
    host[...] = template
    fb.get_pixels_async(0, host)
    rast.sync()
    return np.array(host)


def test_band_cull_drops_meshlets_and_changing_the_scissor_between_frames(orc, rast_factory):
    """The mesh kernel's band test skips most of the scene for a narrow band; going band A -> band B -> no scissor with the
    frame loop's lazy state (seeds left by the resolve pass, prepared batches, swrb_frame_submit) stays exact."""
    scene = scenes.grid_scene(30, 30, 960, 544)
    rast = rast_factory(enable_binning=True)
    ofb, oc = oracle_render(orc, scene)
    n, w, h = scene.width * scene.height, scene.width, scene.height
    want_d, want_i = ofb.data[1, :n], ofb.data[0, :n]
    gscene = rast.upload_scene(scene.meshlets, scene.materials, scene.textures, scene.lights)
    node = scene.nodes[0]
    batch = rast.create_batch(gscene, [dict(offset=node.meshlet_offset, count=node.meshlet_count, object_to_clip=scene.object_to_clip(node))])
    fb = rast.create_framebuffer(w, h)
    uni = api.Rasterizer.make_uniforms(**scenes.resolve_uniforms(scene, node))
    frame = rast.make_frame(batch, uni, 0xFF000000, 0.0)
    for y0, y1 in [(0, 0), (256, 320), (256, 320), (0, 64), (480, 544), (0, 0), (64, 512)]:
        fb.set_scissor_rows(y0, y1)
        a, b = fb.scissor_rows()
        rast.reset_counters()
        rast.submit_frame(fb, frame)                     # Clear -> Draw -> Resolve; the depth layer is current afterwards
        rast.sync()
        c = rast.counters()
        assert np.array_equal(_rows(fb.download_tiled(1), w, a, b), _rows(want_d, w, a, b)), (y0, y1)
        if (a, b) == (0, h):
            assert c["TrianglesProcessed"] == int(oc[0])
        elif (a, b) == (480, 544):
            assert c["TrianglesProcessed"] < int(oc[0]) // 2, (y0, y1, c)     # the nearest 64 rows of an oblique view of a height field
        else:
            assert c["TrianglesProcessed"] < int(oc[0]), (y0, y1, c)
        # vis-buffer alone (no resolve): ids too
        fb.clear(0xFF000000, 0.0)
        rast.draw_prepared(fb, batch)
        assert np.array_equal(_rows(fb.download_tiled(0), w, a, b), _rows(want_i, w, a, b)), (y0, y1)
        assert np.array_equal(_rows(fb.download_tiled(1), w, a, b), _rows(want_d, w, a, b)), (y0, y1)


def test_bands_fill_one_device_image_and_arguments_are_checked(orc, rast_factory):
    import torch
    scene = scenes.torus_knot_scene(120, 48, 960, 544, tex_size=128)
    rast = rast_factory(enable_binning=True)
    fb_full, _, gscene = gpu_render(rast, scene)
    uni = api.Rasterizer.make_uniforms(**scenes.resolve_uniforms(scene, scene.nodes[0]))
    rast.resolve_prebuilt(fb_full, gscene, uni)
    full = fb_full.get_pixels(0)
    dst = torch.zeros((scene.height, scene.width), dtype=torch.int32, device="cuda")
    side = torch.cuda.Stream()
    fbs = [rast.create_framebuffer(scene.width, scene.height) for _ in range(4)]
    for r, fb in enumerate(fbs):                         # four "GPUs" (framebuffers), one band each, one destination image
        fb.set_scissor_rows(*sharding.band_rows(scene.height, r, 4))
        fb.clear(0xFF000000, 0.0)
        gpu_render(rast, scene, gscene=gscene, fb=fb)
        rast.resolve_prebuilt(fb, gscene, uni)
        if r % 2:
            fb.get_pixels_device(0, dst.data_ptr())
        else:
            fb.get_pixels_device(0, dst.data_ptr(), cuda_stream=side.cuda_stream)       # the on-stream variant moves the band as well
    rast.sync()
    torch.cuda.synchronize()
    assert np.array_equal(dst.cpu().numpy().view(np.uint32), full)
    fb = fbs[0]
    for bad in [(4, 64), (0, 60), (64, 64), (128, 64), (0, scene.height + 8)]:
        with pytest.raises(api.SwrbError):
            fb.set_scissor_rows(*bad)
    fb.set_scissor_rows(0, scene.height)
    assert fb.scissor_rows() == (0, scene.height)
