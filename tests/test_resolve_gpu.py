"""Resolve-pass parity: CUDA k_resolve vs the CPU oracle on the same vis-buffer. Needs a GPU.

Gate (BASELINE north_star): max abs error <= 2/255 per 8-bit channel and PSNR >= 50 dB."""
import numpy as np
import pytest

from glimpsw_b200 import scenes
from helpers import oracle_render, gpu_render, assert_visbuffer_equal

pytestmark = pytest.mark.gpu

MAX_ABS = 2          # of 255
MIN_PSNR = 50.0      # dB


def color_error(a_u32, b_u32):
    a = a_u32.view(np.uint8).reshape(-1, 4).astype(np.int32)
    b = b_u32.view(np.uint8).reshape(-1, 4).astype(np.int32)
    diff = np.abs(a - b)
    mse = float((diff[:, :3].astype(np.float64) ** 2).mean())
    psnr = 99.0 if mse == 0 else 10 * np.log10(255.0 ** 2 / mse)
    return int(diff.max()), psnr, float((diff.max(axis=1) > 0).mean())


def run_scene(orc, rast, scene, exposure=1.0):
    ofb, _ = oracle_render(orc, scene)
    gfb, _, gscene = gpu_render(rast, scene)
    assert_visbuffer_equal(ofb, gfb, scene.name)
    node = scene.nodes[0]
    uni = scenes.resolve_uniforms(scene, node, exposure)
    orc.resolve(ofb, scene.meshlets, scene.materials, scene.textures, scene.lights, **uni)
    rast.resolve(gfb, gscene, **uni)
    n = scene.width * scene.height
    return color_error(ofb.data[0, :n], gfb.download_tiled(0)), ofb, gfb


@pytest.mark.parametrize("binning", [True, False], ids=["binned", "direct"])
def test_config1_textured_knot_1080p(orc, rast_factory, binning):
    """BASELINE config C1 stand-in: ~72K-triangle textured mesh, 1920x1080, vis-buffer + resolve."""
    scene = scenes.torus_knot_scene()
    (max_abs, psnr, frac), ofb, gfb = run_scene(orc, rast_factory(enable_binning=binning), scene)
    assert max_abs <= MAX_ABS and psnr >= MIN_PSNR, f"max abs {max_abs}/255, PSNR {psnr:.1f} dB, {frac:.4%} pixels differ"
    # sky pixels resolve to opaque black, and something was actually shaded
    px = gfb.get_pixels(0)
    assert (px == 0xFF000000).any() and (px != 0xFF000000).mean() > 0.05


def test_resolve_point_and_spot_lights(orc, rast_factory):
    scene = scenes.torus_knot_scene(120, 48, 1280, 720, tex_size=256, extra_lights=True)
    (max_abs, psnr, frac), _, _ = run_scene(orc, rast_factory(), scene, exposure=0.7)
    assert max_abs <= MAX_ABS and psnr >= MIN_PSNR, f"max abs {max_abs}/255, PSNR {psnr:.1f} dB, {frac:.4%} pixels differ"


def test_resolve_many_materials_per_fragment(orc, rast_factory):
    """scenes.patchwork_scene: a fifth of the 4x4 fragments hold two or more of nine materials with different texture
    sizes / layer counts (plus material-less meshlets): the half-warp material waterfall and its per-fragment filter votes."""
    scene = scenes.patchwork_scene()
    (max_abs, psnr, frac), _, _ = run_scene(orc, rast_factory(), scene)
    assert max_abs <= MAX_ABS and psnr >= MIN_PSNR, f"max abs {max_abs}/255, PSNR {psnr:.1f} dB, {frac:.4%} pixels differ"
    # and through the clip-cached path (no read-back between draw and resolve)
    cached, _ = _frame(rast_factory(), scene)
    ofb, _ = oracle_render(orc, scene)
    orc.resolve(ofb, scene.meshlets, scene.materials, scene.textures, scene.lights, **scenes.resolve_uniforms(scene, scene.nodes[0]))
    max_abs, psnr, frac = color_error(ofb.data[0, :scene.width * scene.height], cached.download_tiled(0))
    assert max_abs <= MAX_ABS and psnr >= MIN_PSNR


def test_resolve_untextured_grid(orc, rast_factory):
    """Material-less meshlets (MaterialId == UINT_MAX) shade with albedo 0 -> tonemapped black, still in tolerance."""
    scene = scenes.grid_scene(24, 20, 640, 480, seed=9)
    scene.lights = scenes.default_light()
    (max_abs, psnr, _), _, _ = run_scene(orc, rast_factory(), scene)
    assert max_abs <= MAX_ABS and psnr >= MIN_PSNR


def test_resolve_minified_and_magnified(orc, rast_factory):
    """Tiny textures force mip levels > 0 (nearest path) while a close camera forces bilinear magnification."""
    far = scenes.torus_knot_scene(90, 36, 1024, 576, tex_size=512)
    far.camera.position[:] = (0.6, 3.0, 9.0)
    (max_abs, psnr, _), _, _ = run_scene(orc, rast_factory(), far)
    assert max_abs <= MAX_ABS and psnr >= MIN_PSNR
    near = scenes.torus_knot_scene(90, 36, 1024, 576, tex_size=64)
    near.camera.position[:] = (0.2, 0.9, 2.6)
    (max_abs, psnr, _), _, _ = run_scene(orc, rast_factory(), near)
    assert max_abs <= MAX_ABS and psnr >= MIN_PSNR


# ---- the resolve pass's per-vertex clip cache (k_resolve<.., kClipCached>) ------------------------------------------
def _frame(rast, scene, gscene=None, per_node_calls=False, resolve_node=0, fb=None):
    """clear -> draw -> resolve with NO read-back in between (a read-back unpacks the keys and takes the uncached path)."""
    if gscene is None:
        gscene = rast.upload_scene(scene.meshlets, scene.materials, scene.textures, scene.lights)
    if fb is None:
        fb = rast.create_framebuffer(scene.width, scene.height)
    fb.clear(0xFF000000, 0.0)
    draws = [dict(offset=n.meshlet_offset, count=n.meshlet_count, object_to_clip=scene.object_to_clip(n)) for n in scene.nodes]
    if per_node_calls:
        for d in draws:
            rast.draw_meshlets(fb, gscene, d["offset"], d["count"], d["object_to_clip"])
    else:
        rast.draw_batch(fb, gscene, draws)
    rast.resolve(fb, gscene, **scenes.resolve_uniforms(scene, scene.nodes[resolve_node]))
    return fb, gscene


@pytest.mark.parametrize("binning", [True, False], ids=["binned", "direct"])
def test_resolve_clip_cache_is_bit_identical_and_in_tolerance(orc, rast_factory, binning):
    scene = scenes.torus_knot_scene(120, 48, 1280, 720, tex_size=256, extra_lights=True)
    cached, _ = _frame(rast_factory(enable_binning=binning), scene)
    plain, _ = _frame(rast_factory(enable_binning=binning, resolve_cache=False), scene)
    a, b = cached.download_tiled(0), plain.download_tiled(0)
    assert np.array_equal(a, b), f"{int((a != b).sum())} pixels differ between the cached and the re-transforming resolve"
    ofb, _ = oracle_render(orc, scene)
    orc.resolve(ofb, scene.meshlets, scene.materials, scene.textures, scene.lights, **scenes.resolve_uniforms(scene, scene.nodes[0]))
    max_abs, psnr, frac = color_error(ofb.data[0, :scene.width * scene.height], a)
    assert max_abs <= MAX_ABS and psnr >= MIN_PSNR, f"max abs {max_abs}/255, PSNR {psnr:.1f} dB, {frac:.4%} pixels differ"
    # the depth layer is still intact after a cached resolve (it is unpacked from the keys on demand)
    assert np.array_equal(cached.download_tiled(1), ofb.data[1, :scene.width * scene.height])


def test_resolve_clip_cache_falls_back_when_matrices_differ(rast_factory):
    """ShadingContext::Resolve transforms every pixel with the context's CURRENT ObjectToClipMat (Shading.cpp:509-511),
    also pixels of nodes drawn with another matrix; the cache must not change that."""
    from glimpsw_b200 import textures as tx
    from glimpsw_b200.layout import MATERIAL_DTYPE
    scene = scenes.instanced_scene(subdivisions=3, instances=27, width=640, height=360)
    scene.meshlets["MaterialId"] = 0
    scene.materials = np.zeros(1, dtype=MATERIAL_DTYPE)
    scene.materials["AlphaCutoff"] = 255
    scene.textures = [tx.procedural_material_texture(128, seed=3)]
    scene.lights = scenes.default_light()
    for per_node in (False, True):
        cached, _ = _frame(rast_factory(), scene, per_node_calls=per_node, resolve_node=1)
        plain, _ = _frame(rast_factory(resolve_cache=False), scene, per_node_calls=per_node, resolve_node=1)
        a, b = cached.download_tiled(0), plain.download_tiled(0)
        assert (a != 0xFF000000).mean() > 0.02, "nothing was shaded"
        assert np.array_equal(a, b)


def test_resolve_after_meshlet_update_uses_new_positions(rast_factory):
    scene = scenes.torus_knot_scene(60, 24, 640, 360, tex_size=128)
    moved = scene.meshlets.copy()
    moved["Positions"][:, 1, :] += 0.05
    imgs = []
    for cache in (True, False):
        rast = rast_factory(resolve_cache=cache)
        gscene = rast.upload_scene(scene.meshlets, scene.materials, scene.textures, scene.lights)
        fb = rast.create_framebuffer(scene.width, scene.height)
        fb.clear(0xFF000000, 0.0)
        n = scene.nodes[0]
        rast.draw_meshlets(fb, gscene, n.meshlet_offset, n.meshlet_count, scene.object_to_clip(n))
        gscene.update_meshlets(moved, 0)           # between draw and resolve: the reference would shade the moved vertices
        rast.resolve(fb, gscene, **scenes.resolve_uniforms(scene, n))
        imgs.append(fb.download_tiled(0))
    assert np.array_equal(imgs[0], imgs[1])
