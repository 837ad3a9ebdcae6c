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
