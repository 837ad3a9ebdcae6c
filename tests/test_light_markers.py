"""Tail of ShadingContext::Resolve (Shading.cpp:690-731): point / spot lights inside the frustum are drawn as soft discs
over the resolved image where the light is not occluded. CPU: hand-derived properties of the oracle's pass. GPU: the
CUDA resolve (k_resolve + k_light_marker) against the oracle within the resolve tolerance."""
import numpy as np
import pytest

from glimpsw_b200 import scenes, camera as cam
from glimpsw_b200.layout import detile
from helpers import oracle_render, gpu_render, assert_visbuffer_equal

BG = 0xFF102030


def marker_scene():
    scene = scenes.torus_knot_scene(120, 48, 1280, 720, tex_size=64)
    scene.lights = np.concatenate([scenes.default_light(),
                                   scenes.make_light(1, position=(0.2, 0.8, 2.0), color=(1.0, 0.5, 0.2), intensity=3000.0, radius=10.0),    # in front of everything
                                   scenes.make_light(1, position=(1.0, 0.6, 1.5), color=(0.2, 0.9, 1.0), intensity=3000.0, radius=10.0),    # partly behind the knot
                                   scenes.make_light(2, position=(-1.0, 0.6, 1.5), direction=(0.5, -0.5, -0.7), color=(0.9, 0.9, 0.1),
                                                     intensity=5000.0, radius=15.0, inner=0.2, outer=0.5),
                                   scenes.make_light(1, position=(0.0, 0.0, 30.0), color=(1.0, 1.0, 1.0), intensity=1.0, radius=1.0)])     # behind the camera: skipped
    return scene


def test_marker_disc_geometry_and_occlusion(orc):
    scene = marker_scene()
    uni = scenes.resolve_uniforms(scene, scene.nodes[0])
    w2c = np.asarray(uni["world_to_clip"], dtype=np.float64)          # [c, r]
    # empty depth buffer: every visible light shows its whole disc
    fb = orc.Framebuffer(scene.width, scene.height)
    fb.clear(BG, 0.0)
    orc.draw_light_markers(fb, scene.lights[1:2], uni["world_to_clip"])
    img = detile(fb.data[0], scene.width, scene.height)
    changed = img != BG
    pos = np.array([0.2, 0.8, 2.0, 1.0])
    clip = pos @ w2c
    cx, cy = (clip[0] / clip[3] * 0.5 + 0.5) * scene.width, (clip[1] / clip[3] * 0.5 + 0.5) * scene.height
    radius = max(scene.width, scene.height) / 30.0 / clip[3]
    ys, xs = np.nonzero(changed)
    assert abs(xs.mean() + 0.5 - cx) < 0.6 and abs(ys.mean() + 0.5 - cy) < 0.6
    assert abs(changed.sum() - np.pi * radius * radius) < 0.05 * np.pi * radius * radius
    d2 = (xs + 0.5 - cx) ** 2 + (ys + 0.5 - cy) ** 2
    assert d2.max() < radius * radius
    # centre pixel: alpha = 1 - a^2 with a ~ 0 -> the light colour; alpha channel stays 255
    c = int(img[int(cy), int(cx)])
    assert abs((c & 255) - 255) <= 2 and abs(((c >> 8) & 255) - 128) <= 3 and abs(((c >> 16) & 255) - 51) <= 3 and (c >> 24) == 255
    # directional lights and lights outside the frustum draw nothing
    fb.clear(BG, 0.0)
    orc.draw_light_markers(fb, scene.lights[[0, 4]], uni["world_to_clip"])
    assert (fb.data[0, :scene.width * scene.height] == BG).all()
    # occlusion: with the scene's depth, the second light loses part of its disc, and only where geometry is nearer
    ofb, _ = oracle_render(orc, scene)
    depth = detile(ofb.data[1], scene.width, scene.height).view(np.float32)
    ofb.data[0, :] = BG
    orc.draw_light_markers(ofb, scene.lights[2:3], uni["world_to_clip"])
    occluded = detile(ofb.data[0], scene.width, scene.height) != BG
    fb.clear(BG, 0.0)
    orc.draw_light_markers(fb, scene.lights[2:3], uni["world_to_clip"])
    full = detile(fb.data[0], scene.width, scene.height) != BG
    assert 0.1 * full.sum() < occluded.sum() < 0.9 * full.sum()
    clip2 = np.array([1.0, 0.6, 1.5, 1.0]) @ w2c
    light_depth = clip2[2] / clip2[3]
    assert (depth[full & ~occluded] >= light_depth - 1e-6).all() and (depth[occluded] < light_depth + 1e-6).all()


@pytest.mark.gpu
@pytest.mark.parametrize("binning", [True, False], ids=["binned", "direct"])
def test_gpu_resolve_with_light_markers(orc, rast_factory, binning):
    from test_resolve_gpu import color_error, MAX_ABS, MIN_PSNR
    scene = marker_scene()
    rast = rast_factory(enable_binning=binning)
    ofb, _ = oracle_render(orc, scene)
    gfb, _, gscene = gpu_render(rast, scene)
    assert_visbuffer_equal(ofb, gfb, scene.name)
    uni = scenes.resolve_uniforms(scene, scene.nodes[0], 0.8)
    n = scene.width * scene.height
    plain = orc.Framebuffer(scene.width, scene.height)
    plain.data[:] = ofb.data
    orc.resolve(plain, scene.meshlets, scene.materials, scene.textures, scene.lights, **{**uni, "world_to_clip": None})
    orc.resolve(ofb, scene.meshlets, scene.materials, scene.textures, scene.lights, **uni)
    marked = int((plain.data[0, :n] != ofb.data[0, :n]).sum())
    assert marked > 3000                                    # the three visible lights really are in the image
    rast.resolve(gfb, gscene, **uni)
    max_abs, psnr, frac = color_error(ofb.data[0, :n], gfb.download_tiled(0))
    assert max_abs <= MAX_ABS and psnr >= MIN_PSNR, f"max abs {max_abs}/255, PSNR {psnr:.1f} dB, {frac:.4%} pixels differ"
    # resolving from materialised layers (depth / id read back first) gives the same picture
    gfb2, _, _ = gpu_render(rast, scene, gscene=gscene)
    gfb2.download_tiled(1)
    rast.resolve(gfb2, gscene, **uni)
    assert np.array_equal(gfb2.download_tiled(0), gfb.download_tiled(0))
