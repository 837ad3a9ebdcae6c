"""The threaded AVX-512 CPU baseline (oracle/baseline_mt.cpp) must equal the scalar oracle bit for bit."""
import numpy as np
import pytest

from glimpsw_b200 import scenes


@pytest.mark.parametrize("threads", [1, 3, 8])
def test_mt_baseline_equals_scalar_oracle(orc, threads):
    scene = scenes.grid_scene(30, 24, 960, 540, seed=5, flip_fraction=0.15)
    node = scene.nodes[0]
    m = scene.object_to_clip(node)
    fb = orc.Framebuffer(scene.width, scene.height)
    fb.clear(0xFF000000, 0.0)
    c = orc.draw_meshlets(fb, scene.meshlets, 0, len(scene.meshlets), m)
    base = orc.Baseline(threads)
    fb2 = orc.Framebuffer(scene.width, scene.height)
    base.clear(fb2, 0xFF000000, 0.0)
    c2 = base.draw_meshlets(fb2, scene.meshlets, 0, len(scene.meshlets), m)
    assert np.array_equal(fb.data, fb2.data)
    assert list(c[:3]) == list(c2[:3])
    base.close()


def test_mt_baseline_resolve_equals_scalar(orc):
    scene = scenes.torus_knot_scene(60, 24, 640, 360, tex_size=128, extra_lights=True)
    node = scene.nodes[0]
    m = scene.object_to_clip(node)
    uni = scenes.resolve_uniforms(scene, node)
    fbs = []
    base = orc.Baseline(4)
    for impl in ("scalar", "mt"):
        fb = orc.Framebuffer(scene.width, scene.height)
        fb.clear(0xFF000000, 0.0)
        if impl == "scalar":
            orc.draw_meshlets(fb, scene.meshlets, 0, len(scene.meshlets), m, materials=scene.materials)
            orc.resolve(fb, scene.meshlets, scene.materials, scene.textures, scene.lights, **uni)
        else:
            base.draw_meshlets(fb, scene.meshlets, 0, len(scene.meshlets), m, materials=scene.materials)
            base.resolve(fb, scene.meshlets, scene.materials, scene.textures, scene.lights, **uni)
        fbs.append(fb.data.copy())
    assert np.array_equal(fbs[0], fbs[1])
    base.close()
