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


@pytest.mark.parametrize("seed,spread", [(1, 0.6), (3, 40.0)])
def test_mt_baseline_equals_scalar_on_ragged_meshlets(orc, seed, spread):
    """Empty / partially filled meshlets, sub-pixel to guard-band-sized triangles, vertices behind the camera plane: the
    AVX-512 packet path (16-wide packets with masked tails) and the scalar spec agree bit for bit, counters included."""
    from test_kat_gpu import _ragged_meshlets
    meshlets = _ragged_meshlets(seed, 97, spread)
    m = np.zeros((4, 4), dtype=np.float32)
    m[0, 0], m[1, 1], m[2, 3], m[3, 2] = 1.0, 1.0, 1.0, 0.01
    fb = orc.Framebuffer(1000, 564)
    fb.clear(0xFF000000, 0.0)
    c = orc.draw_meshlets(fb, meshlets, 0, len(meshlets), m)
    base = orc.Baseline(3)
    fb2 = orc.Framebuffer(1000, 564)
    base.clear(fb2, 0xFF000000, 0.0)
    c2 = base.draw_meshlets(fb2, meshlets, 0, len(meshlets), m)
    n = 1000 * 564                       # (the layer stride is padded to 64 words; only the pixels are compared)
    assert np.array_equal(fb.data[:, :n], fb2.data[:, :n]) and list(c[:3]) == list(c2[:3])
    assert int(c[0]) == int(meshlets["NumTriangles"].astype(np.int64).sum()) and int(c[2]) > 0
    base.close()


def test_mt_baseline_resolve_equals_scalar_on_many_materials(orc):
    """scenes.patchwork_scene: nine materials (textures 32^2..512^2, with and without a normal / metal-rough layer) plus
    material-less meshlets, two lights; a fifth of the 4x4 fragments hold several materials. The 16-lane resolve with its
    scalar material waterfall must still equal the scalar spec bit for bit."""
    scene = scenes.patchwork_scene()
    node = scene.nodes[0]
    m, uni = scene.object_to_clip(node), scenes.resolve_uniforms(scene, node)
    n = scene.width * scene.height
    fb = orc.Framebuffer(scene.width, scene.height)
    fb.clear(0xFF000000, 0.0)
    orc.draw_meshlets(fb, scene.meshlets, 0, len(scene.meshlets), m, materials=scene.materials)
    ids, depth = fb.data[0, :n].copy(), fb.data[1, :n].view(np.float32)
    mats = np.where(depth > 0, scene.meshlets["MaterialId"][np.minimum(ids // 128, len(scene.meshlets) - 1)], 0xFFFFFFFE).reshape(-1, 16)
    assert np.mean([len(set(f.tolist()) - {0xFFFFFFFE}) >= 2 for f in mats[::11]]) > 0.1      # the scene does what it says
    orc.resolve(fb, scene.meshlets, scene.materials, scene.textures, scene.lights, **uni)
    base = orc.Baseline(3)
    fb2 = orc.Framebuffer(scene.width, scene.height)
    base.clear(fb2, 0xFF000000, 0.0)
    base.draw_meshlets(fb2, scene.meshlets, 0, len(scene.meshlets), m, materials=scene.materials)
    base.resolve(fb2, scene.meshlets, scene.materials, scene.textures, scene.lights, **uni)
    assert np.array_equal(fb.data[:, :n], fb2.data[:, :n])
    base.close()
