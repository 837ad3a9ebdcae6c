"""The alpha-tested fragment program FS_EncodeSurfaceId<true> (Shading.cpp:309-331) on the CUDA path vs the
oracle: the vis-buffer must stay bit-exact, including which fragments the texture's alpha channel rejects."""
import numpy as np
import pytest

from glimpsw_b200 import scenes
from helpers import oracle_render, gpu_render, assert_visbuffer_equal

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("binning", [True, False], ids=["binned", "direct"])
@pytest.mark.parametrize("tex_size,cam", [(256, (0.2, 1.0, 3.2)), (64, (0.1, 0.7, 2.2)), (512, (0.6, 3.0, 9.0))],
                         ids=["mixed", "magnified", "minified"])
def test_alpha_tested_material(orc, rast_factory, binning, tex_size, cam):
    scene = scenes.torus_knot_scene(120, 48, 960, 540, tex_size=tex_size, alpha_material=True)
    scene.camera.position[:] = cam
    ofb, oc = oracle_render(orc, scene)
    # the alpha test really rejects fragments: the opaque program covers more pixels
    opaque = orc.Framebuffer(scene.width, scene.height)
    opaque.clear(0xFF000000, 0.0)
    node = scene.nodes[0]
    orc.draw_meshlets(opaque, scene.meshlets, 0, len(scene.meshlets), scene.object_to_clip(node), materials=scene.materials)
    n = scene.width * scene.height
    assert (opaque.data[1, :n].view(np.float32) > 0).sum() > (ofb.data[1, :n].view(np.float32) > 0).sum()
    gfb, gc, gscene = gpu_render(rast_factory(enable_binning=binning), scene)
    assert_visbuffer_equal(ofb, gfb, f"alpha tex={tex_size}")
    assert [gc["TrianglesProcessed"], gc["TrianglesRasterized"], gc["TrianglesClipped"]] == [int(oc[0]), int(oc[1]), int(oc[2])]
    # and the resolve pass on top of it stays in tolerance
    uni = scenes.resolve_uniforms(scene, node)
    orc.resolve(ofb, scene.meshlets, scene.materials, scene.textures, scene.lights, **uni)
    rast_factory  # (same rasterizer object owns gfb)
    gfb.rast.resolve(gfb, gscene, **uni)
    a = gfb.download_tiled(0).view(np.uint8).astype(np.int32)
    b = ofb.data[0, :n].view(np.uint8).astype(np.int32)
    assert np.abs(a - b).max() <= 2
