"""Bit-exact vis-buffer parity: CUDA path (through the C ABI) vs the CPU oracle. Needs a GPU."""
import numpy as np
import pytest

from glimpsw_b200 import scenes
from helpers import oracle_render, gpu_render, assert_visbuffer_equal

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("binning", [True, False], ids=["binned", "direct"])
def test_grid_small(orc, rast_factory, binning):
    scene = scenes.grid_scene(20, 16, 640, 360, seed=3, flip_fraction=0.2)
    ofb, oc = oracle_render(orc, scene)
    rast = rast_factory(enable_binning=binning)
    gfb, gc, _ = gpu_render(rast, scene)
    assert_visbuffer_equal(ofb, gfb, f"grid small binning={binning}")
    assert [gc["TrianglesProcessed"], gc["TrianglesRasterized"], gc["TrianglesClipped"]] == [int(oc[0]), int(oc[1]), int(oc[2])]


@pytest.mark.parametrize("binning", [True, False], ids=["binned", "direct"])
def test_config2_grid_1m_1080p(orc, rast_factory, binning):
    """BASELINE config C2: 999,600-triangle meshlet grid at 1920x1080, depth + triangle id, bit-exact."""
    scene = scenes.grid_scene()
    assert scene.num_triangles == 999600
    ofb, oc = oracle_render(orc, scene)
    rast = rast_factory(enable_binning=binning)
    gfb, gc, _ = gpu_render(rast, scene)
    assert_visbuffer_equal(ofb, gfb, f"C2 binning={binning}")
    assert [gc["TrianglesProcessed"], gc["TrianglesRasterized"], gc["TrianglesClipped"]] == [int(oc[0]), int(oc[1]), int(oc[2])]
    # GetPixels de-tiling agrees with the oracle's
    assert np.array_equal(gfb.get_pixels(0), ofb.get_pixels(0))


def test_config4_instanced_10m_culled_1080p(orc, rast_factory):
    """BASELINE config C4: ~10 M-triangle instanced scene (122 draws), frustum cull bitmap per node, 1920x1080."""
    scene = scenes.instanced_scene()
    assert scene.num_triangles > 9_900_000
    ofb, oc = oracle_render(orc, scene, cull=True)
    assert int(oc[0]) < 0.4 * scene.num_triangles            # heavy culling: >= 60 % of the triangles never reach setup
    for binning in (True, False):
        gfb, gc, _ = gpu_render(rast_factory(enable_binning=binning), scene, cull=True)
        assert_visbuffer_equal(ofb, gfb, f"C4 binning={binning}")
        assert [gc["TrianglesProcessed"], gc["TrianglesRasterized"], gc["TrianglesClipped"]] == [int(oc[0]), int(oc[1]), int(oc[2])]


def test_config3_textured_alpha_1440p(orc, rast_factory):
    """BASELINE config C3 stand-in: ~260 K textured triangles incl. a double-sided alpha-masked material, 2560x1440."""
    scene = scenes.torus_knot_scene(600, 216, 2560, 1440, tex_size=512, alpha_material=True)
    assert scene.num_triangles == 259200
    ofb, oc = oracle_render(orc, scene)
    gfb, gc, _ = gpu_render(rast_factory(), scene)
    assert_visbuffer_equal(ofb, gfb, "C3")
    assert gc["TrianglesRasterized"] == int(oc[1])


def test_max_render_size_2896(orc, rast_factory):
    """The reference's fixed-point limit (Rasterizer.h:203): 2896 x 2896."""
    scene = scenes.grid_scene(40, 40, 2896, 2896, seed=12)
    ofb, oc = oracle_render(orc, scene)
    gfb, gc, _ = gpu_render(rast_factory(), scene)
    assert_visbuffer_equal(ofb, gfb, "2896^2")
