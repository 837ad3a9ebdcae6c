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
