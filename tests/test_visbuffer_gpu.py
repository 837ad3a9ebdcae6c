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


@pytest.mark.parametrize("blocks", [1, 2, 3])
def test_mesh_occupancy_does_not_change_results(orc, rast_factory, blocks):
    """swrb_device_set_mesh_occupancy only resizes the mesh kernel's persistent grid; depth / ids / counters stay bit-exact."""
    scene = scenes.grid_scene(60, 50, 1280, 720, seed=5, flip_fraction=0.1)
    ofb, oc = oracle_render(orc, scene)
    rast = rast_factory()
    rast.set_mesh_occupancy(blocks)
    gfb, gc, _ = gpu_render(rast, scene)
    assert_visbuffer_equal(ofb, gfb, f"mesh occupancy {blocks}")
    assert [gc["TrianglesProcessed"], gc["TrianglesRasterized"]] == [int(oc[0]), int(oc[1])]
    from glimpsw_b200 import api
    with pytest.raises(api.SwrbError):
        rast.set_mesh_occupancy(5)


def test_send_pixels_with_flow_control_on_one_gpu(orc, rast_factory):
    """swrb_fb_send_pixels (the multi-GPU composite transfer kernel) with destination and both flags in local memory:
    it must wait for the ack value, de-tile exactly like GetPixels, and raise the ready flag after its last store."""
    import torch
    scene = scenes.grid_scene(20, 16, 640, 360, seed=3)
    rast = rast_factory()
    stream, side = torch.cuda.Stream(), torch.cuda.Stream()
    rast.set_stream(stream.cuda_stream)
    gfb, _, _ = gpu_render(rast, scene)
    want = gfb.get_pixels(0)
    dst = torch.zeros((scene.height, scene.width), dtype=torch.int32, device="cuda")
    flags = torch.zeros(2, dtype=torch.int64, device="cuda")            # [0] = ack (waited on), [1] = ready (raised)
    torch.cuda.synchronize()
    gfb.send_pixels(0, dst.data_ptr(), side.cuda_stream, wait_flag=flags.data_ptr(), wait_value=7,
                    signal_flag=flags.data_ptr() + 8, signal_value=3)
    import time
    time.sleep(0.2)
    assert not side.query(), "the kernel must still be waiting for the ack flag"
    with torch.cuda.stream(stream):
        flags[0:1].fill_(7)                                               # the consumer releases the slot
    side.synchronize()
    assert int(flags[1].item()) == 3
    assert np.array_equal(dst.cpu().numpy().view(np.uint32), want)
    # the consumer-side kernel: waits for ready >= expected, then writes the ack values
    acks = torch.zeros(3, dtype=torch.int64, device="cuda")
    ready = torch.full((3,), 5, dtype=torch.int64, device="cuda")
    rast.peer_collect(side.cuda_stream, ready.data_ptr(), 3, 5, [acks.data_ptr() + 8 * i for i in range(3)], 9)
    side.synchronize()
    assert acks.tolist() == [9, 9, 9]
