"""The BASELINE configs exactly as bench.py runs them (glimpsw_b200/workloads.py): the CUDA path, through prepared batches and
swrb_frame_submit, against the committed oracle fixtures (tests/golden/bench_configs.json + colour_*.npz, written by
tests/golden/make_golden_bench.py on the CPU oracle). Vis-buffer and counters exact, resolved colour <= 2/255. Needs a GPU.

The `-m "not gpu"` half pins the fixtures themselves: the workload generators still produce the meshlets the fixtures were
made from, and the CPU oracle still produces the committed hashes for the small configs."""
import hashlib
import json
import os

import numpy as np
import pytest

from glimpsw_b200 import api, workloads

GOLDEN = json.load(open(workloads.GOLDEN))
HERE = os.path.dirname(os.path.abspath(__file__))


def _colour_fixture(name):
    z = np.load(os.path.join(HERE, "golden", name))
    return z["rgb"], int(z["stride"])


def _check(rast, wl, gscene, fb, want, view=None):
    batch = rast.create_batch(gscene, workloads.view_draws(rast, wl, view))
    fb.clear(0xFF000000, 0.0)
    rast.reset_counters()
    rast.draw_prepared(fb, batch)
    depth, ids = fb.download_tiled(1), fb.download_tiled(0)
    c = rast.counters()
    assert [c["TrianglesProcessed"], c["TrianglesRasterized"], c["TrianglesClipped"]] == want["counters"]
    assert int((depth.view(np.float32) > 0).sum()) == want["covered_pixels"]
    assert hashlib.sha256(depth.tobytes()).hexdigest() == want["depth_sha256"], "depth layer differs from the oracle fixture"
    assert hashlib.sha256(ids.tobytes()).hexdigest() == want["id_sha256"], "surface-id layer differs from the oracle fixture"
    if wl.resolve:
        uni = api.Rasterizer.make_uniforms(**workloads.view_uniforms(wl, view))
        frame = rast.make_frame(batch, uni)
        for _ in range(2):                                   # the second frame starts from the seeds the first resolve left
            rast.submit_frame(fb, frame)
        img = fb.get_pixels(0).view(np.uint8).reshape(fb.height, fb.width, 4)[..., :3]
        if "colour" in want:
            ref, stride = _colour_fixture(want["colour"])
            err = np.abs(img[::stride, ::stride].astype(np.int32) - ref.astype(np.int32))
            assert int(err.max()) <= 2, f"resolved colour off by {int(err.max())}/255"
            mse = float((err.astype(np.float64) ** 2).mean())
            assert mse == 0 or 10 * np.log10(255.0 ** 2 / mse) >= 50.0
        assert np.array_equal(fb.download_tiled(1), depth)   # the resolve pass stored the depth layer it read from the keys
    batch.destroy()


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["c1_knot", "c1_sponza", "c2_grid", "c3_knot"])
@pytest.mark.parametrize("binning", [True, False], ids=["binned", "direct"])
def test_single_view_configs(rast_factory, name, binning):
    wl = workloads.build(name)
    if not binning and name == "c1_sponza":
        pytest.skip("the fixture is the binned frame; the unbinned + clipped Sponza frame is checked against the live oracle below")
    rast = rast_factory(enable_binning=binning, enable_clipping=False)
    scene = wl.scene
    gscene = rast.upload_scene(scene.meshlets, scene.materials if len(scene.materials) else None, scene.textures, scene.lights)
    fb = rast.create_framebuffer(scene.width, scene.height)
    _check(rast, wl, gscene, fb, GOLDEN[name])


@pytest.mark.gpu
def test_sponza_lowpoly_unbinned_with_clipping(orc, rast_factory):
    """The mode RasterBench.cpp:77 selects (EnableBinning = false) on the reference's own asset: 87 triangles cross the
    camera plane or leave the guard band and go through the clipper; vis-buffer and counters equal the live oracle."""
    from helpers import oracle_render, gpu_render, assert_visbuffer_equal
    wl = workloads.build("c1_sponza")
    ofb, oc = oracle_render(orc, wl.scene, binned=False, clipping=True)
    rast = rast_factory(enable_binning=False, enable_clipping=True)
    gfb, gc, _ = gpu_render(rast, wl.scene)
    assert_visbuffer_equal(ofb, gfb, "Sponza_LowPoly unbinned + clipping")
    assert [gc["TrianglesProcessed"], gc["TrianglesRasterized"], gc["TrianglesClipped"]] == [int(oc[0]), int(oc[1]), int(oc[2])]
    assert int(oc[2]) > 0


@pytest.mark.gpu
def test_view_batches(rast_factory):
    """c4_views (the bench headline) at 1080p — three of the 64 views — and one C5 view at 2048^2, fused frustum cull."""
    rast = rast_factory(fused_frustum_cull=True)
    for name, views in (("c4_views", [0, 21, 63]), ("c5_views", [0])):
        wl = workloads.build(name)
        scene = wl.scene
        gscene = rast.upload_scene(scene.meshlets, scene.materials, scene.textures, scene.lights)
        fb = rast.create_framebuffer(scene.width, scene.height)
        for v in views:
            _check(rast, wl, gscene, fb, GOLDEN[name]["views"][str(v)], v)
        fb.destroy()
        gscene.destroy()


@pytest.mark.gpu
def test_bench_frame_loop_several_contexts_in_flight(rast_factory):
    """What bench.py's timed region does — three render contexts on their own streams taking turns over the views of the
    batch, GetPixels on a shared side stream into a ring of device buffers, no host synchronisation in between — produces
    the same composites as rendering each view alone."""
    import torch
    wl = workloads.build("c4_views")
    scene = wl.scene
    views = [0, 21, 42, 63, 5, 11]
    ctxs = []
    for _ in range(3):
        r = rast_factory(fused_frustum_cull=True)
        st = torch.cuda.Stream()
        r.set_stream(st.cuda_stream)
        r.set_mesh_occupancy(2)
        ctxs.append((r, st, r.upload_scene(scene.meshlets, scene.materials, scene.textures, scene.lights), r.create_framebuffer(scene.width, scene.height)))
    frames = []
    for i, v in enumerate(views):
        r, st, gs, fb = ctxs[i % 3]
        frames.append(r.make_frame(r.create_batch(gs, workloads.view_draws(r, wl, v)), api.Rasterizer.make_uniforms(**workloads.view_uniforms(wl, v))))
    # reference composites: one view at a time, synchronised
    want = []
    for i, v in enumerate(views):
        r, st, gs, fb = ctxs[i % 3]
        r.submit_frame(fb, frames[i])
        want.append(fb.get_pixels(0).copy())
    comm = torch.cuda.Stream()
    ring = [torch.zeros((scene.height, scene.width), dtype=torch.int32, device="cuda") for _ in range(len(views))]
    for rep in range(3):
        for i, v in enumerate(views):
            r, st, gs, fb = ctxs[i % 3]
            r.submit_frame(fb, frames[i])
            fb.get_pixels_device(0, ring[i].data_ptr(), cuda_stream=comm.cuda_stream)
    torch.cuda.synchronize()
    for i in range(len(views)):
        assert np.array_equal(ring[i].cpu().numpy().view(np.uint32), want[i]), f"view {views[i]}"
    c0 = GOLDEN["c4_views"]["views"]["0"]
    ref, stride = _colour_fixture(c0["colour"])
    img = want[0].view(np.uint8).reshape(scene.height, scene.width, 4)[::stride, ::stride, :3]
    assert int(np.abs(img.astype(np.int32) - ref.astype(np.int32)).max()) <= 2


def test_fixtures_match_the_generators():
    """CPU: the scenes the fixtures were generated from are the scenes the workloads build today."""
    for name in ("c1_knot", "c1_sponza", "c2_grid"):
        wl = workloads.build(name)
        assert hashlib.sha256(np.ascontiguousarray(wl.scene.meshlets).tobytes()).hexdigest() == GOLDEN[name]["meshlets_sha256"], name
        assert wl.scene.num_triangles == GOLDEN[name]["triangles"]
    assert GOLDEN["c1_sponza"]["triangles"] == 63084 and GOLDEN["c4_views"]["triangles"] == 9994240
    assert len(GOLDEN["c4_views"]["views"]) == workloads.NUM_VIEWS


def test_oracle_reproduces_the_small_fixtures(orc):
    """CPU: oracle.cpp still produces the committed hashes (C1 knot, Sponza_LowPoly, C2)."""
    import sys
    sys.path.insert(0, os.path.join(HERE, "golden"))
    import make_golden_bench
    for name in ("c1_sponza", "c2_grid"):
        wl = workloads.build(name)
        frame, colour = make_golden_bench.oracle_frame(orc, wl)
        for k in ("depth_sha256", "id_sha256", "counters", "covered_pixels"):
            assert frame[k] == GOLDEN[name][k], (name, k)
        if colour is not None:
            ref, stride = _colour_fixture(GOLDEN[name]["colour"])
            assert np.array_equal(colour, ref)
