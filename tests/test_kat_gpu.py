"""Adversarial known-answer cases through the CUDA path (both raster modes) vs the oracle. Needs a GPU."""
import numpy as np
import pytest

from glimpsw_b200 import scenes, camera as cam
from glimpsw_b200.layout import MATERIAL_DTYPE
from helpers import oracle_render, gpu_render, assert_visbuffer_equal, raster_mode
from test_oracle_kat import meshlet_from_clip_tris, tri_px, fixed_to_ndc, IDENT

pytestmark = pytest.mark.gpu
MODES = pytest.mark.parametrize("binning", [True, False], ids=["binned", "direct"])


def compare(orc, rast, meshlets, w, h, materials=None, clear=(0xFFFFFFFF, 0.0)):
    ofb = orc.Framebuffer(w, h)
    ofb.clear(*clear)
    oc = orc.draw_meshlets(ofb, meshlets, 0, len(meshlets), IDENT, materials=materials, **raster_mode(rast))
    gscene = rast.upload_scene(meshlets, materials)
    gfb = rast.create_framebuffer(w, h)
    gfb.clear(*clear)
    rast.reset_counters()
    rast.draw_meshlets(gfb, gscene, 0, len(meshlets), IDENT)
    assert_visbuffer_equal(ofb, gfb)
    gc = rast.counters()
    assert [gc["TrianglesProcessed"], gc["TrianglesRasterized"], gc["TrianglesClipped"]] == [int(oc[0]), int(oc[1]), int(oc[2])]
    return ofb


@MODES
def test_fill_rule_and_quirk_cases(orc, rast_factory, binning):
    rast = rast_factory(enable_binning=binning)
    w = h = 32
    sq = [tri_px([(8.5, 8.5), (8.5, 24.5), (24.5, 8.5)], w, h), tri_px([(24.5, 8.5), (8.5, 24.5), (24.5, 24.5)], w, h)]
    compare(orc, rast, meshlet_from_clip_tris(sq), w, h)
    # bbox carry quirk
    w, h = 32, 36
    pts = [(-3, 24), (-3, 24 + 96), (-3 + 96, 24)]
    compare(orc, rast, meshlet_from_clip_tris([[(*fixed_to_ndc(x, y, w, h), 0.5) for (x, y) in pts]]), w, h)
    # zero / negative depth, equal-depth ties, degenerate and sliver triangles
    w = h = 64
    tris = [tri_px([(4, 4), (4, 60), (60, 4)], w, h, z=0.0), tri_px([(4, 4), (4, 60), (60, 4)], w, h, z=0.25),
            tri_px([(4, 4), (4, 60), (60, 4)], w, h, z=0.25), tri_px([(10, 10), (10, 10), (30, 30)], w, h),
            tri_px([(5, 40.5), (60, 40.5), (30, 40.6)], w, h), tri_px([(0, 0), (0, 64), (64, 0)], w, h, z=-0.5)]
    compare(orc, rast, meshlet_from_clip_tris(tris), w, h)


@MODES
def test_guard_band_giants_and_int32_wrap(orc, rast_factory, binning):
    """Triangles as large as the guard band: the reference's int32 edge functions may wrap;
    the CUDA path must reproduce the wrapped arithmetic bit for bit."""
    rast = rast_factory(enable_binning=binning)
    w, h = 1920, 1080
    mats = np.zeros(1, dtype=MATERIAL_DTYPE)
    mats["IsDoubleSided"], mats["AlphaCutoff"], mats["TextureId"] = 1, 255, -1
    tris = [[(-1.5, -2.6, 0.3), (1.5, -2.6, 0.3), (0.0, 2.6, 0.3)],
            [(-1.5, 2.6, 0.5), (1.5, 2.67, 0.5), (1.5, -2.67, 0.6)],
            [(-1.49, -2.67, 0.2), (-1.49, 2.67, 0.7), (1.507, 0.0, 0.4)],
            [(-0.9, -0.9, 0.8), (-0.9, 0.9, 0.8), (0.9, -0.9, 0.1)],
            [(1.6, -0.5, 0.5), (0.2, 0.5, 0.5), (0.2, -0.5, 0.5)]]      # beyond the guard band -> dropped (binned) / clipped (direct)
    compare(orc, rast, meshlet_from_clip_tris(tris, material_id=0), w, h, materials=mats)


@MODES
def test_room_big_triangles_and_clipped_count(orc, rast_factory, binning):
    scene = scenes.room_scene()
    rast = rast_factory(enable_binning=binning)
    ofb, oc = oracle_render(orc, scene, **raster_mode(rast))      # direct mode: EnableClipping -> the crossers are clipped and drawn
    assert int(oc[2]) > 0            # the scene does contain guard-band / near-plane crossers
    gfb, gc, _ = gpu_render(rast, scene)
    assert_visbuffer_equal(ofb, gfb, "room")
    assert [gc["TrianglesProcessed"], gc["TrianglesRasterized"], gc["TrianglesClipped"]] == [int(oc[0]), int(oc[1]), int(oc[2])]


@MODES
def test_multi_draw_order_and_no_clear_between(orc, rast_factory, binning):
    """Several DrawMeshlets calls into one framebuffer: later draws only win with strictly greater depth."""
    scene = scenes.instanced_scene(subdivisions=3, instances=8, width=1280, height=720)
    ofb, oc = oracle_render(orc, scene)
    rast = rast_factory(enable_binning=binning)
    for batch in (True, False):
        gfb, gc, _ = gpu_render(rast, scene, batch=batch)
        assert_visbuffer_equal(ofb, gfb, f"batch={batch}")
        assert gc["TrianglesRasterized"] == int(oc[1])
    # drawing the same scene a second time must not change a single pixel (equal depth never passes)
    gfb, _, gscene = gpu_render(rast, scene, batch=False)
    before = (gfb.download_tiled(0), gfb.download_tiled(1))
    gpu_render(rast, scene, batch=False, gscene=gscene, fb=gfb)
    assert np.array_equal(before[0], gfb.download_tiled(0)) and np.array_equal(before[1], gfb.download_tiled(1))


@MODES
def test_cull_bitmap_fused_cull_and_host_meshlets(orc, rast_factory, binning):
    scene = scenes.instanced_scene(subdivisions=3, instances=27, width=1280, height=720)
    ofb, oc = oracle_render(orc, scene, cull=True)
    rast = rast_factory(enable_binning=binning)
    gfb, gc, gscene = gpu_render(rast, scene, cull=True)
    assert_visbuffer_equal(ofb, gfb, "cull bitmap")
    assert gc["TrianglesProcessed"] == int(oc[0]) < scene.num_triangles
    # GPU bitmap == oracle bitmap, visible counts equal
    proj, view = scene.view_proj()
    for node in scene.nodes[:5]:
        ob, on = orc.cull_meshlets(scene.meshlets[node.meshlet_offset:node.meshlet_offset + node.meshlet_count],
                                   orc.frustum_planes(proj, view, node.model))
        gb, gn = rast.cull_meshlets(gscene, node.meshlet_offset, node.meshlet_count, proj, view, node.model)
        assert gn == on and np.array_equal(gb, ob)
        assert np.array_equal(rast.frustum_planes(proj, view, node.model), orc.frustum_planes(proj, view, node.model)[:5])
    # fused frustum test inside the mesh kernel gives the same frame and counters
    rast2 = rast_factory(enable_binning=binning, fused_frustum_cull=True)
    g2 = rast2.upload_scene(scene.meshlets)
    fb2 = rast2.create_framebuffer(scene.width, scene.height)
    fb2.clear(0xFF000000, 0.0)
    rast2.draw_batch(fb2, g2, [dict(offset=n.meshlet_offset, count=n.meshlet_count, object_to_clip=scene.object_to_clip(n),
                                    planes=rast2.frustum_planes(proj, view, n.model)) for n in scene.nodes])
    assert_visbuffer_equal(ofb, fb2, "fused cull")
    assert rast2.counters()["TrianglesProcessed"] == int(oc[0])
    # literal drop-in form: host meshlet pointer uploaded by the call
    node = scene.nodes[3]
    ofb3 = orc.Framebuffer(scene.width, scene.height)
    ofb3.clear(0, 0.0)
    ms = scene.meshlets[node.meshlet_offset:node.meshlet_offset + node.meshlet_count].copy()
    orc.draw_meshlets(ofb3, ms, 0, len(ms), scene.object_to_clip(node))
    fb3 = rast.create_framebuffer(scene.width, scene.height)
    fb3.clear(0, 0.0)
    rast.draw_meshlets_host(fb3, ms, scene.object_to_clip(node))
    assert_visbuffer_equal(ofb3, fb3, "host meshlets")


def test_framebuffer_ops_and_errors(rast_factory):
    from glimpsw_b200 import api
    rast = rast_factory()
    fb = rast.create_framebuffer(64, 32, layers=3)
    fb.clear(0x11223344, 0.5)
    assert np.all(fb.get_pixels(0) == 0x11223344) and np.all(fb.get_pixels(1).view(np.float32) == 0.5)
    fb.clear_layer(2, 7)
    assert np.all(fb.download_tiled(2) == 7)
    pattern = np.arange(64 * 32, dtype=np.uint32)
    fb.upload_tiled(0, pattern)
    from glimpsw_b200.layout import detile
    assert np.array_equal(fb.get_pixels(0), detile(pattern, 64, 32))
    for bad in ((30, 32), (64, 0), (4000, 64)):
        with pytest.raises(api.SwrbError):
            rast.create_framebuffer(*bad)
    with pytest.raises(api.SwrbError):
        fb.clear(0, -1.0)


def _ragged_meshlets(seed, count, spread):
    """Random meshlets with every vertex / triangle count from empty to full (0..128 triangles, 3..64 vertices), random
    windings, tiny to guard-band-sized triangles, some behind the camera plane."""
    from glimpsw_b200.layout import MESHLET_DTYPE
    rng = np.random.default_rng(seed)
    m = np.zeros(count, dtype=MESHLET_DTYPE)
    m["MaterialId"] = 0xFFFFFFFF
    m["AlphaCutoff"] = 255
    for i in range(count):
        nv = int(rng.integers(3, 65))
        nt = int(rng.choice([0, 1, 15, 16, 17, 31, 32, 33, 97, 127, 128, int(rng.integers(0, 129))]))
        centre = rng.uniform(-0.9, 0.9, 3) * (1.0, 1.0, 0.0) + (0.0, 0.0, rng.uniform(0.05, 0.9))
        size = float(rng.choice([0.002, 0.01, 0.05, 0.3, spread]))
        m["Positions"][i, :, :nv] = (centre[:, None] + rng.uniform(-size, size, (3, nv))).astype(np.float32)
        m["Indices"][i, :, :nt] = rng.integers(0, nv, (3, nt))
        m["NumVertices"][i], m["NumTriangles"][i] = nv, nt
        m["BoundCenter"][i] = centre
        m["BoundRadius"][i] = size * 2
    return m


@MODES
@pytest.mark.parametrize("seed,spread", [(1, 0.6), (2, 3.0), (3, 40.0)])
def test_ragged_and_empty_meshlets(orc, rast_factory, binning, seed, spread):
    """Edge cases of the input format: empty meshlets, counts that are not multiples of the 16-wide packet or the 32-wide
    round, degenerate / repeated indices, triangles from sub-pixel to far beyond the guard band, w <= 0 vertices."""
    from glimpsw_b200 import camera as cam
    meshlets = _ragged_meshlets(seed, 97, spread)
    w, h = 1000, 564
    # a perspective-ish matrix: w = z, so vertices with z <= 0 are behind the camera plane
    m = np.zeros((4, 4), dtype=np.float32)
    m[0, 0], m[1, 1], m[2, 2], m[2, 3], m[3, 2] = 1.0, 1.0, 0.0, 1.0, 0.01
    ofb = orc.Framebuffer(w, h)
    ofb.clear(0xFF000000, 0.0)
    clipping = not binning
    oc = orc.draw_meshlets(ofb, meshlets, 0, len(meshlets), m, binned=binning, clipping=clipping)
    rast = rast_factory(enable_binning=binning, enable_clipping=clipping)
    gscene = rast.upload_scene(meshlets)
    fb = rast.create_framebuffer(w, h)
    fb.clear(0xFF000000, 0.0)
    rast.reset_counters()
    rast.draw_meshlets(fb, gscene, 0, len(meshlets), m)
    assert_visbuffer_equal(ofb, fb, f"ragged seed {seed}")
    c = rast.counters()
    assert [c["TrianglesProcessed"], c["TrianglesRasterized"], c["TrianglesClipped"]] == [int(oc[0]), int(oc[1]), int(oc[2])]
    assert int(oc[0]) == int(meshlets["NumTriangles"].astype(np.int64).sum())
    # a draw of zero meshlets and an all-zero cull bitmap leave the framebuffer alone
    before = (fb.download_tiled(0), fb.download_tiled(1))
    rast.draw_batch(fb, gscene, [])
    rast.draw_meshlets(fb, gscene, 5, 0, m)
    rast.draw_meshlets(fb, gscene, 0, len(meshlets), m, cull_bitmap=np.zeros((len(meshlets) + 15) // 16, dtype=np.uint16))
    assert np.array_equal(before[0], fb.download_tiled(0)) and np.array_equal(before[1], fb.download_tiled(1))


def test_ragged_fuzz_many_seeds(orc, rast_factory):
    """24 more seeds of the ragged generator over four framebuffer sizes (incl. the 2896^2 limit) and four triangle-size
    regimes, binned and unbinned + clipping: vis-buffer and counters bit-exact every time."""
    m = np.zeros((4, 4), dtype=np.float32)
    m[0, 0], m[1, 1], m[2, 3], m[3, 2] = 1.0, 1.0, 1.0, 0.01
    rasts = {True: rast_factory(enable_binning=True, enable_clipping=False), False: rast_factory(enable_binning=False, enable_clipping=True)}
    for seed in range(100, 124):
        spread = [0.3, 1.5, 8.0, 60.0][seed % 4]
        w, h = [(1000, 564), (640, 360), (1280, 720), (64, 64)][(seed // 4) % 4]
        if seed == 108:
            w, h = 2896, 2896
        meshlets = _ragged_meshlets(seed, 61, spread)
        for binning, rast in rasts.items():
            ofb = orc.Framebuffer(w, h)
            ofb.clear(0xFF000000, 0.0)
            oc = orc.draw_meshlets(ofb, meshlets, 0, len(meshlets), m, binned=binning, clipping=not binning)
            gscene = rast.upload_scene(meshlets)
            fb = rast.create_framebuffer(w, h)
            fb.clear(0xFF000000, 0.0)
            rast.reset_counters()
            rast.draw_meshlets(fb, gscene, 0, len(meshlets), m)
            assert_visbuffer_equal(ofb, fb, f"fuzz seed {seed} {w}x{h} binning={binning}")
            c = rast.counters()
            assert [c["TrianglesProcessed"], c["TrianglesRasterized"], c["TrianglesClipped"]] == [int(oc[0]), int(oc[1]), int(oc[2])], seed
            fb.destroy()
            gscene.destroy()
