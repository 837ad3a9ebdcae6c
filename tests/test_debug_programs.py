"""The other shader table and resolve program of the path: OverdrawShader = FS_Overdraw
(Shading.cpp:333-342, :656) and ShadingContext::ResolveDebug (Shading.cpp:734-773).

CPU part: known answers for the oracle restatement. GPU part: libswrb.so against the oracle — the overdraw
counters and its depth layer are integer / order-independent work and must be bit-exact; the debug layers that
only re-colour the colour word (MeshletId, TriangleId, OverdrawPixel, OverdrawQuad) are exact too; the layers that
run ResolveSurface (BaseColor, Normals, MetallicRoughness) are held to the resolve tolerance (<= 2/255)."""
import numpy as np
import pytest

from glimpsw_b200 import api, scenes
from glimpsw_b200 import camera as cam
from glimpsw_b200.layout import MESHLET_DTYPE
from helpers import oracle_render


def _quad_scene(width=64, height=48):
    """Two triangles covering the pixel rectangle [8,24) x [4,20) exactly (vertices on pixel corners)."""
    m = np.zeros(1, dtype=MESHLET_DTYPE)
    m["MaterialId"] = 0xFFFFFFFF
    m["AlphaCutoff"] = 255
    def ndc(px, py):
        return (px / (width / 2) - 1.0, py / (height / 2) - 1.0)
    corners = [ndc(8, 4), ndc(24, 4), ndc(24, 20), ndc(8, 20)]
    for i, (x, y) in enumerate(corners):
        m["Positions"][0, :, i] = (x, y, 0.5)
    m["NumVertices"] = 4
    m["NumTriangles"] = 2
    # front faces have det > 0 in this screen space (y down): pick the winding the oracle keeps
    m["Indices"][0, :, 0] = (0, 2, 1)
    m["Indices"][0, :, 1] = (0, 3, 2)
    return m, np.eye(4, dtype=np.float32)


def test_oracle_overdraw_counts_pixels_and_helper_lanes(orc):
    m, ident = _quad_scene()
    fb = orc.Framebuffer(64, 48)
    fb.clear(0, 0.0)
    c = orc.draw_meshlets(fb, m, 0, 1, ident, overdraw=True)
    if int(c[1]) == 0:      # winding was the culled one: flip
        m["Indices"][0, :, 0] = (0, 1, 2)
        m["Indices"][0, :, 1] = (0, 2, 3)
        c = orc.draw_meshlets(fb, m, 0, 1, ident, overdraw=True)
    assert int(c[1]) == 2
    n = 64 * 48
    col = fb.data[0, :n]
    pix, helper = col >> 16, col & 0xFFFF
    # every pixel of the 16x16 quad is covered exactly once (shared diagonal: top-left rule), nothing outside
    assert int(pix.sum()) == 256 and int(pix.max()) == 1
    # each triangle touches the 4x4 fragments its half of the quad intersects: lanes visited = 16 per touched fragment
    assert int(pix.sum() + helper.sum()) % 16 == 0
    # fragments on the diagonal are touched by both triangles: their 16 lanes are visited twice (1 px + 1 helper)
    assert int(helper.sum()) == 4 * 16        # 4 diagonal fragments, each seen twice, 16 extra lane visits each
    depth = fb.data[1, :n].view(np.float32)
    assert np.all(depth[pix > 0] == 0.5) and np.all(depth[(pix == 0) & (helper == 0)] == 0.0)


def test_oracle_overdraw_saturates_u16_halves(orc):
    m, ident = _quad_scene()
    fb = orc.Framebuffer(64, 48)
    fb.clear(0xFFFE0000 | 0xFFFF, 0.0)       # pixel count 65534, helper count already saturated
    for _ in range(3):
        orc.draw_meshlets(fb, m, 0, 1, ident, overdraw=True)
        m2 = m.copy(); m2["Indices"][0, :, 0] = (0, 1, 2); m2["Indices"][0, :, 1] = (0, 2, 3)
        orc.draw_meshlets(fb, m2, 0, 1, ident, overdraw=True)
    col = fb.data[0, :64 * 48]
    assert int((col >> 16).max()) == 0xFFFF and int((col & 0xFFFF).min()) == 0xFFFF


def test_oracle_debug_layers_known_answers(orc):
    scene = scenes.torus_knot_scene(40, 16, 320, 200, tex_size=64)
    ofb, _ = oracle_render(orc, scene)
    n = scene.width * scene.height
    ids, depth = ofb.data[0, :n].copy(), ofb.data[1, :n].view(np.float32).copy()
    uni = scenes.resolve_uniforms(scene, scene.nodes[0])
    for layer, key in (("MeshletId", ids // 128), ("TriangleId", ids)):
        fb = orc.Framebuffer(scene.width, scene.height); fb.data[...] = ofb.data
        orc.resolve_debug(fb, scene.meshlets, scene.materials, scene.textures, layer, **uni)
        out = fb.data[0, :n]
        want = ((key.astype(np.uint64) * 123456789) & 0xFFFFFF).astype(np.uint32) | 0xFF000000    # Unpack -> Pack round-trips bytes
        surf = depth > 0
        assert np.array_equal(out[surf], want[surf])
        assert set(np.unique(out[~surf]).tolist()) <= {0xFFA0A0A0, 0xFFFFFFFF}
    # checkerboard: 4x4 fragments alternate with (x ^ y) & 4; in tiled order 16 consecutive words are one fragment
    sky_frag = (~(depth > 0)).reshape(-1, 16).all(axis=1)
    frag_x = (np.arange(n // 16) % (scene.width // 4)) * 4
    frag_y = (np.arange(n // 16) // (scene.width // 4)) * 4
    want_bg = np.where(((frag_x ^ frag_y) & 4) != 0, 0xFFA0A0A0, 0xFFFFFFFF).astype(np.uint32)
    assert np.array_equal(out.reshape(-1, 16)[sky_frag][:, 0], want_bg[sky_frag])


# ---------------------------------------------------------------------------------------------------------------------
def _draws(scene):
    return [dict(offset=nd.meshlet_offset, count=nd.meshlet_count, object_to_clip=scene.object_to_clip(nd)) for nd in scene.nodes]


def _oracle_overdraw(orc, scene, binned, clipping, clear=(0, 0.0)):
    fb = orc.Framebuffer(scene.width, scene.height)
    fb.clear(*clear)
    counters = np.zeros(4, dtype=np.uint64)
    for nd in scene.nodes:
        orc.draw_meshlets(fb, scene.meshlets, nd.meshlet_offset, nd.meshlet_count, scene.object_to_clip(nd), materials=scene.materials,
                          counters=counters, binned=binned, clipping=clipping, overdraw=True)
    return fb, counters


@pytest.mark.gpu
@pytest.mark.parametrize("mode", ["binned", "direct_clip", "direct_noclip"])
def test_overdraw_program_bit_exact(orc, rast_factory, mode):
    binned, clipping = mode == "binned", mode == "direct_clip"
    scene = scenes.instanced_scene(subdivisions=3, instances=27, width=640, height=360)    # camera inside: near-plane crossers
    rast = rast_factory(enable_binning=binned, enable_clipping=clipping)
    gscene = rast.upload_scene(scene.meshlets, scene.materials, scene.textures, scene.lights)
    fb = rast.create_framebuffer(scene.width, scene.height)
    fb.clear(0, 0.0)
    rast.reset_counters()
    rast.draw_batch(fb, gscene, _draws(scene), program=api.PROGRAM_OVERDRAW)
    ofb, oc = _oracle_overdraw(orc, scene, binned, clipping)
    n = scene.width * scene.height
    got_c, got_d = fb.download_tiled(0), fb.download_tiled(1)
    assert int((got_c >> 16).max()) >= 2, "scene has no overdraw"
    assert np.array_equal(got_c, ofb.data[0, :n]), f"{int((got_c != ofb.data[0, :n]).sum())} counter words differ"
    assert np.array_equal(got_d, ofb.data[1, :n]), f"{int((got_d != ofb.data[1, :n]).sum())} depth words differ"
    c = rast.counters()
    assert [c["TrianglesProcessed"], c["TrianglesRasterized"], c["TrianglesClipped"]] == [int(oc[0]), int(oc[1]), int(oc[2])]
    # second batch on top of the first: counters accumulate (saturating), no clear in between
    rast.draw_batch(fb, gscene, _draws(scene), program=api.PROGRAM_OVERDRAW)
    for nd in scene.nodes:
        orc.draw_meshlets(ofb, scene.meshlets, nd.meshlet_offset, nd.meshlet_count, scene.object_to_clip(nd), materials=scene.materials,
                          binned=binned, clipping=clipping, overdraw=True)
    assert np.array_equal(fb.download_tiled(0), ofb.data[0, :n])


@pytest.mark.gpu
def test_overdraw_saturation_and_big_triangles(orc, rast_factory):
    """Counters start near the u16 limit; the grid scene seen from close up has triangles spanning many fragments."""
    scene = scenes.grid_scene(6, 5, 512, 384, seed=11)
    rast = rast_factory()
    gscene = rast.upload_scene(scene.meshlets, scene.materials, scene.textures, scene.lights)
    fb = rast.create_framebuffer(scene.width, scene.height)
    fb.clear(0xFFFEFFFD, 0.0)
    rast.draw_batch(fb, gscene, _draws(scene), program=api.PROGRAM_OVERDRAW)
    ofb, _ = _oracle_overdraw(orc, scene, True, False, clear=(0xFFFEFFFD, 0.0))
    n = scene.width * scene.height
    got = fb.download_tiled(0)
    assert np.array_equal(got, ofb.data[0, :n])
    assert int((got >> 16).max()) == 0xFFFF
    assert np.array_equal(fb.download_tiled(1), ofb.data[1, :n])


@pytest.mark.gpu
@pytest.mark.parametrize("layer", ["OverdrawPixel", "OverdrawQuad", "MeshletId", "TriangleId"])
def test_resolve_debug_colour_word_layers_exact(orc, rast_factory, layer):
    scene = scenes.torus_knot_scene(60, 24, 640, 360, tex_size=128)
    rast = rast_factory()
    gscene = rast.upload_scene(scene.meshlets, scene.materials, scene.textures, scene.lights)
    fb = rast.create_framebuffer(scene.width, scene.height)
    uni = scenes.resolve_uniforms(scene, scene.nodes[0])
    overdraw = layer.startswith("Overdraw")
    if overdraw:
        fb.clear(0, 0.0)
        rast.draw_batch(fb, gscene, _draws(scene), program=api.PROGRAM_OVERDRAW)
        ofb, _ = _oracle_overdraw(orc, scene, True, False)
    else:
        fb.clear(0xFF000000, 0.0)
        rast.draw_batch(fb, gscene, _draws(scene))
        ofb, _ = oracle_render(orc, scene)
    rast.resolve_debug(fb, gscene, layer, **uni)
    orc.resolve_debug(ofb, scene.meshlets, scene.materials, scene.textures, layer, **uni)
    n = scene.width * scene.height
    got, want = fb.download_tiled(0), ofb.data[0, :n]
    assert np.array_equal(got, want), f"{layer}: {int((got != want).sum())} pixels differ"
    assert len(np.unique(got)) > 3


@pytest.mark.gpu
@pytest.mark.parametrize("layer", ["BaseColor", "Normals", "MetallicRoughness"])
@pytest.mark.parametrize("read_back_first", [False, True], ids=["from_keys", "from_layers"])
def test_resolve_debug_surface_layers_in_tolerance(orc, rast_factory, layer, read_back_first):
    scene = scenes.torus_knot_scene(60, 24, 640, 360, tex_size=128)
    rast = rast_factory()
    gscene = rast.upload_scene(scene.meshlets, scene.materials, scene.textures, scene.lights)
    fb = rast.create_framebuffer(scene.width, scene.height)
    fb.clear(0xFF000000, 0.0)
    rast.draw_batch(fb, gscene, _draws(scene))
    if read_back_first:
        fb.download_tiled(1)
    uni = scenes.resolve_uniforms(scene, scene.nodes[0])
    rast.resolve_debug(fb, gscene, layer, **uni)
    ofb, _ = oracle_render(orc, scene)
    orc.resolve_debug(ofb, scene.meshlets, scene.materials, scene.textures, layer, **uni)
    n = scene.width * scene.height
    a = fb.download_tiled(0).view(np.uint8).astype(np.int32)
    b = ofb.data[0, :n].view(np.uint8).astype(np.int32)
    assert np.abs(a - b).max() <= 2, f"{layer}: max abs {np.abs(a - b).max()}/255"
    assert (np.abs(a - b).reshape(-1, 4).max(axis=1) > 0).mean() < 0.02


@pytest.mark.gpu
def test_resolve_debug_rejects_layer_none(rast_factory):
    scene = scenes.torus_knot_scene(20, 8, 64, 64, tex_size=16)
    rast = rast_factory()
    gscene = rast.upload_scene(scene.meshlets, scene.materials, scene.textures, scene.lights)
    fb = rast.create_framebuffer(64, 64)
    with pytest.raises(api.SwrbError):
        rast.resolve_debug(fb, gscene, 0, **scenes.resolve_uniforms(scene, scene.nodes[0]))
