"""Pins the CPU oracle (oracle/oracle*.cpp, the restatement every GPU parity test is checked against) to OUTPUT OF THE
REFERENCE'S OWN CODE: oracle/_ref/libswr_ref.so is src/SwRast/{Rasterizer,Shading,ImageHelpers}.cpp compiled with g++ from
where they lie under /root/reference (oracle/ref_build.py; 26 one-line edits of Clang-only syntax listed in
oracle/ref_edits.py, none arithmetic) in the canonical arithmetic (IEEE, no contraction, approx_rcp = 1/x).

Every comparison is word for word: vis-buffer depth + surface ids, the perf counters, cull bitmaps, depth-pyramid texels,
mip chains, sampled texels — and the resolved COLOUR, which the restatement reproduces bit for bit as well.
CPU only; skipped where neither /root/reference nor a prebuilt library exists, or the CPU lacks AVX-512 VBMI.
"""
import numpy as np
import pytest

from glimpsw_b200 import scenes, textures as tx
from glimpsw_b200.layout import MATERIAL_DTYPE
from helpers import oracle_render
from test_oracle_kat import IDENT, meshlet_from_clip_tris, tri_px

try:
    from oracle import ref
    _HAVE_REF = ref.available()
except Exception:                                     # pragma: no cover
    ref, _HAVE_REF = None, False

pytestmark = pytest.mark.skipif(not _HAVE_REF, reason="reference sources / prebuilt oracle/_ref not on this machine")


@pytest.fixture(scope="module")
def R():
    r = ref.Rasterizer(1)
    yield r
    r.close()


def ref_render(R, scene, cull=False, clear=(0xFF000000, 0.0), binned=True, clipping=False, overdraw=False):
    """helpers.oracle_render with the reference's own Rasterizer / CullMeshlets."""
    fb = ref.Framebuffer(scene.width, scene.height)
    ref.fb_clear(fb, *clear)
    counters = np.zeros(4, dtype=np.uint64)
    proj, view = scene.view_proj()
    for node in scene.nodes:
        bitmap = None
        if cull:
            ms = scene.meshlets[node.meshlet_offset:node.meshlet_offset + node.meshlet_count]
            bitmap, _ = ref.cull_meshlets(ms, proj, view, node.model, view, scene.width, scene.height)
        R.draw_meshlets(fb, scene.meshlets, node.meshlet_offset, node.meshlet_count, scene.object_to_clip(node), cull_bitmap=bitmap,
                        materials=scene.materials, counters=counters, textures=scene.textures if len(scene.textures) else None,
                        binned=binned, clipping=clipping, overdraw=overdraw)
    return fb, counters


def assert_same_fb(a, b, what):
    n = a.width * a.height
    bad_d, bad_c = int((a.data[1, :n] != b.data[1, :n]).sum()), int((a.data[0, :n] != b.data[0, :n]).sum())
    assert bad_d == 0 and bad_c == 0, f"{what}: {bad_d} depth words and {bad_c} id/colour words differ of {n}"


MODES = [dict(binned=True, clipping=False), dict(binned=False, clipping=True), dict(binned=False, clipping=False)]
MODE_IDS = ["binned", "direct_clip", "direct_noclip"]


def test_build_is_the_canonical_one():
    info = ref.build_info()
    assert "Rasterizer" in info and "canonical" in info
    assert "fast" in ref.build_info(fast=True)


# ---- single triangles: classification, early setup, bounding box, edge equations -------------------------------------------
def _random_probe_triangles(n, seed):
    rng = np.random.default_rng(seed)
    tris = []
    for k in range(n):
        kind = k % 5
        w = rng.uniform(0.05, 20.0, 3).astype(np.float32)
        if kind == 0:      # small, on screen
            c = rng.uniform(-0.9, 0.9, 2)
            xy = c + rng.uniform(-0.02, 0.02, (3, 2))
        elif kind == 1:    # medium
            xy = rng.uniform(-1.2, 1.2, (3, 2))
        elif kind == 2:    # guard-band sized
            xy = rng.uniform(-3.5, 3.5, (3, 2))
        elif kind == 3:    # crossing the camera plane
            xy = rng.uniform(-2, 2, (3, 2))
            w[rng.integers(0, 3)] *= -1
        else:              # snapped to pixel centres / tile edges: ties of the fill rule and the bbox carry quirk
            xy = (rng.integers(-60, 60, (3, 2)) * 8 + rng.integers(-1, 2, (3, 2))) / np.array([960.0 * 16, 540.0 * 16]) * 16
        z = rng.uniform(0.0, 1.0, 3).astype(np.float32)
        v = np.zeros((3, 4), dtype=np.float32)
        v[:, 0], v[:, 1], v[:, 2], v[:, 3] = xy[:, 0] * w, xy[:, 1] * w, z * np.abs(w), w
        tris.append(v)
    return tris


@pytest.mark.parametrize("cull_mode", [0, 1, 2], ids=["None", "FrontCCW", "FrontCW"])
@pytest.mark.parametrize("guardband", [True, False], ids=["gb", "nogb"])
def test_probe_triangles_match(orc, cull_mode, guardband):
    """Clipper::ComputeClipCodes, TrianglePacket::Setup, GetRenderBoundingBox, TriangleEdgeVars::Setup on 3,000 triangles."""
    kept = 0
    for v in _random_probe_triangles(3000, 100 + cull_mode):
        a = orc.probe_triangle(v, 1920, 1080, cull_mode, guardband)
        b = ref.probe_triangle(v, 1920, 1080, cull_mode, guardband)
        assert a["cc"] == b["cc"], (v, a, b)
        if not (a["cc"] & 1):
            continue                                   # the reference only sets up accepted lanes
        assert a["keep"] == b["keep"], (v, a, b)
        if a["keep"]:
            kept += 1
            for k in ("pos", "bbox", "edges"):
                assert np.array_equal(a[k], b[k]), (k, v, a, b)
            assert np.array_equal(a["zw"].view(np.uint32), b["zw"].view(np.uint32)), (v, a, b)
    assert kept > 300


# ---- whole frames: vis-buffer + counters ---------------------------------------------------------------------------------
def _scenes_small():
    return [scenes.grid_scene(24, 20, 640, 360), scenes.room_scene(640, 360), scenes.torus_knot_scene(100, 40, 640, 360, tex_size=128),
            scenes.closeup_alpha_scene(), scenes.patchwork_scene(20, 16, 640, 360),
            scenes.torus_knot_scene(120, 48, 960, 540, tex_size=64, alpha_material=True)]


@pytest.mark.parametrize("mode", MODES, ids=MODE_IDS)
def test_visbuffer_and_counters_match_on_scenes(orc, R, mode):
    """Heightfield (sub-pixel to 30 px triangles), room (wall-sized triangles through the camera plane and the guard band),
    textured knot, alpha-tested knots (FS_EncodeSurfaceId<true>: implicit-LOD sampling, clipped-piece barycentric remap),
    nine-material patchwork — in the binned mode and both unbinned modes."""
    for scene in _scenes_small():
        ofb, oc = oracle_render(orc, scene, **mode)
        rfb, rc = ref_render(R, scene, **mode)
        assert_same_fb(ofb, rfb, scene.name)
        assert list(oc[:3]) == list(rc[:3]), f"{scene.name}: counters {oc[:3]} vs {rc[:3]}"
        assert int((ofb.data[1] != 0).sum()) > 1000


def test_config2_full_size_bit_exact(orc, R):
    """BASELINE config C2 as benchmarked: 999,600 triangles at 1920x1080."""
    scene = scenes.grid_scene()
    ofb, oc = oracle_render(orc, scene)
    rfb, rc = ref_render(R, scene)
    assert_same_fb(ofb, rfb, "C2")
    assert list(oc[:3]) == list(rc[:3]) and int(oc[0]) == 999600


def test_config4_bench_view_bit_exact(orc, R):
    """One view of the benchmark batch (BASELINE config C4: 9,994,240 triangles, 126,880 meshlets, 122 draws, frustum-culled by
    CullMeshlets) at 1920x1080, and its resolved colour."""
    from glimpsw_b200 import workloads
    wl = workloads.build("c4_views")
    scene = wl.scene
    scene.camera = wl.cameras[21]
    ofb, oc = oracle_render(orc, scene, cull=True)
    rfb, rc = ref_render(R, scene, cull=True)
    assert_same_fb(ofb, rfb, "C4 view 21")
    assert list(oc[:3]) == list(rc[:3]) and int(oc[0]) > 5_000_000
    uni = scenes.resolve_uniforms(scene, scene.nodes[0])
    orc.resolve(ofb, scene.meshlets, scene.materials, scene.textures, scene.lights, **uni)
    R.resolve(rfb, scene.meshlets, scene.materials, scene.textures, scene.lights, **uni)
    assert_same_fb(ofb, rfb, "C4 view 21 colour")


def test_instanced_scene_with_frustum_cull_and_several_workers(orc, R):
    """C4-style scene (camera inside the lattice): CullMeshlets bitmaps, then the draw, also on 4 worker threads (bins are
    drained independently, so the frame must not depend on the worker count)."""
    scene = scenes.instanced_scene(subdivisions=4, instances=40, width=960, height=540)
    proj, view = scene.view_proj()
    culled = 0
    for node in scene.nodes:
        ms = scene.meshlets[node.meshlet_offset:node.meshlet_offset + node.meshlet_count]
        ob, on = orc.cull_meshlets(ms, orc.frustum_planes(proj, view, node.model))
        rb, rn = ref.cull_meshlets(ms, proj, view, node.model, view, scene.width, scene.height)
        assert on == rn and np.array_equal(ob, rb)
        culled += len(ms) - on
    assert culled > 100
    ofb, oc = oracle_render(orc, scene, cull=True)
    rfb, rc = ref_render(R, scene, cull=True)
    assert_same_fb(ofb, rfb, scene.name)
    assert list(oc[:3]) == list(rc[:3])
    R4 = ref.Rasterizer(4)
    rfb4, rc4 = ref_render(R4, scene, cull=True)
    R4.close()
    assert_same_fb(ofb, rfb4, scene.name + " (4 workers)")
    assert list(oc[:3]) == list(rc4[:3])


def test_maximum_render_size_and_wrapping_edges(orc, R):
    """2896 x 2896 (Rasterizer::MaxRenderSize): guard-band-sized triangles whose 32-bit edge products wrap."""
    w = h = 2896
    tris = [tri_px([(3, 5), (2890, 40), (700, 2893)], w, h, 0.5), tri_px([(-2000, -1500), (4800, 300), (1200, 5200)], w, h, 0.25),
            tri_px([(100.5, 100.5), (108.5, 100.5), (100.5, 108.5)], w, h, 0.75)]
    mats = np.zeros(1, dtype=MATERIAL_DTYPE)
    mats["IsDoubleSided"], mats["AlphaCutoff"], mats["TextureId"] = 1, 255, -1
    m = meshlet_from_clip_tris(tris, material_id=0)
    for mode in MODES:
        ofb = orc.Framebuffer(w, h)
        ofb.clear(0xFFFFFFFF, 0.0)
        oc = orc.draw_meshlets(ofb, m, 0, 1, IDENT, materials=mats, **mode)
        rfb = ref.Framebuffer(w, h)
        rfb.clear(0xFFFFFFFF, 0.0)
        rc = R.draw_meshlets(rfb, m, 0, 1, IDENT, materials=mats, **mode)
        assert_same_fb(ofb, rfb, f"2896^2 {mode}")
        assert list(oc[:3]) == list(rc[:3])
    assert int((ofb.data[1] != 0).sum()) > 1_000_000


def test_ragged_fuzz_matches(orc, R):
    """The ragged generator of tests/test_kat_gpu.py (every vertex / triangle count from empty to full, random windings and
    repeated indices, sub-pixel to guard-band-sized triangles, vertices behind the camera plane) — 36 seeds over four
    framebuffer sizes incl. 2896^2 and four triangle-size regimes, in all three raster modes, with and without a double-sided
    material: vis-buffer and counters identical."""
    from test_kat_gpu import _ragged_meshlets
    m4 = np.zeros((4, 4), dtype=np.float32)                      # w = z: vertices with z <= 0 are behind the camera plane
    m4[0, 0], m4[1, 1], m4[2, 2], m4[2, 3], m4[3, 2] = 1.0, 1.0, 0.0, 1.0, 0.01
    mats = np.zeros(1, dtype=MATERIAL_DTYPE)
    mats["IsDoubleSided"], mats["AlphaCutoff"], mats["TextureId"] = 1, 255, -1
    clipped_total = 0
    for seed in range(200, 236):
        spread = [0.3, 1.5, 8.0, 60.0][seed % 4]
        w, h = [(1000, 564), (640, 360), (1280, 720), (64, 64)][(seed // 4) % 4]
        if seed == 208:
            w = h = 2896
        meshlets = _ragged_meshlets(seed, 41, spread)
        double_sided = seed % 3 == 0
        if double_sided:
            meshlets["MaterialId"] = 0
        for mode in MODES:
            ofb = orc.Framebuffer(w, h)
            ofb.clear(0xFF000000, 0.0)
            oc = orc.draw_meshlets(ofb, meshlets, 0, len(meshlets), m4, materials=mats if double_sided else None, **mode)
            rfb = ref.Framebuffer(w, h)
            rfb.clear(0xFF000000, 0.0)
            rc = R.draw_meshlets(rfb, meshlets, 0, len(meshlets), m4, materials=mats if double_sided else None, **mode)
            assert_same_fb(ofb, rfb, f"fuzz seed {seed} {w}x{h} {mode}")
            assert list(oc[:3]) == list(rc[:3]), (seed, mode, oc[:3], rc[:3])
            clipped_total += int(oc[2])
    assert clipped_total > 1000


def test_stale_cull_mode_of_material_less_meshlets_is_the_one_deliberate_difference(orc, R):
    """the reference's binned path reuses one ShadedMeshlet across meshlets (Rasterizer.cpp:521) and
    ShadeMeshlet only writes CullMode / FragmentShaderId for meshlets WITH a material (Shading.cpp:302-306), so a material-less
    meshlet inherits the previous meshlet's cull mode on that worker — scheduling-dependent state. The restatement (and the
    product) use the struct defaults instead (FrontCCW, Rasterizer.h:85-87), which is what the reference's unbinned path does
    (it constructs a fresh ShadedMeshlet per meshlet, :160). Pinned here so the difference stays exactly this one."""
    from glimpsw_b200.layout import NO_MATERIAL
    mats = np.zeros(1, dtype=MATERIAL_DTYPE)
    mats["IsDoubleSided"], mats["AlphaCutoff"], mats["TextureId"] = 1, 255, -1
    front = [(-0.8, -0.8, 0.5), (-0.8, -0.2, 0.5), (-0.2, -0.8, 0.5)]
    back = [(0.2, 0.2, 0.5), (0.8, 0.2, 0.5), (0.2, 0.8, 0.5)]                  # the other winding
    m = scenes.concat_meshlets([meshlet_from_clip_tris([front], material_id=0), meshlet_from_clip_tris([back], material_id=NO_MATERIAL)])

    def covered(fn, **mode):
        fb = ref.Framebuffer(64, 64)
        fb.clear(0xFFFFFFFF, 0.0)
        c = fn(fb, m, 0, 2, IDENT, materials=mats, **mode)
        ids = fb.data[0][fb.data[1] != 0]
        return sorted(set(int(i) >> 7 for i in ids)), int(c[1])
    drawn = {}
    for name, mode in zip(MODE_IDS, MODES):
        drawn[name] = (covered(orc.draw_meshlets, **mode), covered(R.draw_meshlets, **mode))
    # which of the two windings is the front face is the reference's business; whichever it is, the double-sided meshlet 0 shows
    assert all(0 in d[0][0] and 0 in d[1][0] for d in drawn.values())
    # unbinned: the reference agrees with the restatement about meshlet 1
    assert drawn["direct_clip"][0] == drawn["direct_clip"][1] and drawn["direct_noclip"][0] == drawn["direct_noclip"][1]
    # binned: if the restatement culls meshlet 1, the reference draws it anyway (inherited CullMode::None); otherwise they agree
    o, r = drawn["binned"]
    if 1 not in o[0]:
        assert r[0] == [0, 1] and r[1] == o[1] + 1
    else:
        assert o == r


def test_clipped_and_unclipped_depth_ties_in_one_packet(orc, R):
    """Two triangles of one 16-packet with bit-equal depth on shared pixels, one of them crossing the right guard-band
    plane: the reference draws the packet's accepted lanes first and its clipped pieces afterwards (Rasterizer.cpp:181-249),
    so the accepted triangle wins the tie whatever its primitive index."""
    mats = np.zeros(1, dtype=MATERIAL_DTYPE)
    mats["IsDoubleSided"], mats["AlphaCutoff"], mats["TextureId"] = 1, 255, -1
    big = [(0.1, -0.6, 0.5), (40.0, 0.1, 0.5), (0.1, 0.8, 0.5)]           # leaves the guard band on the right -> clipped
    small = [(0.0, -0.5, 0.5), (0.9, 0.0, 0.5), (0.0, 0.7, 0.5)]          # accepted
    for order in ([big, small], [small, big]):
        m = meshlet_from_clip_tris(order, material_id=0)
        ofb = orc.Framebuffer(256, 128)
        ofb.clear(0xFFFFFFFF, 0.0)
        oc = orc.draw_meshlets(ofb, m, 0, 1, IDENT, materials=mats, binned=False, clipping=True)
        rfb = ref.Framebuffer(256, 128)
        rfb.clear(0xFFFFFFFF, 0.0)
        rc = R.draw_meshlets(rfb, m, 0, 1, IDENT, materials=mats, binned=False, clipping=True)
        assert_same_fb(ofb, rfb, "tie")
        assert list(oc[:3]) == list(rc[:3]) and int(oc[2]) == 1
        ids = np.unique(rfb.data[0][rfb.data[1] != 0])
        assert len(ids) == 2                                                # both visible somewhere, overlap decided by the rule


@pytest.mark.parametrize("mode", MODES, ids=MODE_IDS)
def test_overdraw_program_matches(orc, R, mode):
    """OverdrawShader = FS_Overdraw (Shading.cpp:333-342): pixel / helper-lane counters with u16 saturation + max depth."""
    scene = scenes.instanced_scene(subdivisions=3, instances=27, width=640, height=360)
    ofb = orc.Framebuffer(scene.width, scene.height)
    ofb.clear(0, 0.0)
    oc = np.zeros(4, dtype=np.uint64)
    for nd in scene.nodes:
        orc.draw_meshlets(ofb, scene.meshlets, nd.meshlet_offset, nd.meshlet_count, scene.object_to_clip(nd), materials=scene.materials,
                          counters=oc, overdraw=True, **mode)
    rfb, rc = ref_render(R, scene, clear=(0, 0.0), overdraw=True, **mode)
    assert_same_fb(ofb, rfb, "overdraw")
    assert list(oc[:3]) == list(rc[:3]) and int((ofb.data[0] >> 16).max()) >= 2


@pytest.mark.parametrize("mode", MODES, ids=MODE_IDS)
def test_deferred_gbuffer_program_matches(orc, R, mode):
    """DeferredShader = FS_EncodeGBuffer (Shading.cpp:344-414): all three layers — base colour, depth, packed world normal +
    metallic / roughness — on textured, alpha-tested and clipped geometry and on a scene that mixes in material-less meshlets."""
    for scene in (scenes.torus_knot_scene(100, 40, 640, 360, tex_size=128), scenes.torus_knot_scene(120, 48, 960, 540, tex_size=64, alpha_material=True),
                  scenes.patchwork_scene(20, 16, 640, 360), scenes.closeup_alpha_scene()):
        ofb = orc.Framebuffer(scene.width, scene.height, 3)
        ofb.clear(0xFF000000, 0.0)
        ofb.data[2, :] = 0x12345678
        rfb = ref.Framebuffer(scene.width, scene.height, 3)
        rfb.data[:] = ofb.data
        oc, rc = np.zeros(4, dtype=np.uint64), np.zeros(4, dtype=np.uint64)
        for nd in scene.nodes:
            kw = dict(materials=scene.materials, textures=scene.textures, deferred=True, object_to_world3=np.ascontiguousarray(nd.model[0:3, 0:3]), **mode)
            orc.draw_meshlets(ofb, scene.meshlets, nd.meshlet_offset, nd.meshlet_count, scene.object_to_clip(nd), counters=oc, **kw)
            R.draw_meshlets(rfb, scene.meshlets, nd.meshlet_offset, nd.meshlet_count, scene.object_to_clip(nd), counters=rc, **kw)
        n = scene.width * scene.height
        for layer in range(3):
            bad = int((ofb.data[layer, :n] != rfb.data[layer, :n]).sum())
            assert bad == 0, f"{scene.name}: {bad} words of layer {layer} differ"
        assert list(oc[:3]) == list(rc[:3])
        assert len(np.unique(ofb.data[2, :n])) > 1000


# ---- resolve pass -------------------------------------------------------------------------------------------------------
def _resolve_both(orc, R, scene, exposure=1.0, skybox=None, debug_layer=0):
    ofb, _ = oracle_render(orc, scene)
    rfb = ref.Framebuffer(scene.width, scene.height)
    rfb.data[:] = ofb.data
    uni = scenes.resolve_uniforms(scene, scene.nodes[0], exposure)
    if debug_layer:
        orc.resolve_debug(ofb, scene.meshlets, scene.materials, scene.textures, debug_layer, **uni)
    else:
        orc.resolve(ofb, scene.meshlets, scene.materials, scene.textures, scene.lights, skybox=skybox, **uni)
    R.resolve(rfb, scene.meshlets, scene.materials, scene.textures, scene.lights, skybox=skybox, debug_layer=debug_layer, **uni)
    return ofb, rfb


def test_resolve_colour_is_bit_identical(orc, R):
    """ShadingContext::Resolve: IntersectTriangle, UV gradients, CalcMipLevel, nearest / bilinear sampling with the per-fragment
    filter vote, normal mapping, EvalLighting with directional + point + spot lights, tonemap, pack, light markers."""
    cases = [(scenes.torus_knot_scene(100, 40, 640, 360, tex_size=128), 1.0),
             (scenes.torus_knot_scene(120, 48, 960, 540, tex_size=256, extra_lights=True), 0.7),
             (scenes.patchwork_scene(20, 16, 640, 360), 1.0), (scenes.room_scene(640, 360), 1.3),
             (scenes.torus_knot_scene(120, 48, 960, 540, tex_size=64, alpha_material=True), 1.0)]
    for scene, exposure in cases:
        ofb, rfb = _resolve_both(orc, R, scene, exposure)
        assert_same_fb(ofb, rfb, scene.name + " colour")
        n = scene.width * scene.height
        assert len(scene.textures) == 0 or len(np.unique(ofb.data[0, :n])) > 500


def test_resolve_with_reference_side_matrix_inverse(orc, R):
    """The same with Resolve inverting WorldToClipMat itself (GetInverseScreenProjMatrix on the GLM stand-in, float32
    adjugate) instead of taking the float64-derived matrix the tests pass around: within one 8-bit step on a few pixels."""
    scene = scenes.torus_knot_scene(120, 48, 960, 540, tex_size=256, extra_lights=True)
    ofb, _ = _resolve_both(orc, R, scene)
    rfb = ref.Framebuffer(scene.width, scene.height)
    base, _ = oracle_render(orc, scene)
    rfb.data[:] = base.data
    R.resolve(rfb, scene.meshlets, scene.materials, scene.textures, scene.lights, derive_inverse=True, **scenes.resolve_uniforms(scene, scene.nodes[0]))
    n = scene.width * scene.height
    d = np.abs(ofb.data[0, :n].view(np.uint8).astype(np.int32) - rfb.data[0, :n].view(np.uint8).astype(np.int32))
    assert d.max() <= 1 and (d.reshape(-1, 4).max(axis=1) > 0).mean() < 1e-3


def test_skybox_resolve_is_bit_identical(orc, R):
    scene = scenes.torus_knot_scene(100, 40, 640, 360, tex_size=128)
    ofb, rfb = _resolve_both(orc, R, scene, 0.9, skybox=tx.procedural_sky_texture(256))
    assert_same_fb(ofb, rfb, "skybox")


@pytest.mark.parametrize("layer", range(1, 8), ids=["BaseColor", "Normals", "MetallicRoughness", "MeshletId", "TriangleId", "OverdrawPixel", "OverdrawQuad"])
def test_resolve_debug_layers_are_bit_identical(orc, R, layer):
    scene = scenes.patchwork_scene(20, 16, 640, 360)
    if layer >= 6:          # the overdraw views read the counters FS_Overdraw left in layer 0
        ofb = orc.Framebuffer(scene.width, scene.height)
        ofb.clear(0, 0.0)
        for nd in scene.nodes:
            orc.draw_meshlets(ofb, scene.meshlets, nd.meshlet_offset, nd.meshlet_count, scene.object_to_clip(nd), materials=scene.materials, overdraw=True)
        rfb = ref.Framebuffer(scene.width, scene.height)
        rfb.data[:] = ofb.data
        uni = scenes.resolve_uniforms(scene, scene.nodes[0])
        orc.resolve_debug(ofb, scene.meshlets, scene.materials, scene.textures, layer, **uni)
        R.resolve(rfb, scene.meshlets, scene.materials, scene.textures, scene.lights, debug_layer=layer, **uni)
    else:
        ofb, rfb = _resolve_both(orc, R, scene, debug_layer=layer)
    assert_same_fb(ofb, rfb, f"debug layer {layer}")


# ---- textures, depth pyramid, HiZ, framebuffer helpers -------------------------------------------------------------------
def test_mip_chain_and_implicit_lod_sampling_match(orc):
    tex = tx.procedural_material_texture(256, seed=3, with_nmr=True)
    assert np.array_equal(ref.generate_mips(tex), np.ascontiguousarray(tex.data, dtype=np.uint32)), "Texture2D::GenerateMips vs textures.generate_mips"
    rng = np.random.default_rng(9)
    for k in range(400):
        scale = 10.0 ** rng.uniform(-4, 0.5)                                 # magnified ... heavily minified
        u0, v0 = rng.uniform(-2, 2, 2)
        du, dv = rng.uniform(-1, 1, (2, 2)) * scale
        xs, ys = np.meshgrid(np.arange(4), np.arange(4))
        u = (u0 + xs * du[0] + ys * du[1]).astype(np.float32).reshape(16)
        v = (v0 + xs * dv[0] + ys * dv[1]).astype(np.float32).reshape(16)
        want = np.zeros(16, dtype=np.uint32)
        descs, keep = orc._texture_descs([tex])
        orc.lib().orc_sample_implicit_lod_4x4(descs, orc._p(u), orc._p(v), orc._p(want))
        assert np.array_equal(ref.sample_implicit_lod_4x4(tex, u, v), want), f"fragment {k}"


def test_depth_pyramid_and_hiz_cull_match(orc):
    scene = scenes.instanced_scene(subdivisions=4, instances=64, width=1920, height=1080)
    fb, _ = oracle_render(orc, scene)
    hw, hh = orc.hiz_dims(scene.width, scene.height)
    opyr, rpyr = tx.create_texture(hw, hh, 16, 1), tx.create_texture(hw, hh, 16, 1)
    orc.downsample_depth(fb, opyr)
    ref.downsample_depth(fb, rpyr)
    for m in range(opyr.mip_levels):                                         # every texel the reference's recursion writes
        t = 1 << (m + 1)
        blk = (4 if m == opyr.mip_levels - 1 else 8) * t
        ty, txx = -(-scene.height // blk) * (blk // t), -(-scene.width // blk) * (blk // t)
        y, x = np.meshgrid(np.arange(ty), np.arange(txx), indexing="ij")
        off = int(opyr.mip_offsets[m]) + tx.texel_offset(x, y, opyr.row_shift - m)
        assert np.array_equal(np.asarray(opyr.data)[off], np.asarray(rpyr.data)[off]), f"pyramid level {m}"
    proj, view = scene.view_proj()
    removed = 0
    for node in scene.nodes:
        ms = scene.meshlets[node.meshlet_offset:node.meshlet_offset + node.meshlet_count]
        ob, on = orc.cull_meshlets_hiz(ms, proj, view, node.model, view, scene.width, scene.height, opyr)
        rb, rn = ref.cull_meshlets(ms, proj, view, node.model, view, scene.width, scene.height, rpyr)
        assert on == rn and np.array_equal(ob, rb)
        fb_, fn = ref.cull_meshlets(ms, proj, view, node.model, view, scene.width, scene.height)
        removed += fn - rn
    assert removed > 50


def test_framebuffer_clear_and_get_pixels_match(orc):
    fb = orc.Framebuffer(64, 40)
    fb.data[:] = np.random.default_rng(2).integers(0, 2**32, fb.data.shape, dtype=np.uint32)
    for layer in (0, 1):
        assert np.array_equal(ref.fb_get_pixels(fb, layer), fb.get_pixels(layer))
    a, b = orc.Framebuffer(64, 40), ref.Framebuffer(64, 40)
    a.clear(0x12345678, 0.25)
    ref.fb_clear(b, 0x12345678, 0.25)
    n = 64 * 40
    assert np.array_equal(a.data[:, :n], b.data[:, :n])


# ---- what the upstream build flags change (information, loosely bounded) ------------------------------------------------------
def test_fast_math_build_changes_depth_bits_only(orc):
    """The same sources compiled -Ofast -mrecip with the hardware's vrcp14ps / vrsqrt14ps (the reference builds with
    -ffast-math, src/SwRast/CMakeLists.txt:6): coverage and counters stay, a handful of ids flip on depth near-ties, depth
    words move in the low bits. This is why parity is defined on the canonical arithmetic."""
    scene = scenes.grid_scene(40, 40, 960, 540)
    ofb, oc = oracle_render(orc, scene)
    RF = ref.Rasterizer(1, fast=True)
    rfb, rc = ref_render(RF, scene)
    RF.close()
    n = scene.width * scene.height
    covered = ofb.data[1, :n] != 0
    assert np.array_equal(covered, rfb.data[1, :n] != 0)
    assert (ofb.data[0, :n] != rfb.data[0, :n]).mean() < 1e-3
    assert abs(int(oc[1]) - int(rc[1])) <= 16 and int(oc[0]) == int(rc[0])
    rel = np.abs(ofb.data[1, :n].view(np.float32)[covered] / rfb.data[1, :n].view(np.float32)[covered] - 1.0)
    assert float(np.quantile(rel, 0.999)) < 1e-3
