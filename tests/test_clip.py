"""Polygon clipper of the unbinned path: Clipper::ClipTriangles + DrawTriangle<FS, true>
(Rasterizer.cpp:209-249, :398-491). CPU tests pin the oracle's clipper to properties that do not depend on its own
implementation; the GPU tests demand bit-equality of the CUDA path (k_clip_triangles) with the oracle."""
import numpy as np
import pytest

from glimpsw_b200 import scenes
from glimpsw_b200.layout import MATERIAL_DTYPE, detile
from helpers import oracle_render, gpu_render, assert_visbuffer_equal, raster_mode
from test_oracle_kat import meshlet_from_clip_tris, IDENT

# object (x, y, z) -> clip (x, y, 0.1, z): w follows the object-space z, like a reverse-Z projection with near 0.1
W_FROM_Z = np.zeros((4, 4), dtype=np.float32)      # [c, r]
W_FROM_Z[0, 0] = W_FROM_Z[1, 1] = 1.0
W_FROM_Z[2, 3] = 1.0
W_FROM_Z[3, 2] = 0.1


def render(orc, tris, matrix, w, h, **kw):
    fb = orc.Framebuffer(w, h)
    fb.clear(0xFFFFFFFF, 0.0)
    mats = np.zeros(1, dtype=MATERIAL_DTYPE)
    mats["IsDoubleSided"], mats["AlphaCutoff"], mats["TextureId"] = 1, 255, -1
    c = orc.draw_meshlets(fb, meshlet_from_clip_tris(tris, material_id=0), 0, 1, matrix, materials=mats, **kw)
    return detile(fb.data[0], w, h), detile(fb.data[1], w, h).view(np.float32), c


def test_counters_follow_the_three_modes(orc):
    """Binned: non-trivial triangles are counted and dropped (:567-569). Unbinned without clipping: dropped, not
    counted (:209-210). Unbinned with clipping: counted, and every surviving piece counts as rasterized (:247)."""
    tris = [[(-0.5, -0.5, 1.0), (0.5, -0.5, 1.0), (0.0, 0.5, 1.0)],        # w = 1 everywhere: trivially accepted
            [(-0.3, -0.3, 2.0), (0.3, -0.3, 2.0), (0.0, 0.2, -1.0)]]       # third vertex behind the camera
    ids_b, _, cb = render(orc, tris, W_FROM_Z, 64, 64)
    ids_n, _, cn = render(orc, tris, W_FROM_Z, 64, 64, binned=False, clipping=False)
    ids_c, _, cc = render(orc, tris, W_FROM_Z, 64, 64, binned=False, clipping=True)
    assert [int(x) for x in cb[:3]] == [2, 1, 1]
    assert [int(x) for x in cn[:3]] == [2, 1, 0]
    assert int(cc[0]) == 2 and int(cc[2]) == 1 and int(cc[1]) >= 3          # the crosser becomes >= 2 pieces
    assert np.array_equal(ids_b, ids_n)
    assert (ids_c == 1).sum() > 0 and (ids_b == 1).sum() == 0                # surface id of the pieces = the original prim
    assert np.array_equal(ids_c == 0, ids_b == 0) or ((ids_c == 0) <= (ids_b == 0)).all()


def test_frustum_clip_matches_guard_band_rendering(orc):
    """A triangle that leaves the viewport but stays inside the guard band is rasterized unclipped when the guard band
    is on, and clipped at the frustum planes when it is off (bx = by = 1, Rasterizer.cpp:155-156). Both must show
    the same surface in the viewport; only pixels on the re-snapped edges may differ."""
    w, h = 256, 128
    tris = [[(-1.8, -0.7, 0.5), (1.6, -1.4, 0.25), (0.3, 1.9, 0.75)],
            [(0.2, -2.5, 0.9), (2.2, 0.4, 0.6), (-0.4, 0.6, 0.3)]]
    ids_g, d_g, cg = render(orc, tris, IDENT, w, h, binned=False, clipping=True, guardband=True)
    ids_c, d_c, cc = render(orc, tris, IDENT, w, h, binned=False, clipping=True, guardband=False)
    assert int(cg[2]) == 0 and int(cc[2]) == 2 and int(cc[1]) > 2
    differ = ids_g != ids_c
    assert differ.mean() < 2e-3
    same = ~differ & (ids_g != 0xFFFFFFFF)
    assert same.sum() > 0.5 * w * h
    assert np.abs(d_g[same] - d_c[same]).max() < 3e-4     # 1/16-px vertex snap x depth slope


def test_camera_plane_clip_matches_analytic_preclip(orc):
    """A triangle with one vertex behind the camera: the oracle's clipped pieces must cover what an analytically
    pre-clipped polygon (float64 intersection with w = near, then drawn as ordinary triangles) covers."""
    w, h = 128, 128
    a, b, c = np.array([-0.6, -0.5, 1.5]), np.array([0.7, -0.4, 1.2]), np.array([0.1, 0.3, -0.4])
    ids_c, d_c, cc = render(orc, [[a, b, c]], W_FROM_Z, w, h, binned=False, clipping=True)
    assert int(cc[2]) == 1
    near = 0.1

    def cut(p, q):       # point on pq with object z (= clip w) == near
        t = (p[2] - near) / (p[2] - q[2])
        return p + t * (q - p)
    bc, ca = cut(b, c), cut(a, c)
    ids_p, d_p, cp = render(orc, [[a, b, bc], [a, bc, ca]], W_FROM_Z, w, h, binned=False, clipping=True)
    cov_c, cov_p = ids_c != 0xFFFFFFFF, ids_p != 0xFFFFFFFF
    assert cov_p.sum() > 1000
    assert (cov_c != cov_p).mean() < 3e-3
    both = cov_c & cov_p
    assert np.abs(d_c[both] - d_p[both]).max() < 3e-3


def test_fully_outside_and_degenerate_results(orc):
    """Triangles whose every vertex is beyond one plane are rejected before the clipper; a crosser whose visible part
    vanishes under clipping yields no piece; nothing may touch the framebuffer."""
    tris = [[(3.0, 0.0, 1.0), (4.0, 0.0, 1.0), (3.5, 1.0, 1.0)],            # all right of the frustum
            [(0.0, 0.0, -1.0), (1.0, 0.0, -2.0), (0.0, 1.0, -3.0)],         # all behind the camera
            [(0.0, 0.0, 0.1), (0.2, 0.0, 0.05), (0.0, 0.2, 0.05)]]          # touches the camera plane in one point
    ids, _, c = render(orc, tris, W_FROM_Z, 64, 64, binned=False, clipping=True)
    assert int(c[0]) == 3 and int(c[1]) == 0
    assert (ids == 0xFFFFFFFF).all()


# ---- CUDA path -----------------------------------------------------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("guardband", [True, False], ids=["guardband", "frustum"])
def test_gpu_clipped_kat_triangles(orc, rast_factory, guardband):
    rast = rast_factory(enable_binning=False, enable_clipping=True, enable_guardband=guardband)
    mats = np.zeros(1, dtype=MATERIAL_DTYPE)
    mats["IsDoubleSided"], mats["AlphaCutoff"], mats["TextureId"] = 1, 255, -1
    cases = [(IDENT, 256, 128, [[(-1.8, -0.7, 0.5), (1.6, -1.4, 0.25), (0.3, 1.9, 0.75)], [(0.2, -2.5, 0.9), (2.2, 0.4, 0.6), (-0.4, 0.6, 0.3)]]),
             (IDENT, 1920, 1080, [[(1.6, -0.5, 0.5), (0.2, 0.5, 0.5), (0.2, -0.5, 0.5)], [(-4.0, -3.0, 0.4), (4.0, -3.0, 0.4), (0.0, 5.0, 0.6)]]),
             (W_FROM_Z, 128, 128, [[(-0.6, -0.5, 1.5), (0.7, -0.4, 1.2), (0.1, 0.3, -0.4)], [(-0.3, -0.3, 2.0), (0.3, -0.3, 2.0), (0.0, 0.2, -1.0)],
                                   [(0.0, 0.0, 0.1), (0.2, 0.0, 0.05), (0.0, 0.2, 0.05)], [(-30.0, -20.0, 0.5), (25.0, -22.0, 0.02), (1.0, 40.0, 3.0)]])]
    for matrix, w, h, tris in cases:
        meshlets = meshlet_from_clip_tris(tris, material_id=0)
        ofb = orc.Framebuffer(w, h)
        ofb.clear(0xFFFFFFFF, 0.0)
        oc = orc.draw_meshlets(ofb, meshlets, 0, 1, matrix, materials=mats, guardband=guardband, binned=False, clipping=True)
        gscene = rast.upload_scene(meshlets, mats)
        gfb = rast.create_framebuffer(w, h)
        gfb.clear(0xFFFFFFFF, 0.0)
        rast.reset_counters()
        rast.draw_meshlets(gfb, gscene, 0, 1, matrix)
        assert_visbuffer_equal(ofb, gfb, f"{w}x{h}")
        gc = rast.counters()
        assert [gc["TrianglesProcessed"], gc["TrianglesRasterized"], gc["TrianglesClipped"]] == [int(oc[0]), int(oc[1]), int(oc[2])]
        assert int(oc[2]) > 0 or (guardband and w == 256)      # (the first case stays inside the guard band)


@pytest.mark.gpu
@pytest.mark.parametrize("clipping", [True, False], ids=["clip", "noclip"])
def test_gpu_room_and_closeup_alpha_scene(orc, rast_factory, clipping):
    """Whole scenes on the unbinned path: big triangles crossing the camera plane and the guard band, opaque (room) and
    alpha-tested with the barycentric remap of clipped pieces (close-up knot). clipping=False: dropped, uncounted."""
    for scene in (scenes.room_scene(), scenes.closeup_alpha_scene()):
        rast = rast_factory(enable_binning=False, enable_clipping=clipping)
        ofb, oc = oracle_render(orc, scene, **raster_mode(rast))
        gfb, gc, _ = gpu_render(rast, scene)
        assert_visbuffer_equal(ofb, gfb, scene.name)
        assert [gc["TrianglesProcessed"], gc["TrianglesRasterized"], gc["TrianglesClipped"]] == [int(oc[0]), int(oc[1]), int(oc[2])]
        assert (int(oc[2]) > 0) == clipping
