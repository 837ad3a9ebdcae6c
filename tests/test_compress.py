"""Meshlet compression (Shading.cpp:292-294): the packed transport format of include/swr_types.h.

CPU: the encoder's error bound, what stays verbatim, and the oracle's decode against the numpy one. GPU: the decode kernel
reproduces the oracle's floats bit for bit, and a packed scene renders (vis-buffer, counters, colour) exactly like its
host-decoded meshlets through the ordinary path."""
import numpy as np
import pytest

from glimpsw_b200 import compress, scenes
from glimpsw_b200.layout import MESHLET_DTYPE, PACKED_MESHLET_DTYPE
from helpers import oracle_render, gpu_render, assert_visbuffer_equal


def _scenes():
    return [scenes.grid_scene(24, 20, 640, 360), scenes.torus_knot_scene(100, 40, 640, 360, tex_size=128),
            scenes.instanced_scene(subdivisions=3, instances=12, width=640, height=360)]


def test_packed_layout_and_error_bound(orc):
    assert PACKED_MESHLET_DTYPE.itemsize == 1376 and MESHLET_DTYPE.itemsize == 1728
    for scene in _scenes():
        m = scene.meshlets
        packed = compress.pack_meshlets(m)
        dec = orc.unpack_meshlets(packed)
        assert np.array_equal(dec.view(np.uint8).reshape(len(m), 1728)[:, :64], m.view(np.uint8).reshape(len(m), 1728)[:, :64])
        for f in ("TexCoords", "NormalTangents", "Indices"):
            assert np.array_equal(dec[f], m[f])
        valid = (np.arange(64)[None, :] < m["NumVertices"].astype(int)[:, None])[:, None, :]
        err = np.abs(dec["Positions"].astype(np.float64) - m["Positions"].astype(np.float64))
        step = packed["Scale"].astype(np.float64)[:, :, None]
        mag = np.abs(m["Positions"]).max()
        assert np.all(np.where(valid, err, 0) <= 0.5 * step + 2e-7 * mag), "quantization error above half a step"
        assert float(np.where(valid, err, 0).max()) > 0                        # it IS lossy
        # the numpy decode and the oracle's C decode agree bit for bit (one rounding: fmaf)
        assert np.array_equal(compress.unpack_meshlets(packed)["Positions"].view(np.uint32), dec["Positions"].view(np.uint32))


def test_degenerate_meshlets_pack(orc):
    m = np.zeros(3, dtype=MESHLET_DTYPE)                                        # empty, single vertex, flat in one axis
    m["NumVertices"] = [0, 1, 3]
    m["Positions"][1, :, 0] = (1.5, -2.25, 1e-3)
    m["Positions"][2, 0, :3] = (0.0, 1.0, 2.0)
    m["Positions"][2, 1, :3] = 7.0
    dec = orc.unpack_meshlets(compress.pack_meshlets(m))
    assert np.array_equal(dec["Positions"][1, :, 0], m["Positions"][1, :, 0])
    assert np.array_equal(dec["Positions"][2, 1, :3], m["Positions"][2, 1, :3]) and np.allclose(dec["Positions"][2, 0, :3], (0, 1, 2), atol=2e-5)


@pytest.mark.gpu
def test_gpu_decode_is_bit_identical_and_renders_like_the_decoded_scene(orc, rast_factory):
    rast = rast_factory()
    for scene in _scenes():
        packed = compress.pack_meshlets(scene.meshlets)
        want = orc.unpack_meshlets(packed)
        gscene = rast.upload_scene(packed, scene.materials, scene.textures, scene.lights)
        got = gscene.download_meshlets()
        assert np.array_equal(got.view(np.uint8), want.view(np.uint8)), "k_unpack_meshlets differs from orc_unpack_meshlets"
        # the decoded scene through the ordinary path
        dscene = scenes.SceneData(scene.name + " (decoded)", want, scene.nodes, scene.camera, scene.width, scene.height)
        dscene.materials, dscene.textures, dscene.lights = scene.materials, scene.textures, scene.lights
        ofb, oc = oracle_render(orc, dscene)
        gfb, gc, _ = gpu_render(rast, dscene, gscene=gscene)
        assert_visbuffer_equal(ofb, gfb, scene.name)
        assert [gc["TrianglesProcessed"], gc["TrianglesRasterized"], gc["TrianglesClipped"]] == [int(oc[0]), int(oc[1]), int(oc[2])]
    # update path: overwrite a range with other packed meshlets
    scene = _scenes()[0]
    packed = compress.pack_meshlets(scene.meshlets)
    gscene = rast.upload_scene(scene.meshlets)
    gscene.update_meshlets(packed[10:50], first=10)
    got = gscene.download_meshlets()
    dec = orc.unpack_meshlets(packed[10:50])
    for f in MESHLET_DTYPE.names:                                              # (field by field: numpy does not carry padding bytes along)
        want = scene.meshlets[f].copy()
        want[10:50] = dec[f]
        assert np.array_equal(np.ascontiguousarray(got[f]).view(np.uint8), np.ascontiguousarray(want).view(np.uint8)), f
