"""ShadingContext::DeferredShader (FS_EncodeGBuffer, Shading.cpp:344-414, :655) on the CUDA path (SWRB_PROGRAM_DEFERRED) vs the
oracle, which tests/test_ref_pin.py pins to the reference's own code: base colour (layer 0), depth (layer 1) and the packed
world normal / metallic / roughness (layer 2) must all be bit-exact, as must the counters. Needs a GPU."""
import numpy as np
import pytest

from glimpsw_b200 import api, scenes
from glimpsw_b200.layout import MATERIAL_DTYPE, NO_MATERIAL

pytestmark = pytest.mark.gpu

MODES = {"binned": dict(enable_binning=True), "direct_clip": dict(enable_binning=False, enable_clipping=True),
         "direct_noclip": dict(enable_binning=False, enable_clipping=False)}


def _oracle_gbuffer(orc, scene, binned, clipping, clear=(0xFF000000, 0.0), ch2=0x12345678, fb=None):
    if fb is None:
        fb = orc.Framebuffer(scene.width, scene.height, 3)
        fb.clear(*clear)
        fb.data[2, :] = ch2
    counters = np.zeros(4, dtype=np.uint64)
    for nd in scene.nodes:
        orc.draw_meshlets(fb, scene.meshlets, nd.meshlet_offset, nd.meshlet_count, scene.object_to_clip(nd), materials=scene.materials,
                          textures=scene.textures, counters=counters, deferred=True, object_to_world3=np.ascontiguousarray(nd.model[0:3, 0:3]),
                          binned=binned, clipping=clipping)
    return fb, counters


def _draws(scene):
    return [dict(offset=nd.meshlet_offset, count=nd.meshlet_count, object_to_clip=scene.object_to_clip(nd),
                 object_to_world3=np.ascontiguousarray(nd.model[0:3, 0:3])) for nd in scene.nodes]


def _gpu_gbuffer(rast, scene, clear=(0xFF000000, 0.0), ch2=0x12345678):
    gscene = rast.upload_scene(scene.meshlets, scene.materials, scene.textures, scene.lights)
    fb = rast.create_framebuffer(scene.width, scene.height, 3)
    fb.clear(*clear)
    fb.upload_tiled(2, np.full(scene.width * scene.height, ch2, dtype=np.uint32))
    rast.reset_counters()
    rast.draw_batch(fb, gscene, _draws(scene), program=api.PROGRAM_DEFERRED)
    return fb, gscene


def _assert_layers(ofb, gfb, what):
    n = ofb.width * ofb.height
    for layer in range(3):
        got = gfb.download_tiled(layer)
        bad = int((got != ofb.data[layer, :n]).sum())
        assert bad == 0, f"{what}: {bad} words of layer {layer} differ of {n}"


@pytest.mark.parametrize("mode", list(MODES), ids=list(MODES))
def test_gbuffer_bit_exact_on_scenes(orc, rast_factory, mode):
    """Textured knot (2 materials, normal maps), alpha-tested double-sided knot, nine-material patchwork with material-less
    meshlets in between (a batch of many runs), close-up with triangles through the camera plane (clipped pieces interpolate
    UVs and normals through the barycentric remap)."""
    binned, clipping = mode == "binned", mode == "direct_clip"
    rast = rast_factory(**MODES[mode])
    for scene in (scenes.torus_knot_scene(120, 48, 960, 540, tex_size=256), scenes.torus_knot_scene(120, 48, 960, 540, tex_size=64, alpha_material=True),
                  scenes.patchwork_scene(20, 16, 640, 360), scenes.closeup_alpha_scene()):
        ofb, oc = _oracle_gbuffer(orc, scene, binned, clipping)
        gfb, _ = _gpu_gbuffer(rast, scene)
        _assert_layers(ofb, gfb, f"{scene.name} [{mode}]")
        c = rast.counters()
        assert [c["TrianglesProcessed"], c["TrianglesRasterized"], c["TrianglesClipped"]] == [int(oc[0]), int(oc[1]), int(oc[2])]
        n = scene.width * scene.height
        assert len(np.unique(ofb.data[2, :n])) > 1000 and (ofb.data[1, :n] != 0).mean() > 0.1


def test_gbuffer_rotated_instances_and_second_batch(orc, rast_factory):
    """Per-draw ObjectToWorld (the normals of rotated instances) and a second batch on top of the first without a clear: the
    depth layer carries over, earlier fragments keep ties."""
    scene = scenes.torus_knot_scene(80, 32, 640, 360, tex_size=128)
    base = scene.nodes[0]
    nodes = []
    for k, ang in enumerate((0.0, 0.9, 2.1)):
        m = np.eye(4, dtype=np.float32)
        c, s = np.cos(ang), np.sin(ang)
        m[0, 0], m[0, 2], m[2, 0], m[2, 2] = c, s, -s, c
        m[3, 0] = (k - 1) * 1.4                                   # (column-major: model[c][r])
        nodes.append(scenes.DrawNode(base.meshlet_offset, base.meshlet_count, np.ascontiguousarray(m)))
    scene.nodes = nodes
    rast = rast_factory()
    ofb, _ = _oracle_gbuffer(orc, scene, True, False)
    gfb, gscene = _gpu_gbuffer(rast, scene)
    _assert_layers(ofb, gfb, "rotated instances")
    scene.nodes = nodes[::-1]                                     # same geometry again, other order: every fragment ties and loses
    _oracle_gbuffer(orc, scene, True, False, fb=ofb)
    rast.draw_batch(gfb, gscene, _draws(scene), program=api.PROGRAM_DEFERRED)
    _assert_layers(ofb, gfb, "second batch")


def test_gbuffer_material_less_scene_writes_depth_only(orc, rast_factory):
    scene = scenes.grid_scene(24, 20, 640, 360)                   # no materials at all
    rast = rast_factory()
    ofb, _ = _oracle_gbuffer(orc, scene, True, False, clear=(0xAABBCCDD, 0.0), ch2=0x0BADF00D)
    gfb, _ = _gpu_gbuffer(rast, scene, clear=(0xAABBCCDD, 0.0), ch2=0x0BADF00D)
    _assert_layers(ofb, gfb, "material-less")
    n = scene.width * scene.height
    assert np.all(gfb.download_tiled(0) == 0xAABBCCDD) and np.all(gfb.download_tiled(2) == 0x0BADF00D) and (ofb.data[1, :n] != 0).mean() > 0.5


def test_gbuffer_errors(rast_factory):
    rast = rast_factory()
    scene = scenes.torus_knot_scene(40, 16, 320, 200, tex_size=64)
    gscene = rast.upload_scene(scene.meshlets, scene.materials, scene.textures, scene.lights)
    fb2 = rast.create_framebuffer(scene.width, scene.height, 2)
    with pytest.raises(api.SwrbError):                            # needs three layers
        rast.draw_batch(fb2, gscene, _draws(scene), program=api.PROGRAM_DEFERRED)
    mats = np.zeros(1, dtype=MATERIAL_DTYPE)
    mats["AlphaCutoff"], mats["TextureId"] = 255, -1
    m = scene.meshlets.copy()
    m["MaterialId"] = 0
    g2 = rast.upload_scene(m, mats)
    fb3 = rast.create_framebuffer(scene.width, scene.height, 3)
    with pytest.raises(api.SwrbError):                            # Material::Texture == nullptr is dereferenced upstream
        rast.draw_batch(fb3, g2, _draws(scene), program=api.PROGRAM_DEFERRED)
