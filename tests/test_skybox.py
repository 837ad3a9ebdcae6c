"""The skybox branch of ShadingContext::Resolve (Shading.cpp:676-679): sky pixels take
SkyboxTex->SampleOctLevel<EnvSampler>(worldPos - ViewPos, 1) from an octahedron-mapped Texture2D<R11G11B10f>.

CPU: known answers for the restated pieces (R11G11B10f bit layout, MapOctahedron, the level-1 bilinear sample, a
constant sky resolving to one hand-computed colour). GPU: libswrb.so against the oracle within the resolve tolerance."""
import numpy as np
import pytest

from glimpsw_b200 import scenes, textures as tx
from helpers import oracle_render, gpu_render

f32 = np.float32


def test_r11g11b10f_bit_layout_known_answers():
    # 1.0 = exponent 15, mantissa 0 in both the 11-bit (5e6m) and the 10-bit (5e5m) field (Texture.h:145-182)
    assert int(tx.pack_r11g11b10f(1.0, 1.0, 1.0)) == (0x3C0 << 21) | (0x3C0 << 10) | 0x1E0
    r, g, b = tx.unpack_r11g11b10f(tx.pack_r11g11b10f(1.0, 0.5, 2.0))
    assert (float(r), float(g), float(b)) == (1.0, 0.5, 2.0)
    # truncation, not rounding: 1 + 63/64 + eps keeps mantissa 63 in 6 bits, 1.999 has mantissa 31 in 5 bits
    r, g, b = tx.unpack_r11g11b10f(tx.pack_r11g11b10f(1.999, 1.999, 1.999))
    assert (float(r), float(b)) == (1.0 + 63 / 64, 1.0 + 31 / 32)
    # clamps: below 2^-15 -> 2^-15; above the largest value -> 130048 / 129024
    r, g, b = tx.unpack_r11g11b10f(tx.pack_r11g11b10f(0.0, 1e9, 1e9))
    assert (float(r), float(g), float(b)) == (2.0 ** -15, 130048.0, 129024.0)


def test_map_octahedron_known_answers(orc):
    assert orc.map_octahedron((0, 0, 1)).tolist() == [0.5, 0.5]            # +Z is the centre of the map
    assert orc.map_octahedron((1, 0, 0)).tolist() == [1.0, 0.5]
    assert orc.map_octahedron((0, -1, 0)).tolist() == [0.5, 0.0]
    assert orc.map_octahedron((0, 0, -5)).tolist() == [1.0, 1.0]            # -Z folds to the corners; length does not matter
    u, v = orc.map_octahedron((0.3, 0.2, 0.5))                             # upper hemisphere: plain L1 projection
    assert abs(u - (0.3 * 0.5 + 0.5)) < 1e-6 and abs(v - (0.2 * 0.5 + 0.5)) < 1e-6


def test_sample_skybox_is_bilinear_on_level_1(orc):
    sky = tx.create_texture(16, 16, 4, 1)
    ramp = np.tile(np.arange(16, dtype=f32)[None, :], (16, 1))            # value = x on level 0
    tx.set_pixels(sky, tx.pack_r11g11b10f(ramp + 1, ramp + 1, ramp + 1), 0)
    tx.generate_mips_hdr(sky)
    lvl1 = tx.unpack_r11g11b10f(tx.get_pixels(sky, 0, 1))[0]
    assert lvl1[0].tolist() == [1.5 + 2 * k for k in range(8)]           # 2x2 box of the ramp
    # direction straight up the +Z axis -> uv (0.5, 0.5) -> texel coordinate 4.0 on level 1 minus the half-texel
    # offset (127/256): between texels 3 and 4 with fraction 129/256
    got = orc.sample_skybox(sky, (0, 0, 1))
    want = 7.5 + (9.5 - 7.5) * (129 / 256)
    assert np.allclose(got, want, rtol=0, atol=1e-6)


def test_constant_sky_resolves_to_tonemapped_constant(orc):
    scene = scenes.torus_knot_scene(40, 16, 320, 200, tex_size=64)
    sky = tx.create_texture(32, 32, 3, 1)
    tx.set_pixels(sky, np.full((32, 32), tx.pack_r11g11b10f(0.5, 0.25, 2.0), dtype=np.uint32), 0)
    tx.generate_mips_hdr(sky)
    ofb, _ = oracle_render(orc, scene)
    n = scene.width * scene.height
    is_sky = ofb.data[1, :n].view(np.float32) <= 0
    exposure = 0.8
    orc.resolve(ofb, scene.meshlets, scene.materials, scene.textures, scene.lights, skybox=sky,
                **scenes.resolve_uniforms(scene, scene.nodes[0], exposure))
    def tone(c):                                                          # Tonemap_Unreal (Shading.cpp:221-226) + Pack
        x = f32(c) * f32(exposure)
        return int(np.clip(np.rint(x / (x + f32(0.155)) * f32(1.019) * f32(255.0)), 0, 255))
    want = 0xFF000000 | tone(0.5) | (tone(0.25) << 8) | (tone(2.0) << 16)
    assert is_sky.any() and np.all(ofb.data[0, :n][is_sky] == want)
    assert np.any(ofb.data[0, :n][~is_sky] != want)


@pytest.mark.gpu
@pytest.mark.parametrize("read_back_first", [False, True], ids=["from_keys", "from_layers"])
def test_skybox_resolve_in_tolerance(orc, rast_factory, read_back_first):
    scene = scenes.torus_knot_scene(120, 48, 1280, 720, tex_size=256)
    sky = tx.procedural_sky_texture(256)
    uni = scenes.resolve_uniforms(scene, scene.nodes[0], 0.9)
    ofb, _ = oracle_render(orc, scene)
    n = scene.width * scene.height
    is_sky = ofb.data[1, :n].view(np.float32) <= 0
    orc.resolve(ofb, scene.meshlets, scene.materials, scene.textures, scene.lights, skybox=sky, **uni)
    rast = rast_factory()
    gfb, _, gscene = gpu_render(rast, scene)          # (gpu_render does not read the framebuffer back)
    if read_back_first:
        gfb.download_tiled(1)
    gscene.set_skybox(sky)
    rast.resolve(gfb, gscene, **uni)
    a = gfb.download_tiled(0).view(np.uint8).reshape(-1, 4).astype(np.int32)
    b = ofb.data[0, :n].view(np.uint8).reshape(-1, 4).astype(np.int32)
    diff = np.abs(a - b)
    assert diff.max() <= 2, f"max abs {diff.max()}/255"
    mse = float((diff[:, :3].astype(np.float64) ** 2).mean())
    assert mse == 0 or 10 * np.log10(255.0 ** 2 / mse) >= 50.0
    assert len(np.unique(gfb.download_tiled(0)[is_sky])) > 10           # the sky is really textured
    # removing the skybox restores colour 0 for sky pixels
    gfb2, _, _ = gpu_render(rast, scene, gscene=gscene)
    gscene.set_skybox(None)
    rast.resolve(gfb2, gscene, **uni)
    assert np.all(gfb2.download_tiled(0)[is_sky] == 0xFF000000)
