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


# ---- panorama -> octahedron import (texutil::LoadOctahedronFromPanoramaHDR, ImageHelpers.cpp:73-104) -------------------------
def _test_panorama(h=128, w=256):
    v, u = np.meshgrid((np.arange(h) + 0.5) / h, (np.arange(w) + 0.5) / w, indexing="ij")
    theta, phi = (u - 0.5) * 2 * np.pi, (v - 0.5) * np.pi
    r = 0.3 + 0.5 * (0.5 + 0.5 * np.sin(theta * 2)) + 6 * np.exp(-((theta - 1.0) ** 2 + (phi + 0.4) ** 2) * 30)      # an HDR sun lobe
    g = 0.4 + 0.4 * (0.5 + 0.5 * np.cos(phi * 3))
    b = 0.6 + 0.3 * (0.5 + 0.5 * np.sin(theta + phi))
    return np.stack([r, g, b], -1).astype(np.float32)


def test_octahedron_from_panorama_looks_up_the_panorama(orc):
    """Sampling the imported octahedron map in direction d gives the panorama's colour at (atan2(d.z, d.x) / tau + 0.5,
    asin(-d.y) / pi + 0.5), up to the two bilinear filters and the 6-bit mantissas in between."""
    pano = _test_panorama()
    sky = tx.octahedron_from_panorama(pano, 6)
    assert (sky.width, sky.height, sky.mip_levels) == (256, 256, 6)
    rng = np.random.default_rng(4)
    worst, errs = 0.0, []
    for _ in range(200):
        d = rng.normal(size=3)
        d /= np.linalg.norm(d)
        if abs(d[1]) > 0.95:
            continue                                                   # the panorama's poles are a single stretched row
        got = orc.sample_skybox(sky, d)                                # level 1 of the map, bilinear
        pu, pv = np.arctan2(d[2], d[0]) / (2 * np.pi) + 0.5, np.arcsin(-d[1]) / np.pi + 0.5
        x, y = pu * 256 - 0.5, pv * 128 - 0.5
        x0, y0 = int(np.floor(x)) % 256, int(np.clip(np.floor(y), 0, 126))
        fx, fy = x - np.floor(x), np.clip(y - y0, 0, 1)
        want = ((pano[y0, x0] * (1 - fx) + pano[y0, (x0 + 1) % 256] * fx) * (1 - fy)
                + (pano[y0 + 1, x0] * (1 - fx) + pano[y0 + 1, (x0 + 1) % 256] * fx) * fy)
        errs.append(float(np.max(np.abs(got - want) / (want + 0.05))))
    # (the sample is level 1 of the map: a 2 x 2 average, which smooths the sun lobe's flank)
    assert max(errs) < 0.25 and float(np.median(errs)) < 0.03, (max(errs), float(np.median(errs)))


def test_octahedron_from_panorama_matches_the_reference_pieces():
    """The same import with the loop of ImageHelpers.cpp:73-104 around the reference's own UnmapOctahedron / SampleLevel /
    WriteTile / GenerateMips (oracle/_ref): every texel of every mip level identical."""
    try:
        from oracle import ref
        if not ref.available():
            raise RuntimeError
    except Exception:
        pytest.skip("reference sources / prebuilt oracle/_ref not on this machine")
    pano = _test_panorama()
    sky = tx.octahedron_from_panorama(pano, 6)
    want = ref.octahedron_from_panorama(tx.hdr_texture_from_pixels(pano, 1), sky)
    n = sky.layer_stride
    assert np.array_equal(np.asarray(sky.data, dtype=np.uint32)[:n], want[:n])
