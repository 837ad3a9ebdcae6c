"""Known answers for the resolve-pass restatement (oracle/oracle_resolve.cpp), derived from the reference SOURCE's
formulas in float64 / integer numpy, not from the oracle's own output:

  * a screen-filling quad textured 1:1 must reproduce the texture texel for texel (IntersectTriangle + UV interpolation +
    SampleLevel/SampleLinear addressing, Shading.cpp:417-464, :516-527; Texture.h:412-459, :506-575);
  * a quad whose clip-space w varies must interpolate UVs perspective-correctly (the analytic projective map), where an
    affine interpolation would be ~10 texels off;
  * the decoded vertex normal of a flat quad is the hand-computed octahedron decode (Texture.h:289-296);
  * the shaded colour of a flat quad under one directional light equals the Filament-style BRDF of Shading.cpp:17-33,
    :602-645 evaluated in float64 (GGX D, Smith-GGX-correlated-fast V, Schlick F, Lambert, 0.05 ambient, Unreal
    tonemap :221-226, RNE pack), within one 8-bit step.

CPU only. The probes are ShadingContext::ResolveDebug layers (BaseColor, Normals) and ShadingContext::Resolve itself."""
import numpy as np
import pytest

from glimpsw_b200 import camera as cam, scenes, textures as tx
from glimpsw_b200.layout import MATERIAL_DTYPE, LIGHT_DTYPE, detile

f32 = np.float32


def quad_scene(size, w_slope=0.0, tex=None):
    """Two triangles over object-space [-1,1]^2 at z = 0, UV = (x+1)/2, (y+1)/2, normal +Z, tangent +X; clip = (x, y, 0.5, 1 + w_slope*x)."""
    pos = np.array([[-1, -1, 0], [1, -1, 0], [1, 1, 0], [-1, 1, 0]], dtype=f32)
    uv = (pos[:, :2] + 1) / 2
    nrm = np.tile(np.array([0, 0, 1], dtype=f32), (4, 1))
    tan = np.tile(np.array([1, 0, 0, 1], dtype=f32), (4, 1))
    meshlets = scenes.meshletize(pos, np.array([[0, 2, 1], [0, 3, 2], [0, 1, 2], [0, 2, 3]]), uv=uv, normals=nrm, tangents=tan, material_id=0)
    m = np.zeros((4, 4), dtype=f32)                     # column-major [c, r]
    m[0, 0], m[1, 1], m[2, 2] = 1.0, 1.0, 0.1           # (the z column only keeps the matrix invertible: the quad has z = 0)
    m[0, 3] = w_slope                                   # w += w_slope * x
    m[3, 2], m[3, 3] = 0.5, 1.0                         # z_clip = 0.5 + 0.1 z, w = 1 + w_slope * x
    mats = np.zeros(1, dtype=MATERIAL_DTYPE)
    mats["AlphaCutoff"] = 255                           # single sided: both windings are in the list, exactly one of each pair survives culling
    return meshlets, m, mats, [tex]


def render(orc, size, meshlets, m, mats, textures):
    fb = orc.Framebuffer(size, size)
    fb.clear(0xFF000000, 0.0)
    c = orc.draw_meshlets(fb, meshlets, 0, len(meshlets), m, materials=mats)
    assert int(c[1]) == 2, "exactly the two front-facing triangles are rasterized"
    return fb


def uniforms(m, size, view_pos=(0.0, 0.0, 3.0), exposure=1.0):
    return dict(world_to_clip=m, object_to_clip=m, object_to_world3=np.eye(3, dtype=f32),
                inv_screen_proj=cam.inverse_screen_proj(m, size, size), view_pos=np.asarray(view_pos, dtype=f32), exposure=exposure)


def gradient_texture(size, layers=1):
    tex = tx.create_texture(size, size, 8, layers)
    y, x = np.meshgrid(np.arange(size), np.arange(size), indexing="ij")
    tx.set_pixels(tex, ((x * 3 + 10) | ((y * 3 + 20) << 8) | (77 << 16) | (255 << 24)).astype(np.uint32), 0)
    tx.generate_mips(tex)
    return tex


def test_one_to_one_textured_quad_reproduces_the_texture(orc):
    size = 64
    tex = gradient_texture(size)
    meshlets, m, mats, textures = quad_scene(size, tex=tex)
    fb = render(orc, size, meshlets, m, mats, textures)
    orc.resolve_debug(fb, meshlets, mats, textures, "BaseColor", **uniforms(m, size))
    img = detile(fb.data[0, :size * size], size, size)
    # pixel (px, py) centre -> u = (px + 0.5) / 64 -> fixed-point 256 * px + 128 -> minus the 127 half-texel offset = texel px
    # with fraction 1/256: lerp16 adds ((3 << 7) + (1 << 14)) >> 15 = 0 for neighbours that differ by 3 -> the texel itself
    assert np.array_equal(img, tx.get_pixels(tex, 0, 0))


def test_perspective_correct_uv(orc):
    size, tsize = 256, 64
    tex = tx.create_texture(tsize, tsize, 1, 1)
    y, x = np.meshgrid(np.arange(tsize), np.arange(tsize), indexing="ij")
    tx.set_pixels(tex, ((x * 4) | ((y * 4) << 8) | (255 << 24)).astype(np.uint32), 0)      # texel index in R and G
    meshlets, m, mats, textures = quad_scene(size, w_slope=0.5, tex=tex)
    fb = render(orc, size, meshlets, m, mats, textures)
    covered = detile(fb.data[1, :size * size], size, size).view(np.float32) > 0
    orc.resolve_debug(fb, meshlets, mats, textures, "BaseColor", **uniforms(m, size))
    img = detile(fb.data[0, :size * size], size, size)
    # ground truth: NDC x' = x / (1 + 0.5 x)  =>  x = x' / (1 - 0.5 x'), y = y' (1 + 0.5 x); u = (x + 1) / 2, v = (y + 1) / 2
    py, px = np.meshgrid(np.arange(size) + 0.5, np.arange(size) + 0.5, indexing="ij")
    xn, yn = px / (size / 2) - 1.0, py / (size / 2) - 1.0
    xo = xn / (1.0 - 0.5 * xn)
    yo = yn * (1.0 + 0.5 * xo)
    u_true, v_true = (xo + 1) / 2, (yo + 1) / 2
    u_affine = (xn + 1) / 2                                               # what screen-space-linear interpolation would give
    inner = covered & (u_true > 0.05) & (u_true < 0.95) & (v_true > 0.05) & (v_true < 0.95)
    assert inner.sum() > 10000
    # bilinear magnification: channel = 4 * (u * 64 - 0.5) inside the texture, up to the 8-bit lerp rounding
    u_got = ((img & 255).astype(np.float64) / 4 + 0.5) / tsize
    v_got = (((img >> 8) & 255).astype(np.float64) / 4 + 0.5) / tsize
    assert np.abs(u_got - u_true)[inner].max() < 1.0 / tsize and np.abs(v_got - v_true)[inner].max() < 1.0 / tsize
    assert np.abs(u_affine - u_true)[inner].max() > 5.0 / tsize            # the test can tell the two apart


def test_flat_quad_normal_layer_is_the_hand_decoded_octahedron(orc):
    size = 32
    meshlets, m, mats, textures = quad_scene(size, tex=gradient_texture(16))
    # +Z encodes to octahedron (0.5, 0.5) -> unorm8 128 (rint of 127.5, ties to even); decode: 128/255*2-1 = 1/255 in x and y,
    # z = 1 - 2/255, normalised, then * 0.5 + 0.5 and RNE-packed (Shading.cpp:752, Texture.h:55-67)
    assert int(meshlets["NormalTangents"][0, 0]) & 0xFFFF == 128 | (128 << 8)
    n = np.array([1 / 255, 1 / 255, 1 - 2 / 255])
    n = n / np.linalg.norm(n)
    want = [int(np.rint((c * 0.5 + 0.5) * 255)) for c in n]
    assert want == [128, 128, 255]
    fb = render(orc, size, meshlets, m, mats, textures)
    orc.resolve_debug(fb, meshlets, mats, textures, "Normals", **uniforms(m, size))
    img = fb.data[0, :size * size]
    assert np.all(img == (0xFF000000 | want[0] | (want[1] << 8) | (want[2] << 16)))


@pytest.mark.parametrize("metal,rough", [(0, 128), (255, 64), (128, 230)])
def test_directional_light_matches_float64_brdf(orc, metal, rough):
    size = 64
    tex = tx.create_texture(16, 16, 1, 2)
    albedo = (200, 150, 100)
    tx.set_pixels(tex, np.full((16, 16), albedo[0] | (albedo[1] << 8) | (albedo[2] << 16) | (255 << 24), dtype=np.uint32), 0)
    tx.set_pixels(tex, np.full((16, 16), (metal << 16) | (rough << 24), dtype=np.uint32), 1)      # normal.xy = 0: normal mapping off (:554)
    meshlets, m, mats, textures = quad_scene(size, tex=tex)
    light = np.zeros(1, dtype=LIGHT_DTYPE)
    ldir = np.array([0.3, -0.2, -1.0]); ldir /= np.linalg.norm(ldir)
    light["Type"], light["Direction"], light["Color"], light["Intensity"] = 0, ldir.astype(f32), (1.0, 0.9, 0.8), 900.0
    exposure, view = 0.7, np.array([0.2, -0.1, 3.0])
    fb = render(orc, size, meshlets, m, mats, textures)
    orc.resolve(fb, meshlets, mats, textures, light, **uniforms(m, size, view, exposure))
    img = detile(fb.data[0, :size * size], size, size)

    # ---- float64 model of Shading.cpp:602-645 per pixel
    py, px = np.meshgrid(np.arange(size) + 0.5, np.arange(size) + 0.5, indexing="ij")
    world = np.stack([px / (size / 2) - 1.0, py / (size / 2) - 1.0, np.zeros_like(px)], -1)      # the quad is the z = 0 plane, clip = object
    n = np.array([1 / 255, 1 / 255, 1 - 2 / 255]); n /= np.linalg.norm(n)
    srgb = lambda c: (((c << 8) + 255) ** 2 >> 16) / 65535.0                                    # RGBA8u::UnpackSrgb (Texture.h:37-54)
    base = np.array([srgb(c) for c in albedo])
    metallic, roughness = metal / 255.0, rough / 255.0
    a = max(roughness * roughness, 1e-4)
    f0 = 0.04 + (base - 0.04) * metallic                                                        # lerp(0.16 * 0.5^2, base, metallic)
    diffuse = base * (1 - metallic)
    V = view - world; V /= np.linalg.norm(V, axis=-1, keepdims=True)
    L = -ldir
    NoV = np.abs(V @ n) + 1e-5
    NoL = float(L @ n)
    H = V + L; H /= np.linalg.norm(H, axis=-1, keepdims=True)
    NoH, LoH = np.clip(H @ n, 0, 1), np.clip(H @ L, 0, 1)
    k = a / (1 - NoH ** 2 + (NoH * a) ** 2)
    D = k * k / np.pi
    Vis = 0.5 / ((2 * NoL * NoV) * (1 - a) + (NoL + NoV) * a)
    f = (1 - LoH) ** 5
    atten = 900.0 * (exposure * 0.001)
    out = np.zeros(px.shape + (3,))
    for c in range(3):
        F = f + f0[c] * (1 - f)
        out[..., c] = (diffuse[c] / np.pi + D * Vis * F) * float(light["Color"][0, c]) * max(NoL * atten, 0) + base[c] * 0.05
    x = out * exposure
    want = np.clip(np.rint(x / (x + 0.155) * 1.019 * 255), 0, 255).astype(np.int64)
    got = np.stack([(img >> (8 * c)) & 255 for c in range(3)], -1).astype(np.int64)
    assert np.abs(got - want).max() <= 1, f"max diff {np.abs(got - want).max()} of 255"
    assert (np.abs(got - want).max(axis=-1) > 0).mean() < 0.1
    assert want.max() > 40 and len(np.unique(want[..., 0])) >= 2           # a lit, non-constant image


# ---- the same known answers through libswrb.so (GPU) ---------------------------------------------------------------
def _gpu_frame(rast, size, meshlets, m, mats, textures, lights=None):
    gscene = rast.upload_scene(meshlets, mats, textures, lights)
    fb = rast.create_framebuffer(size, size)
    fb.clear(0xFF000000, 0.0)
    rast.reset_counters()
    rast.draw_meshlets(fb, gscene, 0, len(meshlets), m)
    assert rast.counters()["TrianglesRasterized"] == 2
    return fb, gscene


@pytest.mark.gpu
def test_gpu_one_to_one_textured_quad_and_normals(rast_factory):
    size = 64
    tex = gradient_texture(size)
    meshlets, m, mats, textures = quad_scene(size, tex=tex)
    rast = rast_factory()
    fb, gscene = _gpu_frame(rast, size, meshlets, m, mats, textures)
    rast.resolve_debug(fb, gscene, "BaseColor", **uniforms(m, size))
    assert np.array_equal(fb.get_pixels(0), tx.get_pixels(tex, 0, 0))
    fb, gscene = _gpu_frame(rast, size, meshlets, m, mats, textures)
    rast.resolve_debug(fb, gscene, "Normals", **uniforms(m, size))
    px = fb.get_pixels(0)
    got = np.stack([(px >> (8 * c)) & 255 for c in range(3)], -1).astype(np.int64)
    assert np.abs(got - np.array([128, 128, 255])).max() <= 1


@pytest.mark.gpu
@pytest.mark.parametrize("metal,rough", [(0, 128), (255, 64)])
def test_gpu_directional_light_matches_float64_brdf(orc, rast_factory, metal, rough):
    """The CUDA resolve against the float64 model directly (and, as everywhere, against the oracle within 2/255)."""
    size = 64
    tex = tx.create_texture(16, 16, 1, 2)
    albedo = (200, 150, 100)
    tx.set_pixels(tex, np.full((16, 16), albedo[0] | (albedo[1] << 8) | (albedo[2] << 16) | (255 << 24), dtype=np.uint32), 0)
    tx.set_pixels(tex, np.full((16, 16), (metal << 16) | (rough << 24), dtype=np.uint32), 1)
    meshlets, m, mats, textures = quad_scene(size, tex=tex)
    light = np.zeros(1, dtype=LIGHT_DTYPE)
    ldir = np.array([0.3, -0.2, -1.0]); ldir /= np.linalg.norm(ldir)
    light["Type"], light["Direction"], light["Color"], light["Intensity"] = 0, ldir.astype(f32), (1.0, 0.9, 0.8), 900.0
    uni = uniforms(m, size, np.array([0.2, -0.1, 3.0]), 0.7)
    ofb = render(orc, size, meshlets, m, mats, textures)
    orc.resolve(ofb, meshlets, mats, textures, light, **uni)
    want = detile(ofb.data[0, :size * size], size, size)          # (the oracle itself is within 1/255 of the float64 model, CPU test above)
    rast = rast_factory()
    fb, gscene = _gpu_frame(rast, size, meshlets, m, mats, textures, light)
    rast.resolve(fb, gscene, **uni)
    got = fb.get_pixels(0)
    d = np.abs(got.view(np.uint8).astype(np.int32) - want.view(np.uint8).astype(np.int32))
    assert d.max() <= 2


# ---- the alpha-tested fragment program FS_EncodeSurfaceId<true> (Shading.cpp:309-331) --------------------------------
def _alpha_scene(size):
    tex = tx.create_texture(size, size, 1, 1)
    y, x = np.meshgrid(np.arange(size), np.arange(size), indexing="ij")
    alpha = np.where(((x // 2) + (y // 3)) % 2 == 0, 255, 0)                      # 2 x 3 texel blocks, on / off
    tx.set_pixels(tex, (0x00406080 | (alpha.astype(np.uint32) << 24)), 0)
    meshlets, m, mats, textures = quad_scene(size, tex=tex)
    mats["AlphaCutoff"] = 128
    meshlets["AlphaCutoff"] = 128                                                  # Scene.cpp:246: copied from the material
    return meshlets, m, mats, textures, alpha


def test_alpha_test_coverage_equals_the_texture_alpha_mask(orc):
    """1:1 mapping: pixel (x, y) samples texel (x, y) (bilinear with a 1/256 weight of the right / lower neighbour, which
    moves alpha by at most 1 — Texture.h:506-575, SIMD.h:448-450), so a pixel is written iff its texel's alpha >= cutoff."""
    size = 64
    meshlets, m, mats, textures, alpha = _alpha_scene(size)
    fb = orc.Framebuffer(size, size)
    fb.clear(0xFFFFFFFF, 0.0)
    c = orc.draw_meshlets(fb, meshlets, 0, len(meshlets), m, materials=mats, textures=textures)
    assert int(c[1]) == 2
    depth = detile(fb.data[1, :size * size], size, size).view(np.float32)
    ids = detile(fb.data[0, :size * size], size, size)
    assert np.array_equal(depth > 0, alpha == 255)
    assert np.all(depth[alpha == 255] == 0.5) and np.all(ids[alpha == 0] == 0xFFFFFFFF)


@pytest.mark.gpu
def test_gpu_alpha_test_coverage_equals_the_texture_alpha_mask(rast_factory):
    size = 64
    meshlets, m, mats, textures, alpha = _alpha_scene(size)
    for binning in (True, False):
        rast = rast_factory(enable_binning=binning)
        gscene = rast.upload_scene(meshlets, mats, textures)
        fb = rast.create_framebuffer(size, size)
        fb.clear(0xFFFFFFFF, 0.0)
        rast.draw_meshlets(fb, gscene, 0, len(meshlets), m)
        depth = fb.get_pixels(1).view(np.float32)
        assert np.array_equal(depth > 0, alpha == 255) and np.all(depth[alpha == 255] == 0.5)
        assert np.all(fb.get_pixels(0)[alpha == 0] == 0xFFFFFFFF)
