"""Host-side RGBA8 textures in the reference's storage format (src/SwRast/Texture.h).

Texture2D<RGBA8u, TiledY8>: power-of-two size, 8x8-texel Y-major tiles, a mip chain padded to 64 texels
per level and `NumLayers` layers (layer 0 base colour, layer 1 normal.xy + metallic + roughness,
Scene.h:8-11). Creation mirrors CreateTexture2D (Texture.h:600-636), SetPixels (:331-349) and
GenerateMips/GenerateMip (:378-384, :577-596).
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

f32 = np.float32


@dataclass
class TextureData:
    width: int
    height: int
    mip_levels: int
    num_layers: int
    row_shift: int
    layer_stride: int
    mip_offsets: np.ndarray      # [16] u32, in texels
    data: np.ndarray             # [layer_stride * num_layers + 64] u32


def texel_offset(x, y, stride):
    """Texture2D::GetTexelOffset for TiledY8 (Texture.h:494-501)."""
    return (y & 7) | (x << 3) | ((y & ~7) << stride)


def create_texture(width: int, height: int, max_levels: int = 8, num_layers: int = 1) -> TextureData:
    assert width & (width - 1) == 0 and height & (height - 1) == 0 and max_levels <= 16
    row_shift = int(max(width, 8)).bit_length() - 1
    mip_offsets = np.zeros(16, dtype=np.uint32)
    layer_stride = 0
    mip = 0
    while mip < max_levels:
        if (width >> mip) < 4 or (height >> mip) < 4:
            break
        mip_offsets[mip] = layer_stride
        layer_stride += ((width >> mip) * (height >> mip) + 63) & ~63
        mip += 1
    data = np.zeros(layer_stride * num_layers + 64, dtype=np.uint32)
    return TextureData(width, height, mip, num_layers, row_shift, layer_stride, mip_offsets, data)


def set_pixels(tex: TextureData, pixels: np.ndarray, layer: int = 0, level: int = 0):
    h, w = tex.height >> level, tex.width >> level
    assert pixels.shape == (h, w)
    y, x = np.meshgrid(np.arange(h), np.arange(w), indexing="ij")
    off = layer * tex.layer_stride + int(tex.mip_offsets[level]) + texel_offset(x, y, tex.row_shift - level)
    tex.data[off.reshape(-1)] = pixels.astype(np.uint32).reshape(-1)


def get_pixels(tex: TextureData, layer: int = 0, level: int = 0) -> np.ndarray:
    h, w = tex.height >> level, tex.width >> level
    y, x = np.meshgrid(np.arange(h), np.arange(w), indexing="ij")
    off = layer * tex.layer_stride + int(tex.mip_offsets[level]) + texel_offset(x, y, tex.row_shift - level)
    return tex.data[off]


def generate_mips(tex: TextureData):
    """2x2 box filter in float32, (((a+b)+c)+d)*0.25, RNE pack (Texture.h:577-596, :55-67)."""
    s = f32(1.0) / f32(255.0)
    for layer in range(tex.num_layers):
        for level in range(1, tex.mip_levels):
            src = get_pixels(tex, layer, level - 1)
            out = np.zeros((tex.height >> level, tex.width >> level), dtype=np.uint32)
            for k in range(4):
                ch = ((src >> (8 * k)) & 255).astype(f32) * s
                a, b, c, d = ch[0::2, 0::2], ch[0::2, 1::2], ch[1::2, 0::2], ch[1::2, 1::2]
                avg = (((a + b) + c) + d) * f32(0.25)
                q = np.clip(np.rint(avg * f32(255.0)), 0, 255).astype(np.uint32)
                out |= q << (8 * k)
            set_pixels(tex, out, layer, level)


def pack_rgba(r, g, b, a) -> np.ndarray:
    q = [np.clip(np.rint(np.asarray(c, dtype=np.float64) * 255.0), 0, 255).astype(np.uint32) for c in (r, g, b, a)]
    return q[0] | (q[1] << 8) | (q[2] << 16) | (q[3] << 24)


def procedural_material_texture(size: int = 1024, seed: int = 0, with_nmr: bool = True, alpha_holes: bool = False,
                                max_levels: int = 8) -> TextureData:
    """A two-layer material texture: base colour (checker + stripes + value noise) and a normal/metal/rough layer."""
    tex = create_texture(size, size, max_levels, 2 if with_nmr else 1)
    v, u = np.meshgrid((np.arange(size) + 0.5) / size, (np.arange(size) + 0.5) / size, indexing="ij")
    rng = np.random.default_rng(seed)
    phase = rng.uniform(0, 6.28, 6)
    checker = ((np.floor(u * 16) + np.floor(v * 16)) % 2)
    r = 0.25 + 0.5 * checker + 0.2 * np.sin(u * 40 + phase[0])
    g = 0.35 + 0.4 * np.sin(v * 25 + phase[1]) ** 2 + 0.1 * checker
    b = 0.30 + 0.5 * np.cos((u + v) * 18 + phase[2]) ** 2
    noise = rng.uniform(-0.06, 0.06, (size, size))
    a = np.ones_like(u)
    if alpha_holes:
        a = ((np.sin(u * 50) * np.sin(v * 50)) > -0.3).astype(np.float64)
    set_pixels(tex, pack_rgba(np.clip(r + noise, 0, 1), np.clip(g + noise, 0, 1), np.clip(b + noise, 0, 1), a), 0)
    if with_nmr:
        nx = 0.5 + 0.22 * np.sin(u * 64 + phase[3])
        ny = 0.5 + 0.22 * np.cos(v * 64 + phase[4])
        metal = (checker * 0.8 + 0.1)
        rough = 0.25 + 0.5 * (0.5 + 0.5 * np.sin((u - v) * 12 + phase[5]))
        set_pixels(tex, pack_rgba(nx, ny, metal, rough), 1)
    generate_mips(tex)
    return tex


# ---- HdrTexture2D = Texture2D<R11G11B10f, TiledY8> (Texture.h:133-200, :216): the skybox of ShadingContext::Resolve ----
def pack_r11g11b10f(r, g, b) -> np.ndarray:
    """pixfmt::R11G11B10f::Pack (Texture.h:145-167): clamp to [2^-15, max], truncate the mantissa."""
    def f11(x):
        x = np.minimum(np.maximum(np.asarray(x, dtype=f32), f32(1.0 / (1 << 15))), f32(130048.0))
        return ((x.view(np.uint32) >> 17) & 0x3FFF) - 0x1C00
    def f10(x):
        x = np.minimum(np.maximum(np.asarray(x, dtype=f32), f32(1.0 / (1 << 15))), f32(129024.0))
        return ((x.view(np.uint32) >> 18) & 0x1FFF) - 0x0E00
    return (f11(r).astype(np.uint32) << 21) | (f11(g).astype(np.uint32) << 10) | f10(b).astype(np.uint32)


def unpack_r11g11b10f(p: np.ndarray):
    """pixfmt::R11G11B10f::Unpack (Texture.h:138-144, :170-182)."""
    p = np.asarray(p, dtype=np.uint32)
    r = ((((p >> 21) << 17) & 0x0FFE0000) + 0x38000000).astype(np.uint32).view(f32)
    g = ((((p >> 10) << 17) & 0x0FFE0000) + 0x38000000).astype(np.uint32).view(f32)
    b = (((p << 18) & 0x0FFC0000) + 0x38000000).astype(np.uint32).view(f32)
    return r, g, b


def generate_mips_hdr(tex: TextureData):
    """Texture2D<R11G11B10f>::GenerateMip (Texture.h:577-596): unpack, (((a+b)+c)+d)*0.25 in float32, pack (truncating)."""
    for level in range(1, tex.mip_levels):
        src = get_pixels(tex, 0, level - 1)
        ch = unpack_r11g11b10f(src)
        avg = [(((c[0::2, 0::2] + c[0::2, 1::2]) + c[1::2, 0::2]) + c[1::2, 1::2]) * f32(0.25) for c in ch]
        set_pixels(tex, pack_r11g11b10f(*avg), 0, level)


def unmap_octahedron(u: np.ndarray, v: np.ndarray):
    """texutil::UnmapOctahedron (Texture.h:289-296), float64: texel centre -> direction."""
    x, y = u * 2.0 - 1.0, v * 2.0 - 1.0
    z = 1.0 - np.abs(x) - np.abs(y)
    t = np.maximum(-z, 0.0)
    x, y = x - np.copysign(t, x), y - np.copysign(t, y)
    n = np.sqrt(x * x + y * y + z * z)
    return x / n, y / n, z / n


def hdr_texture_from_pixels(rgb: np.ndarray, max_levels: int = 8) -> TextureData:
    """texutil::LoadImageHDR (ImageHelpers.cpp:46-71) minus the file decoder: float RGB pixels [H, W, 3] -> HdrTexture2D
    (R11G11B10f, truncating pack) with its mip chain."""
    h, w = rgb.shape[:2]
    tex = create_texture(w, h, max_levels, 1)
    rgb = np.asarray(rgb, dtype=f32)
    set_pixels(tex, pack_r11g11b10f(rgb[..., 0], rgb[..., 1], rgb[..., 2]), 0)
    generate_mips_hdr(tex)
    return tex


def _sample_linear_hdr_level0(tex: TextureData, u: np.ndarray, v: np.ndarray):
    """Texture2D<R11G11B10f>::SampleLevel<{Repeat, Linear, Linear}>(u, v, 0, 0) (Texture.h:412-459) with the float branch of
    SampleLinear (:506-575): 8 fractional bits, the +1 texel not wrapped but faded out (fx = 0) at the right edge, the row
    below only taken when it is inside the image; float32 in the source's operation order."""
    W, H = tex.width, tex.height
    su, sv = u.astype(f32) * f32(W << 8), v.astype(f32) * f32(H << 8)
    ix = np.rint(su).astype(np.int64) & ((W << 8) - 1)
    iy = np.rint(sv).astype(np.int64) & ((H << 8) - 1)
    ixf, iyf = np.maximum(ix - 127, 0), np.maximum(iy - 127, 0)
    ix, iy = ixf >> 8, iyf >> 8
    inx, iny = (ix + 1) < W, (iy + 1) < H
    data = np.asarray(tex.data, dtype=np.uint32)
    o00 = texel_offset(ix, iy, tex.row_shift)
    o01 = texel_offset(ix, iy + iny.astype(np.int64), tex.row_shift)
    lim = len(data) - 1
    c00, c10 = unpack_r11g11b10f(data[o00]), unpack_r11g11b10f(data[np.minimum(o00 + 8, lim)])
    c01, c11 = unpack_r11g11b10f(data[o01]), unpack_r11g11b10f(data[np.minimum(o01 + 8, lim)])
    fx = np.where(inx, (ixf & 255).astype(f32) * f32(1.0 / 256), f32(0))
    fy = (iyf & 255).astype(f32) * f32(1.0 / 256)
    out = []
    for k in range(3):
        row_a = c00[k] + (c10[k] - c00[k]) * fx
        row_b = c01[k] + (c11[k] - c01[k]) * fx
        out.append((row_a + (row_b - row_a) * fy).astype(f32))
    return out


def octahedron_from_panorama(pano_rgb: np.ndarray, max_levels: int = 8) -> TextureData:
    """texutil::LoadOctahedronFromPanoramaHDR (ImageHelpers.cpp:73-104) minus the file decoder: an equirectangular float RGB
    panorama [H, W, 3] (power-of-two sizes) -> the octahedron-mapped HdrTexture2D (W x W) ShadingContext::SkyboxTex expects.
    Per texel: u, v = x / (W - 1) + 0.5 / (W - 1); direction = UnmapOctahedron(u, v); panorama coordinates
    atan2(dir.z, dir.x) / tau + 0.5, asin(-dir.y) / pi + 0.5; one bilinear sample of the panorama's level 0; truncating pack."""
    pano = hdr_texture_from_pixels(pano_rgb, 1)
    face = pano.width
    cube = create_texture(face, face, max_levels, 1)
    scale = f32(1.0) / f32(face - 1)
    center = f32(0.5) * scale
    coord = (np.arange(face, dtype=np.int64).astype(f32) * scale + center).astype(f32)
    v, u = np.meshgrid(coord, coord, indexing="ij")
    # UnmapOctahedron in float32, canonical approx_rsqrt = 1 / sqrt (Texture.h:289-296, SIMD.h:439)
    x, y = u * f32(2) - f32(1), v * f32(2) - f32(1)
    z = (f32(1) - np.abs(x) - np.abs(y)).astype(f32)
    t = np.maximum(-z, f32(0))
    x, y = (x - np.copysign(t, x)).astype(f32), (y - np.copysign(t, y)).astype(f32)
    dot = (x * x + (y * y + z * z)).astype(f32)                  # fma chain of simd::dot; the difference is below the 6-bit mantissa
    r = (f32(1) / np.sqrt(dot)).astype(f32)
    x, y, z = x * r, y * r, z * r
    pu = (np.arctan2(z, x).astype(f32) / f32(6.283185307179586) + f32(0.5)).astype(f32)
    pv = (np.arcsin(np.clip(-y, -1, 1)).astype(f32) / f32(3.141592653589793) + f32(0.5)).astype(f32)
    rgb = _sample_linear_hdr_level0(pano, pu, pv)
    set_pixels(cube, pack_r11g11b10f(*rgb), 0)
    generate_mips_hdr(cube)
    return cube


def procedural_sky_texture(size: int = 256, sun_dir=(0.4, 0.7, 0.6), max_levels: int = 6) -> TextureData:
    """An octahedron-mapped HDR sky (stand-in for the reference's panorama -> octahedron import, the HDR files are not in
    the checkout): horizon-to-zenith gradient, a ground tint and an HDR sun lobe well above 1.0."""
    tex = create_texture(size, size, max_levels, 1)
    v, u = np.meshgrid((np.arange(size) + 0.5) / size, (np.arange(size) + 0.5) / size, indexing="ij")
    x, y, z = unmap_octahedron(u, v)
    s = np.asarray(sun_dir, dtype=np.float64)
    s = s / np.linalg.norm(s)
    up = np.clip(y, 0.0, 1.0)
    sun = np.clip(x * s[0] + y * s[1] + z * s[2], 0.0, 1.0) ** 64
    ground = (y < 0).astype(np.float64)
    r = 0.55 - 0.35 * up + 12.0 * sun - 0.35 * ground
    g = 0.70 - 0.25 * up + 10.0 * sun - 0.45 * ground
    b = 0.95 - 0.10 * up + 7.0 * sun - 0.70 * ground
    set_pixels(tex, pack_r11g11b10f(np.maximum(r, 0.02), np.maximum(g, 0.02), np.maximum(b, 0.02)), 0)
    generate_mips_hdr(tex)
    return tex
