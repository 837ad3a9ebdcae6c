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
