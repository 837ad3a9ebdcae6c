"""Byte layouts at the drop-in boundary as numpy dtypes (mirror of include/swr_types.h).

Reference: src/SwRast/Scene.h:15-30 (Meshlet), :7-14 (Material), :52-75 (Light),
src/SwRast/Rasterizer.h:10-78 (Framebuffer, 4x4-tiled layers).
"""
import numpy as np

MAX_VERTICES = 64
MAX_PRIMS = 128
MAX_RENDER_SIZE = 2896
NO_MATERIAL = 0xFFFFFFFF

MESHLET_DTYPE = np.dtype({
    "names": ["BoundCenter", "BoundRadius", "ConeApex", "ConeAxis", "ConeCutoff",
              "NumVertices", "NumTriangles", "AlphaCutoff", "MaterialId", "TangentHandedness",
              "Positions", "TexCoords", "NormalTangents", "Indices"],
    "formats": [("<f4", 3), "<f4", ("<f4", 3), ("<f4", 3), "<f4",
                "u1", "u1", "u1", "<u4", "<u8",
                ("<f4", (3, 64)), ("<u4", 64), ("<u4", 64), ("u1", (3, 128))],
    "offsets": [0, 12, 16, 28, 40, 44, 45, 46, 48, 56, 64, 832, 1088, 1344],
    "itemsize": 1728,
})
assert MESHLET_DTYPE.itemsize == 1728

# swr_meshlet_packed (include/swr_types.h): the Meshlet with 16-bit positions inside its own bounding box, 1376 bytes
PACKED_MESHLET_DTYPE = np.dtype({
    "names": ["Header", "Origin", "Scale", "Q", "TexCoords", "NormalTangents", "Indices"],
    "formats": [("u1", 64), ("<f4", 3), ("<f4", 3), ("<u2", (3, 64)), ("<u4", 64), ("<u4", 64), ("u1", (3, 128))],
    "offsets": [0, 64, 76, 96, 480, 736, 992],
    "itemsize": 1376,
})
assert PACKED_MESHLET_DTYPE.itemsize == 1376

MATERIAL_DTYPE = np.dtype([("TextureId", "<i4"), ("IsDoubleSided", "u1"), ("AlphaCutoff", "u1"), ("_pad", "u1", 2)])
assert MATERIAL_DTYPE.itemsize == 8

LIGHT_DTYPE = np.dtype([("Type", "<u4"), ("Position", "<f4", 3), ("Direction", "<f4", 3), ("Color", "<f4", 3),
                        ("Intensity", "<f4"), ("Radius", "<f4"), ("SpotInnerAngle", "<f4"), ("SpotOuterAngle", "<f4"),
                        ("InvRadiusSq", "<f4"), ("SpotScale", "<f4"), ("SpotOffset", "<f4")])
assert LIGHT_DTYPE.itemsize == 68


def fb_layer_stride(width: int, height: int) -> int:
    """CreateFramebuffer (Rasterizer.h:69): layer stride in u32, padded to 64."""
    return (width * height + 63) & ~63


def fb_pixel_offsets(width: int, height: int) -> np.ndarray:
    """Framebuffer::GetPixelOffset (Rasterizer.h:50-56) for every (y, x) -> [H, W] int64."""
    x = np.arange(width, dtype=np.int64)[None, :]
    y = np.arange(height, dtype=np.int64)[:, None]
    return ((x & ~3) << 2) + (y & ~3) * width + (x & 3) + (y & 3) * 4


def detile(layer: np.ndarray, width: int, height: int) -> np.ndarray:
    """Host-side equivalent of Framebuffer::GetPixels for checking; returns [H, W]."""
    return layer[fb_pixel_offsets(width, height)]
