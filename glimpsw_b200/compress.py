"""Meshlet compression as an import-time / transport format (the reference author's TODO at Shading.cpp:292-294).

`pack_meshlets` quantizes the positions of every meshlet to 16 bits inside the meshlet's own bounding box and keeps every other
byte (bounds, cone, counts, material, UVs, normals / tangents, indices): 1376 instead of 1728 bytes per meshlet. The library
decodes on the device at upload (`Rasterizer.upload_scene_packed` -> swrb_scene_create_packed -> k_unpack_meshlets) with
    position = fmaf(float(q), Scale, Origin)
which the test oracle restates in C; a packed scene renders bit-identically to its host-decoded meshlets.
The quantization error is at most half a step, (max - min) / 131070 per axis — for a meshlet one metre across, 8 micrometres.
"""
from __future__ import annotations

import numpy as np

from .layout import MESHLET_DTYPE, PACKED_MESHLET_DTYPE

f32 = np.float32


def pack_meshlets(meshlets: np.ndarray) -> np.ndarray:
    assert meshlets.dtype.itemsize == MESHLET_DTYPE.itemsize
    meshlets = np.ascontiguousarray(meshlets)
    n = len(meshlets)
    out = np.zeros(n, dtype=PACKED_MESHLET_DTYPE)
    raw = meshlets.view(np.uint8).reshape(n, MESHLET_DTYPE.itemsize)
    out["Header"] = raw[:, :64]
    pos = meshlets["Positions"]                                           # [n, 3, 64]
    valid = (np.arange(64)[None, :] < meshlets["NumVertices"].astype(np.int64)[:, None])[:, None, :]
    lo = np.where(valid, pos, np.inf).min(axis=2)
    hi = np.where(valid, pos, -np.inf).max(axis=2)
    empty = ~np.isfinite(lo)
    lo = np.where(empty, 0, lo).astype(f32)
    hi = np.where(empty, 0, hi).astype(f32)
    scale = ((hi.astype(np.float64) - lo.astype(np.float64)) / 65535.0).astype(f32)
    # the largest code must not decode beyond what the float32 scale reaches; q is chosen against the decode it will get
    safe = np.where(scale > 0, scale, 1).astype(np.float64)
    q = np.rint((np.where(valid, pos, lo[:, :, None]).astype(np.float64) - lo[:, :, None].astype(np.float64)) / safe[:, :, None])
    out["Q"] = np.clip(q, 0, 65535).astype(np.uint16)
    out["Origin"], out["Scale"] = lo, scale
    out["TexCoords"], out["NormalTangents"], out["Indices"] = meshlets["TexCoords"], meshlets["NormalTangents"], meshlets["Indices"]
    return out


def unpack_meshlets(packed: np.ndarray) -> np.ndarray:
    """The decode in numpy (float32 fused multiply-add emulated in float64: q * scale is exact in 53 bits for 16-bit q and a
    24-bit scale, and the sum with the 24-bit origin is rounded once to float32 — the same result as fmaf)."""
    n = len(packed)
    out = np.zeros(n, dtype=MESHLET_DTYPE)
    out.view(np.uint8).reshape(n, MESHLET_DTYPE.itemsize)[:, :64] = packed["Header"]
    prod = packed["Q"].astype(np.float64) * packed["Scale"].astype(np.float64)[:, :, None]
    out["Positions"] = (prod + packed["Origin"].astype(np.float64)[:, :, None]).astype(f32)
    out["TexCoords"], out["NormalTangents"], out["Indices"] = packed["TexCoords"], packed["NormalTangents"], packed["Indices"]
    return out
