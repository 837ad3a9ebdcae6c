"""ctypes binding of oracle/_ref/libswr_ref.so — the REFERENCE's own Rasterizer.cpp / Shading.cpp / ImageHelpers.cpp compiled
by oracle/ref_build.py (test infrastructure; see ref_api.cpp). Same call shapes as oracle/orc.py so a test can run one scene
through the restatement and through the reference's code and compare the framebuffers word for word.

`available()` is False on a machine that has neither /root/reference nor a prebuilt library, or whose CPU lacks the AVX-512
subsets the reference itself requires; tests skip in that case.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import ref_build
from .orc import Framebuffer, _mat, _p, _texture_descs  # noqa: F401  (Framebuffer re-exported: same host mirror)

_LIBS: dict = {}


def available() -> bool:
    if not ref_build.cpu_ok():
        return False
    return ref_build.available() or os.path.exists(ref_build.LIB)


def lib(fast: bool = False) -> C.CDLL:
    """fast=False: canonical arithmetic (IEEE, no contraction, approx_rcp = 1/x) — what parity is defined against.
    fast=True : -Ofast -mrecip with the real vrcp14ps / vrsqrt14ps — what the upstream flags license (sensitivity only)."""
    if fast not in _LIBS:
        ref_build.build()
        path = ref_build.LIB_FAST if fast else ref_build.LIB
        l = C.CDLL(path)
        l.ref_build_info.restype = C.c_char_p
        l.ref_rasterizer_create.restype = C.c_void_p
        l.ref_cull_meshlets.restype = C.c_uint32
        _LIBS[fast] = l
    return _LIBS[fast]


def build_info(fast: bool = False) -> str:
    return lib(fast).ref_build_info().decode()


class Rasterizer:
    """swr::Rasterizer (Rasterizer.h:198-248). threads=1 is the deterministic one-worker order the restatement follows."""

    def __init__(self, threads: int = 1, fast: bool = False):
        self._lib = lib(fast)
        self._h = C.c_void_p(self._lib.ref_rasterizer_create(C.c_uint32(threads)))
        self.threads = int(self._lib.ref_rasterizer_threads(self._h))

    def close(self):
        if self._h:
            self._lib.ref_rasterizer_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def draw_meshlets(self, fb: Framebuffer, meshlets: np.ndarray, meshlet_offset: int, count: int, object_to_clip,
                      cull_bitmap=None, materials=None, guardband: bool = True, counters=None, textures=None,
                      binned: bool = True, clipping: bool = False, overdraw: bool = False, deferred: bool = False,
                      object_to_world3=None) -> np.ndarray:
        """Rasterizer::DrawMeshlets with ShadingContext::VisBufferShader (OverdrawShader / DeferredShader on request)."""
        assert meshlets.dtype.itemsize == 1728
        if counters is None:
            counters = np.zeros(4, dtype=np.uint64)
        cb = None if cull_bitmap is None else _p(np.ascontiguousarray(cull_bitmap, dtype=np.uint16))
        nmat = 0 if materials is None else len(materials)
        descs, keep = (None, None) if not textures else _texture_descs(textures)
        ntex = 0 if not textures else len(textures)
        o2w = None if object_to_world3 is None else _p(np.ascontiguousarray(np.asarray(object_to_world3, dtype=np.float32).reshape(9)))
        flags = (1 if guardband else 0) | (2 if clipping else 0) | (0 if binned else 4) | ((1 if overdraw else (2 if deferred else 0)) << 4)
        rc = self._lib.ref_draw_meshlets(self._h, _p(fb.data), C.c_uint32(fb.width), C.c_uint32(fb.height), C.c_uint32(fb.layers),
                                         _p(meshlets), C.c_uint32(meshlet_offset), C.c_uint32(count), _p(_mat(object_to_clip)), o2w, cb,
                                         _p(materials) if nmat else None, C.c_uint32(nmat), descs, C.c_uint32(ntex),
                                         C.c_uint32(flags), _p(counters))
        assert rc == 0, "texture descriptor does not match CreateTexture2D's layout"
        return counters

    def resolve(self, fb: Framebuffer, meshlets: np.ndarray, materials, textures, lights, object_to_clip, object_to_world3,
                inv_screen_proj, view_pos, exposure: float = 1.0, world_to_clip=None, skybox=None, meshlet_offset: int = 0,
                debug_layer: int = 0, derive_inverse: bool = False, **_unused):
        """ShadingContext::Resolve (or ResolveDebug with `debug_layer`). Resolve always draws the markers of point / spot lights
        with WorldToClipMat (= `world_to_clip`, else `object_to_clip`); orc.resolve only does when given `world_to_clip`. derive_inverse=True lets Resolve invert WorldToClipMat itself (GLM stand-in)."""
        descs, keep = _texture_descs(textures)
        vp = np.ascontiguousarray(np.asarray(view_pos, dtype=np.float32))
        o2w = np.ascontiguousarray(np.asarray(object_to_world3, dtype=np.float32).reshape(9))
        lights = np.ascontiguousarray(lights)
        sky_desc, keep_sky = (None, None) if skybox is None else _texture_descs([skybox])
        w2c = _mat(world_to_clip if world_to_clip is not None else object_to_clip)
        inv = None if derive_inverse else _p(_mat(inv_screen_proj))
        rc = self._lib.ref_resolve(self._h, _p(fb.data), C.c_uint32(fb.width), C.c_uint32(fb.height), C.c_uint32(fb.layers), _p(meshlets),
                                   C.c_uint32(meshlet_offset), _p(materials) if len(materials) else None, C.c_uint32(len(materials)),
                                   descs, C.c_uint32(len(textures)), _p(lights) if len(lights) else None, C.c_uint32(len(lights)),
                                   _p(_mat(object_to_clip)), _p(o2w), _p(w2c), inv, _p(vp), C.c_float(exposure), sky_desc,
                                   C.c_int(debug_layer))
        assert rc == 0, "texture descriptor does not match CreateTexture2D's layout"


def cull_meshlets(meshlets: np.ndarray, proj, view, model, prev_view, frame_w: int, frame_h: int, pyramid=None, fast: bool = False):
    """ShadingContext::CullMeshlets (pyramid=None: frustum part only). Returns (bitmap, visible)."""
    bitmap = np.zeros((len(meshlets) + 15) // 16, dtype=np.uint16)
    descs, keep = (None, None) if pyramid is None else _texture_descs([pyramid])
    n = lib(fast).ref_cull_meshlets(_p(bitmap), _p(meshlets), C.c_uint32(len(meshlets)), _p(_mat(proj)), _p(_mat(view)), _p(_mat(model)),
                                    _p(_mat(prev_view)), C.c_float(frame_w), C.c_float(frame_h), descs)
    assert n != 0xFFFFFFFF, "pyramid descriptor does not match CreateTexture2D's layout"
    return bitmap, int(n)


def downsample_depth(fb: Framebuffer, pyramid) -> None:
    """texutil::DownsampleDepth: fb layer 1 -> `pyramid` (glimpsw_b200.textures.TextureData, R32f bits)."""
    descs, keep = _texture_descs([pyramid])
    data = keep[0].copy()
    rc = lib().ref_downsample_depth(_p(fb.data[1]), C.c_uint32(fb.width), C.c_uint32(fb.height), descs, _p(data))
    assert rc == 0
    pyramid.data = data


def fb_clear(fb: Framebuffer, color: int, depth: float) -> None:
    lib().ref_fb_clear(_p(fb.data), C.c_uint32(fb.width), C.c_uint32(fb.height), C.c_uint32(fb.layers), C.c_uint32(color), C.c_float(depth))


def fb_get_pixels(fb: Framebuffer, layer: int) -> np.ndarray:
    out = np.zeros((fb.height, fb.width), dtype=np.uint32)
    lib().ref_fb_get_pixels(_p(fb.data[layer]), C.c_uint32(fb.width), C.c_uint32(fb.height), _p(out), C.c_uint32(fb.width))
    return out


def generate_mips(tex) -> np.ndarray:
    """Texture2D::GenerateMips on a copy of the texture's data; returns the full texel array."""
    descs, keep = _texture_descs([tex])
    data = keep[0].copy()
    rc = lib().ref_generate_mips(_p(data), descs)
    assert rc == 0
    return data


def sample_implicit_lod_4x4(tex, u, v) -> np.ndarray:
    descs, keep = _texture_descs([tex])
    uu = np.ascontiguousarray(u, dtype=np.float32).reshape(16)
    vv = np.ascontiguousarray(v, dtype=np.float32).reshape(16)
    out = np.zeros(16, dtype=np.uint32)
    rc = lib().ref_sample_implicit_lod_4x4(descs, _p(uu), _p(vv), _p(out))
    assert rc == 0
    return out


def probe_triangle(verts, width: int, height: int, cull_mode: int = 1, guardband: bool = True, fast: bool = False):
    v = np.ascontiguousarray(np.asarray(verts, dtype=np.float32).reshape(12))
    cc = C.c_int(0)
    pos = np.zeros(3, dtype=np.uint32)
    bbox = np.zeros(2, dtype=np.uint32)
    edges = np.zeros(9, dtype=np.int32)
    zw = np.zeros(7, dtype=np.float32)
    keep = lib(fast).ref_probe_triangle(_p(v), width, height, cull_mode, 1 if guardband else 0, C.byref(cc), _p(pos), _p(bbox), _p(edges), _p(zw))
    return dict(keep=int(keep), cc=cc.value, pos=pos, bbox=bbox, edges=edges, zw=zw)


def octahedron_from_panorama(pano_tex, cube_tex) -> np.ndarray:
    """LoadOctahedronFromPanoramaHDR on an in-memory panorama texture (glimpsw_b200.textures.TextureData, R11G11B10f); returns the
    texel array of the octahedron map laid out like `cube_tex`."""
    pd, keep_p = _texture_descs([pano_tex])
    cd, keep_c = _texture_descs([cube_tex])
    out = np.zeros_like(keep_c[0])
    rc = lib().ref_octahedron_from_panorama(pd, cd, _p(out))
    assert rc == 0
    return out
