"""ctypes binding of oracle/liboracle.so (test infrastructure; see oracle.cpp header)."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build(force: bool = False) -> str:
    path = os.path.join(_HERE, "liboracle.so")
    srcs = [os.path.join(_HERE, f) for f in os.listdir(_HERE) if f.endswith(".cpp")]
    stale = (not os.path.exists(path)) or any(os.path.getmtime(s) > os.path.getmtime(path) for s in srcs)
    if force or stale:
        subprocess.run(["make", "-C", _HERE, "-B" if force else "-s", "liboracle.so"], check=True,
                       stdout=subprocess.DEVNULL)
    return path


def lib() -> C.CDLL:
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "liboracle.so")
        if not os.path.exists(path):
            build()
        _LIB = C.CDLL(path)
        _LIB.orc_build_info.restype = C.c_char_p
        _LIB.orc_cull_meshlets.restype = C.c_uint32
    return _LIB


def _p(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)


def _mat(m) -> np.ndarray:
    return np.ascontiguousarray(np.asarray(m, dtype=np.float32).reshape(16))


class Framebuffer:
    """Host mirror of swr::Framebuffer (Rasterizer.h:10-78): [layers, LayerStride] u32, 4x4-tiled."""

    def __init__(self, width: int, height: int, layers: int = 2):
        assert width % 4 == 0 and height % 4 == 0
        self.width, self.height, self.layers = width, height, layers
        self.layer_stride = (width * height + 63) & ~63
        self.data = np.zeros((layers, self.layer_stride), dtype=np.uint32)

    def clear(self, color: int, depth: float):
        self.data[0, :] = np.uint32(color)
        self.data[1, :] = np.float32(depth).view(np.uint32)

    def get_pixels(self, layer: int) -> np.ndarray:
        out = np.zeros((self.height, self.width), dtype=np.uint32)
        lib().orc_fb_get_pixels(_p(self.data[layer]), self.width, self.height, _p(out), self.width)
        return out


def draw_meshlets(fb: Framebuffer, meshlets: np.ndarray, meshlet_offset: int, count: int, object_to_clip,
                  cull_bitmap=None, materials=None, guardband: bool = True, counters=None, textures=None,
                  binned: bool = True, clipping: bool = False, overdraw: bool = False, deferred: bool = False,
                  object_to_world3=None) -> np.ndarray:
    """Rasterizer::DrawMeshlets + VisBufferShader (or, with `overdraw`, OverdrawShader = FS_Overdraw, Shading.cpp:333-342,
    :656) on the CPU. Returns the 4 integer perf counters.

    With `textures`, alpha-tested materials (AlphaCutoff < 255) run FS_EncodeSurfaceId<true>; without, every
    triangle takes the opaque program. binned=False selects DrawMeshletsST's treatment of non-trivial triangles:
    clipped and drawn with `clipping`, else dropped without being counted. `deferred`: DeferredShader (G-buffer)."""
    assert meshlets.dtype.itemsize == 1728
    if counters is None:
        counters = np.zeros(4, dtype=np.uint64)
    m = _mat(object_to_clip)
    cb = None if cull_bitmap is None else _p(np.ascontiguousarray(cull_bitmap, dtype=np.uint16))
    mats = None if materials is None or len(materials) == 0 else _p(materials)
    descs, keep = (None, None) if not textures else _texture_descs(textures)
    if deferred:    # ShadingContext::DeferredShader = FS_EncodeGBuffer (Shading.cpp:344-414, :655) into a 3-layer framebuffer
        assert fb.layers >= 3 and not overdraw
        o2w = np.ascontiguousarray(np.asarray(object_to_world3 if object_to_world3 is not None else np.eye(3), dtype=np.float32).reshape(9))
        lib().orc_draw_meshlets_gbuffer(_p(fb.data[0]), _p(fb.data[1]), _p(fb.data[2]), fb.width, fb.height, _p(meshlets),
                                        C.c_uint32(meshlet_offset), C.c_uint32(count), _p(m), _p(o2w), cb, mats, descs,
                                        C.c_uint32((1 if guardband else 0) | (0 if binned else (2 if clipping else 4))), _p(counters))
        return counters
    lib().orc_draw_meshlets_ex(_p(fb.data[0]), _p(fb.data[1]), fb.width, fb.height, _p(meshlets),
                               C.c_uint32(meshlet_offset), C.c_uint32(count), _p(m), cb, mats, descs,
                               C.c_uint32((1 if guardband else 0) | (0 if binned else (2 if clipping else 4)) | (8 if overdraw else 0)),
                               _p(counters))
    return counters


class _TextureDesc(C.Structure):   # swr_texture_desc (include/swr_types.h)
    _fields_ = [("Width", C.c_uint32), ("Height", C.c_uint32), ("MipLevels", C.c_uint32), ("NumLayers", C.c_uint32),
                ("RowShift", C.c_uint32), ("LayerStride", C.c_uint32), ("MipOffsets", C.c_uint32 * 16),
                ("Data", C.c_void_p)]


def _texture_descs(textures):
    descs = (_TextureDesc * max(len(textures), 1))()
    keep = []
    for i, t in enumerate(textures):
        data = np.ascontiguousarray(t.data, dtype=np.uint32)
        keep.append(data)
        d = descs[i]
        d.Width, d.Height, d.MipLevels, d.NumLayers = t.width, t.height, t.mip_levels, t.num_layers
        d.RowShift, d.LayerStride = t.row_shift, t.layer_stride
        for k in range(16):
            d.MipOffsets[k] = int(t.mip_offsets[k])
        d.Data = data.ctypes.data
    return descs, keep


def resolve(fb: Framebuffer, meshlets: np.ndarray, materials, textures, lights, object_to_clip, object_to_world3,
            inv_screen_proj, view_pos, exposure: float = 1.0, world_to_clip=None, skybox=None, **_unused):
    """ShadingContext::Resolve on the CPU: layer 0 (surface ids) is overwritten with RGBA8 colour; with
    `world_to_clip`, point/spot lights are then drawn as markers like the tail of Resolve (Shading.cpp:690-731)."""
    assert meshlets.dtype.itemsize == 1728
    descs, keep = _texture_descs(textures)
    vp = np.ascontiguousarray(np.asarray(view_pos, dtype=np.float32))
    o2w = np.ascontiguousarray(np.asarray(object_to_world3, dtype=np.float32).reshape(9))
    lights = np.ascontiguousarray(lights)
    if skybox is not None:      # ShadingContext::SkyboxTex (Shading.h:29): an HdrTexture2D = Texture2D<R11G11B10f>
        sky_desc, keep_sky = _texture_descs([skybox])
        lib().orc_resolve_sky(_p(fb.data[0]), _p(fb.data[1]), C.c_uint32(fb.width), C.c_uint32(fb.height), _p(meshlets),
                              _p(materials) if len(materials) else None, descs, _p(lights) if len(lights) else None,
                              C.c_uint32(len(lights)), _p(_mat(object_to_clip)), _p(o2w), _p(_mat(inv_screen_proj)), _p(vp),
                              C.c_float(exposure), sky_desc)
    else:
        lib().orc_resolve(_p(fb.data[0]), _p(fb.data[1]), C.c_uint32(fb.width), C.c_uint32(fb.height), _p(meshlets),
                          _p(materials) if len(materials) else None, descs, _p(lights) if len(lights) else None,
                          C.c_uint32(len(lights)), _p(_mat(object_to_clip)), _p(o2w), _p(_mat(inv_screen_proj)), _p(vp),
                          C.c_float(exposure))
    if world_to_clip is not None and len(lights):
        draw_light_markers(fb, lights, world_to_clip)


DEBUG_LAYERS = ["None", "BaseColor", "Normals", "MetallicRoughness", "MeshletId", "TriangleId", "OverdrawPixel", "OverdrawQuad"]


def resolve_debug(fb: Framebuffer, meshlets: np.ndarray, materials, textures, layer, object_to_clip, object_to_world3,
                  inv_screen_proj, **_unused):
    """ShadingContext::ResolveDebug (Shading.cpp:734-773) on the CPU; `layer` is a DebugLayer name or value (Shading.h:8)."""
    layer = DEBUG_LAYERS.index(layer) if isinstance(layer, str) else int(layer)
    assert 1 <= layer <= 7
    descs, keep = _texture_descs(textures)
    o2w = np.ascontiguousarray(np.asarray(object_to_world3, dtype=np.float32).reshape(9))
    lib().orc_resolve_debug(_p(fb.data[0]), _p(fb.data[1]), C.c_uint32(fb.width), C.c_uint32(fb.height), _p(meshlets),
                            _p(materials) if len(materials) else None, descs, _p(_mat(object_to_clip)), _p(o2w),
                            _p(_mat(inv_screen_proj)), C.c_int(layer))


def map_octahedron(direction) -> np.ndarray:
    """texutil::MapOctahedron (Texture.h:282-288) of one direction."""
    d = np.ascontiguousarray(np.asarray(direction, dtype=np.float32))
    uv = np.zeros(2, dtype=np.float32)
    lib().orc_map_octahedron(_p(d), _p(uv))
    return uv


def sample_skybox(tex, direction) -> np.ndarray:
    """HdrTexture2D::SampleOctLevel<EnvSampler>(dir, 1) (Texture.h:467-480) of one direction."""
    descs, keep = _texture_descs([tex])
    d = np.ascontiguousarray(np.asarray(direction, dtype=np.float32))
    rgb = np.zeros(3, dtype=np.float32)
    lib().orc_sample_skybox(descs, _p(d), _p(rgb))
    return rgb


def draw_light_markers(fb: Framebuffer, lights, world_to_clip) -> None:
    lights = np.ascontiguousarray(lights)
    lib().orc_draw_light_markers(_p(fb.data[0]), _p(fb.data[1]), C.c_uint32(fb.width), C.c_uint32(fb.height), _p(lights),
                                 C.c_uint32(len(lights)), _p(_mat(world_to_clip)))


def generate_mip(tex, layer: int, level: int):
    """Texture2D::GenerateMip on the CPU oracle (validates glimpsw_b200.textures.generate_mips)."""
    descs, keep = _texture_descs([tex])
    data = keep[0]
    lib().orc_generate_mip(_p(data), descs, C.c_uint32(layer), C.c_uint32(level))
    return data


class Baseline:
    """Threaded AVX-512 restatement of the reference's binned CPU path (baseline_mt.cpp) — the timed CPU baseline."""

    def __init__(self, threads: int = 0):
        lib().orc_mt_create.restype = C.c_void_p
        self._h = C.c_void_p(lib().orc_mt_create(C.c_int(threads)))
        self.threads = int(lib().orc_mt_threads(self._h))
        self.avx512 = bool(lib().orc_mt_uses_avx512(self._h))

    def clear(self, fb: Framebuffer, color: int, depth: float):
        bits = int(np.float32(depth).view(np.uint32))
        lib().orc_mt_clear(self._h, _p(fb.data[0]), _p(fb.data[1]), C.c_uint32(fb.width * fb.height), C.c_uint32(color), C.c_uint32(bits))

    def draw_meshlets(self, fb: Framebuffer, meshlets, meshlet_offset, count, object_to_clip, cull_bitmap=None,
                      materials=None, guardband=True, counters=None):
        if counters is None:
            counters = np.zeros(4, dtype=np.uint64)
        cb = None if cull_bitmap is None else _p(np.ascontiguousarray(cull_bitmap, dtype=np.uint16))
        mats = None if materials is None or len(materials) == 0 else _p(materials)
        lib().orc_mt_draw_meshlets(self._h, _p(fb.data[0]), _p(fb.data[1]), C.c_uint32(fb.width), C.c_uint32(fb.height),
                                   _p(meshlets), C.c_uint32(meshlet_offset), C.c_uint32(count), _p(_mat(object_to_clip)),
                                   cb, mats, C.c_uint32(1 if guardband else 0), _p(counters))
        return counters

    def resolve(self, fb: Framebuffer, meshlets, materials, textures, lights, object_to_clip, object_to_world3,
                inv_screen_proj, view_pos, exposure: float = 1.0, world_to_clip=None, **_unused):
        descs, keep = _texture_descs(textures)
        vp = np.ascontiguousarray(np.asarray(view_pos, dtype=np.float32))
        o2w = np.ascontiguousarray(np.asarray(object_to_world3, dtype=np.float32).reshape(9))
        lights = np.ascontiguousarray(lights)
        lib().orc_mt_resolve(self._h, _p(fb.data[0]), _p(fb.data[1]), C.c_uint32(fb.width), C.c_uint32(fb.height),
                             _p(meshlets), _p(materials) if len(materials) else None, descs,
                             _p(lights) if len(lights) else None, C.c_uint32(len(lights)), _p(_mat(object_to_clip)),
                             _p(o2w), _p(_mat(inv_screen_proj)), _p(vp), C.c_float(exposure))
        if world_to_clip is not None and len(lights):
            draw_light_markers(fb, lights, world_to_clip)

    def close(self):
        if self._h:
            lib().orc_mt_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def hiz_dims(width: int, height: int):
    """Size of the R32f depth pyramid the Playground allocates (Main.cpp:54-56)."""
    half_w = 1 << int((width - 1) // 2).bit_length()
    half_h = 1 << int((height - 1) // 2).bit_length()
    return half_w, half_h


def downsample_depth(fb: Framebuffer, pyramid) -> None:
    """texutil::DownsampleDepth on the CPU: fb layer 1 -> `pyramid` (glimpsw_b200.textures.TextureData, R32f bits)."""
    descs, keep = _texture_descs([pyramid])
    data = keep[0]
    lib().orc_downsample_depth(_p(fb.data[1]), C.c_uint32(fb.width), C.c_uint32(fb.height), descs, _p(data))
    pyramid.data = data


def cull_meshlets_hiz(meshlets: np.ndarray, proj, view, model, prev_view, frame_w: int, frame_h: int, pyramid=None):
    """ShadingContext::CullMeshlets incl. the HiZ test (pyramid=None: frustum part only). Returns (bitmap, visible)."""
    bitmap = np.zeros((len(meshlets) + 15) // 16, dtype=np.uint16)
    descs, data = None, None
    if pyramid is not None:
        descs, keep = _texture_descs([pyramid])
        data = _p(keep[0])
    lib().orc_cull_meshlets_hiz.restype = C.c_uint32
    n = lib().orc_cull_meshlets_hiz(_p(bitmap), _p(meshlets), C.c_uint32(len(meshlets)), _p(_mat(proj)), _p(_mat(view)),
                                    _p(_mat(model)), _p(_mat(prev_view)), C.c_float(frame_w), C.c_float(frame_h), descs, data)
    return bitmap, int(n)


def frustum_planes(proj, view, model) -> np.ndarray:
    out = np.zeros((6, 4), dtype=np.float32)
    lib().orc_frustum_planes(_p(_mat(proj)), _p(_mat(view)), _p(_mat(model)), _p(out))
    return out


def cull_meshlets(meshlets: np.ndarray, planes: np.ndarray):
    bitmap = np.zeros((len(meshlets) + 15) // 16, dtype=np.uint16)
    pl = np.ascontiguousarray(planes[:5], dtype=np.float32)
    n = lib().orc_cull_meshlets(_p(bitmap), _p(meshlets), C.c_uint32(len(meshlets)), _p(pl))
    return bitmap, int(n)


def probe_triangle(verts, width: int, height: int, cull_mode: int = 1, guardband: bool = True):
    """One triangle through classify + early setup + bbox + late setup (known-answer probes)."""
    v = np.ascontiguousarray(np.asarray(verts, dtype=np.float32).reshape(12))
    cc = C.c_int(0)
    pos = np.zeros(3, dtype=np.uint32)
    bbox = np.zeros(2, dtype=np.uint32)
    edges = np.zeros(9, dtype=np.int32)
    zw = np.zeros(7, dtype=np.float32)
    keep = lib().orc_probe_triangle(_p(v), width, height, cull_mode, 1 if guardband else 0, C.byref(cc),
                                    _p(pos), _p(bbox), _p(edges), _p(zw))
    return dict(keep=int(keep), cc=cc.value, pos=pos, bbox=bbox, edges=edges, zw=zw)


def set_reciprocal_mode(mode: int) -> bool:
    """0 = canonical arithmetic (default, what parity is defined against). Bit 0: vrcp14ps + one Newton-Raphson step instead
    of IEEE division, the code clang most likely emits for the reference's -ffast-math vector divisions;
    bit 1: the back-face determinant contracted into one FMA (App. B.1b). Sensitivity studies only
    (tools/rcp14_sensitivity.py). Returns False when bit 0 is asked for on a host without AVX-512F (mode unchanged)."""
    return lib().orc_set_reciprocal_mode(C.c_int(mode)) == 0


def unpack_meshlets(packed: np.ndarray) -> np.ndarray:
    """Decode of swr_meshlet_packed (glimpsw_b200.compress.pack_meshlets) into Meshlets: positions = fmaf(q, Scale, Origin)."""
    from glimpsw_b200.layout import MESHLET_DTYPE
    assert packed.dtype.itemsize == 1376
    packed = np.ascontiguousarray(packed)
    out = np.zeros(len(packed), dtype=MESHLET_DTYPE)
    lib().orc_unpack_meshlets(_p(packed), C.c_uint32(len(packed)), _p(out))
    return out
