"""Host-side mirror of the reference's swr:: interface over the C ABI (include/swrb.h).

Names follow the reference so call sites read like its frame loop (src/SwRast/Main.cpp:213-252):

    rast = Rasterizer()                                   # swr::Rasterizer
    fb = rast.create_framebuffer(1920, 1080)              # swr::CreateFramebuffer
    scene = rast.upload_scene(meshlets, materials, ...)   # Scene::{Meshlets,Materials,Textures,Lights}
    fb.clear(0xFF000000, 0.0)                             # Framebuffer::Clear
    bitmap, n = rast.cull_meshlets(scene, off, cnt, P, V, M)   # ShadingContext::CullMeshlets
    rast.draw_meshlets(fb, scene, off, cnt, object_to_clip)    # Rasterizer::DrawMeshlets + VisBufferShader
    rast.resolve(fb, scene, uniforms)                     # ShadingContext::Resolve
    pixels = fb.get_pixels(0)                             # Framebuffer::GetPixels

Everything executes in libswrb.so on the GPU; this module never computes pixels itself and raises
if the library or a CUDA device is missing.
"""
from __future__ import annotations

import ctypes as C
import os
import weakref

import numpy as np

from .layout import MESHLET_DTYPE, PACKED_MESHLET_DTYPE, MATERIAL_DTYPE, LIGHT_DTYPE

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libswrb.so")

FLAG_BINNING = 1 << 0
FLAG_CLIPPING = 1 << 1
FLAG_GUARDBAND = 1 << 2
FLAG_FUSED_FRUSTUM_CULL = 1 << 3
FLAG_NO_RESOLVE_CACHE = 1 << 4
FLAGS_DEFAULT = FLAG_BINNING | FLAG_CLIPPING | FLAG_GUARDBAND

PERF_NAMES = ["TrianglesProcessed", "TrianglesRasterized", "TrianglesClipped", "BinQueueFlushes",
              "DrawTime", "ResolveTime", "ShadowTime", "FrameTime"]
STAGE_NAMES = ["clear", "cull", "mesh", "bin", "raster", "resolve"]

# every symbol include/swrb.h declares (checked by tests/test_abi.py)
EXPORTS = [
    "swrb_device_create", "swrb_device_destroy", "swrb_device_set_stream", "swrb_device_set_flags",
    "swrb_device_reserve", "swrb_sync", "swrb_last_error", "swrb_version", "swrb_get_counters",
    "swrb_reset_counters", "swrb_scene_create", "swrb_scene_update_meshlets", "swrb_scene_destroy",
    "swrb_scene_create_packed", "swrb_scene_update_packed", "swrb_scene_download_meshlets", "swrb_fb_keys_device", "swrb_fb_keys_touched",
    "swrb_fb_create", "swrb_fb_destroy", "swrb_fb_info", "swrb_fb_clear", "swrb_fb_clear_layer",
    "swrb_fb_set_scissor_rows", "swrb_fb_get_scissor_rows",
    "swrb_fb_download_tiled", "swrb_fb_upload_tiled", "swrb_fb_get_pixels", "swrb_fb_get_pixels_device",
    "swrb_cull_meshlets", "swrb_frustum_planes", "swrb_draw_meshlets", "swrb_draw_batch",
    "swrb_draw_meshlets_host", "swrb_resolve", "swrb_timer_begin", "swrb_timer_end", "swrb_flush_l2",
    "swrb_device_enable_stage_timing", "swrb_get_stage_times", "swrb_get_launch_count",
    "swrb_alloc_pinned", "swrb_free_pinned", "swrb_get_draw_stats", "swrb_fb_get_pixels_device_on_stream",
    "swrb_fb_get_pixels_async", "swrb_hiz_create", "swrb_hiz_destroy", "swrb_hiz_info", "swrb_hiz_build",
    "swrb_hiz_download", "swrb_cull_meshlets_hiz", "swrb_draw_batch_program", "swrb_resolve_debug",
    "swrb_fb_send_pixels", "swrb_peer_collect", "swrb_device_set_mesh_occupancy",
    "swrb_scene_set_skybox", "swrb_batch_create", "swrb_batch_destroy", "swrb_draw_prepared", "swrb_frame_submit",
    "swrb_scene_meshlets_device", "swrb_scene_touch",
]

PROGRAM_VISBUFFER, PROGRAM_OVERDRAW, PROGRAM_DEFERRED = 0, 1, 2   # ShadingContext::VisBufferShader / OverdrawShader / DeferredShader (Shading.h:49)
# enum class DebugLayer (Shading.h:8)
DEBUG_LAYERS = ["None", "BaseColor", "Normals", "MetallicRoughness", "MeshletId", "TriangleId", "OverdrawPixel", "OverdrawQuad"]


class SwrbError(RuntimeError):
    pass


class DrawDesc(C.Structure):
    _fields_ = [("MeshletOffset", C.c_uint32), ("MeshletCount", C.c_uint32), ("ObjectToClip", C.c_float * 16),
                ("CullBitmapHost", C.c_void_p), ("UseDeviceCullBitmap", C.c_int32), ("FrustumPlanes", C.c_float * 20),
                ("ObjectToWorld", C.c_float * 9)]


class ShadingUniforms(C.Structure):
    _fields_ = [("WorldToClip", C.c_float * 16), ("ObjectToClip", C.c_float * 16), ("ObjectToWorld", C.c_float * 9),
                ("InvScreenProj", C.c_float * 16), ("ViewPos", C.c_float * 3), ("Exposure", C.c_float)]


class TextureDesc(C.Structure):
    _fields_ = [("Width", C.c_uint32), ("Height", C.c_uint32), ("MipLevels", C.c_uint32), ("NumLayers", C.c_uint32),
                ("RowShift", C.c_uint32), ("LayerStride", C.c_uint32), ("MipOffsets", C.c_uint32 * 16),
                ("Data", C.c_void_p)]


class PeerSync(C.Structure):     # swrb_peer_sync
    _fields_ = [("WaitFlag", C.c_void_p), ("WaitValue", C.c_uint64), ("SignalFlag", C.c_void_p), ("SignalValue", C.c_uint64)]


class FrameDesc(C.Structure):    # swrb_frame_desc
    _fields_ = [("ClearColor", C.c_uint32), ("ClearDepth", C.c_float), ("Batch", C.c_void_p), ("Uniforms", C.c_void_p),
                ("PixelsDevice", C.c_void_p), ("PixelsStride", C.c_uint32), ("PixelsStream", C.c_void_p),
                ("PeerSync", C.c_void_p), ("PixelsHost", C.c_void_p)]


class FbInfo(C.Structure):
    _fields_ = [("Width", C.c_uint32), ("Height", C.c_uint32), ("TileStride", C.c_uint32),
                ("LayerStride", C.c_uint32), ("NumLayers", C.c_uint32)]


_lib = None


def load_library() -> C.CDLL:
    """Loads libswrb.so. No fallback: a missing library is an error."""
    global _lib
    if _lib is None:
        path = os.environ.get("SWRB_LIB", LIB_PATH)      # developer A/B builds only; there is still no non-CUDA path
        if not os.path.exists(path):
            raise SwrbError(f"{path} is missing: build it with `python -m glimpsw_b200.build` "
                            "(nvcc, sm_100a). There is no CPU fallback.")
        lib = C.CDLL(path)
        lib.swrb_last_error.restype = C.c_char_p
        lib.swrb_version.restype = C.c_char_p
        lib.swrb_device_destroy.restype = None
        lib.swrb_scene_destroy.restype = None
        lib.swrb_fb_destroy.restype = None
        lib.swrb_hiz_destroy.restype = None
        lib.swrb_batch_destroy.restype = None
        for name in ("swrb_device_destroy", "swrb_scene_destroy", "swrb_fb_destroy", "swrb_hiz_destroy", "swrb_batch_destroy"):
            getattr(lib, name).argtypes = [C.c_void_p]
        lib.swrb_frame_submit.argtypes = [C.c_void_p, C.c_void_p]
        lib.swrb_draw_prepared.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32]
        lib.swrb_device_reserve.argtypes = [C.c_void_p, C.c_uint64, C.c_uint64]
        _lib = lib
    return _lib


def _check(rc: int):
    if rc != 0:
        raise SwrbError(f"swrb error {rc}: {load_library().swrb_last_error().decode()}")


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _mat(m) -> np.ndarray:
    return np.ascontiguousarray(np.asarray(m, dtype=np.float32).reshape(-1))


class Framebuffer:
    """swr::Framebuffer (Rasterizer.h:10-78) living in HBM."""

    def __init__(self, rast: "Rasterizer", width: int, height: int, layers: int = 2):
        self.rast, self.width, self.height, self.layers = rast, width, height, layers
        self._h = C.c_void_p()
        _check(rast.lib.swrb_fb_create(rast._h, C.c_uint32(width), C.c_uint32(height), C.c_uint32(layers), C.byref(self._h)))
        info = FbInfo()
        _check(rast.lib.swrb_fb_info(self._h, C.byref(info)))
        self.layer_stride = info.LayerStride
        rast._children.add(self)

    def clear(self, color: int, depth: float):
        _check(self.rast.lib.swrb_fb_clear(self._h, C.c_uint32(color), C.c_float(depth)))

    def set_scissor_rows(self, y0: int = 0, y1: int = 0):
        """Only rows [y0, y1) are drawn, resolved and read back from now on (sort-first split, swrb.h); (0, 0) = the whole framebuffer."""
        _check(self.rast.lib.swrb_fb_set_scissor_rows(self._h, C.c_uint32(y0), C.c_uint32(y1)))

    def scissor_rows(self) -> tuple[int, int]:
        a, b = C.c_uint32(0), C.c_uint32(0)
        _check(self.rast.lib.swrb_fb_get_scissor_rows(self._h, C.byref(a), C.byref(b)))
        return int(a.value), int(b.value)

    def clear_layer(self, layer: int, value: int):
        _check(self.rast.lib.swrb_fb_clear_layer(self._h, C.c_uint32(layer), C.c_uint32(value)))

    def download_tiled(self, layer: int) -> np.ndarray:
        """Raw 4x4-tiled layer data (GetLayerData), W*H u32."""
        out = np.empty(self.width * self.height, dtype=np.uint32)
        _check(self.rast.lib.swrb_fb_download_tiled(self._h, C.c_uint32(layer), _ptr(out)))
        return out

    def keys_device(self):
        """(device pointer, words) of the 64-bit key buffer holding the pending vis-buffer (sort-last composition, swrb.h)."""
        p, n = C.c_void_p(), C.c_uint64(0)
        _check(self.rast.lib.swrb_fb_keys_device(self._h, C.byref(p), C.byref(n)))
        return int(p.value or 0), int(n.value)

    def keys_touched(self):
        _check(self.rast.lib.swrb_fb_keys_touched(self._h))

    def upload_tiled(self, layer: int, data: np.ndarray):
        data = np.ascontiguousarray(data, dtype=np.uint32)
        assert data.size >= self.width * self.height
        _check(self.rast.lib.swrb_fb_upload_tiled(self._h, C.c_uint32(layer), _ptr(data)))

    def get_pixels(self, layer: int, out: np.ndarray | None = None) -> np.ndarray:
        """Framebuffer::GetPixels (ImageHelpers.cpp:109-147): row-major [H, W] u32."""
        if out is None:
            out = np.empty((self.height, self.width), dtype=np.uint32)
        _check(self.rast.lib.swrb_fb_get_pixels(self._h, C.c_uint32(layer), _ptr(out), C.c_uint32(out.shape[1])))
        return out

    def get_pixels_async(self, layer: int, out: np.ndarray):
        """GetPixels without the final synchronisation (`out` should be pinned; valid after Rasterizer.sync())."""
        _check(self.rast.lib.swrb_fb_get_pixels_async(self._h, C.c_uint32(layer), _ptr(out), C.c_uint32(out.shape[1])))

    def get_pixels_device(self, layer: int, device_ptr: int, stride: int | None = None, cuda_stream: int | None = None):
        """GetPixels into device (or NVLink peer) memory; `cuda_stream` launches it on another stream."""
        if cuda_stream is None:
            _check(self.rast.lib.swrb_fb_get_pixels_device(self._h, C.c_uint32(layer), C.c_void_p(device_ptr),
                                                           C.c_uint32(stride or self.width)))
        else:
            _check(self.rast.lib.swrb_fb_get_pixels_device_on_stream(self._h, C.c_uint32(layer), C.c_void_p(device_ptr),
                                                                     C.c_uint32(stride or self.width), C.c_void_p(cuda_stream)))

    def send_pixels(self, layer: int, device_ptr: int, cuda_stream: int, wait_flag: int = 0, wait_value: int = 0,
                    signal_flag: int = 0, signal_value: int = 0, stride: int | None = None):
        """GetPixels into another GPU's memory with the slot flow control folded into the kernel (swrb_fb_send_pixels)."""
        ps = PeerSync(wait_flag or None, wait_value, signal_flag or None, signal_value)
        _check(self.rast.lib.swrb_fb_send_pixels(self._h, C.c_uint32(layer), C.c_void_p(device_ptr), C.c_uint32(stride or self.width),
                                                 C.c_void_p(cuda_stream), C.byref(ps)))

    def destroy(self):
        if self._h:
            self.rast.lib.swrb_fb_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.destroy()
        except Exception:
            pass


class Scene:
    """Device-resident Scene::{Meshlets, Materials, Textures, Lights} (Scene.h:117-123)."""

    def __init__(self, rast: "Rasterizer", meshlets: np.ndarray, materials=None, textures=None, lights=None):
        packed = meshlets.dtype == PACKED_MESHLET_DTYPE          # swr_meshlet_packed: decoded on the device at upload
        assert packed or meshlets.dtype == MESHLET_DTYPE
        self.rast = rast
        self.num_meshlets = len(meshlets)
        meshlets = np.ascontiguousarray(meshlets)
        materials = np.zeros(0, MATERIAL_DTYPE) if materials is None else np.ascontiguousarray(materials, dtype=MATERIAL_DTYPE)
        lights = np.zeros(0, LIGHT_DTYPE) if lights is None else np.ascontiguousarray(lights, dtype=LIGHT_DTYPE)
        textures = textures or []
        descs = (TextureDesc * max(len(textures), 1))()
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
        self._h = C.c_void_p()
        create = rast.lib.swrb_scene_create_packed if packed else rast.lib.swrb_scene_create
        _check(create(rast._h, _ptr(meshlets), C.c_uint32(len(meshlets)),
                      _ptr(materials) if len(materials) else None, C.c_uint32(len(materials)),
                      descs if textures else None, C.c_uint32(len(textures)),
                      _ptr(lights) if len(lights) else None, C.c_uint32(len(lights)),
                      C.byref(self._h)))
        rast._children.add(self)

    def set_skybox(self, tex):
        """ShadingContext::SkyboxTex: an octahedron-mapped Texture2D<R11G11B10f> (textures.procedural_sky_texture); None removes it."""
        if tex is None:
            _check(self.rast.lib.swrb_scene_set_skybox(self._h, None))
            return
        data = np.ascontiguousarray(tex.data, dtype=np.uint32)
        d = TextureDesc()
        d.Width, d.Height, d.MipLevels, d.NumLayers = tex.width, tex.height, tex.mip_levels, tex.num_layers
        d.RowShift, d.LayerStride = tex.row_shift, tex.layer_stride
        for k in range(16):
            d.MipOffsets[k] = int(tex.mip_offsets[k])
        d.Data = data.ctypes.data
        _check(self.rast.lib.swrb_scene_set_skybox(self._h, C.byref(d)))

    def update_meshlets(self, meshlets: np.ndarray, first: int = 0):
        meshlets = np.ascontiguousarray(meshlets)
        fn = self.rast.lib.swrb_scene_update_packed if meshlets.dtype == PACKED_MESHLET_DTYPE else self.rast.lib.swrb_scene_update_meshlets
        _check(fn(self._h, _ptr(meshlets), C.c_uint32(first), C.c_uint32(len(meshlets))))

    def download_meshlets(self, first: int = 0, count: int | None = None) -> np.ndarray:
        count = self.num_meshlets - first if count is None else count
        out = np.zeros(count, dtype=MESHLET_DTYPE)
        _check(self.rast.lib.swrb_scene_download_meshlets(self._h, _ptr(out), C.c_uint32(first), C.c_uint32(count)))
        return out

    def meshlets_device_ptr(self) -> int:
        """Device address of the scene's meshlet array (num_meshlets x 1728 bytes), for callers that fill it on the device."""
        p = C.c_void_p()
        _check(self.rast.lib.swrb_scene_meshlets_device(self._h, C.byref(p)))
        return int(p.value or 0)

    def touch(self, first: int = 0, count: int | None = None):
        """Declares meshlets [first, first + count) rewritten on the device (derived tables are rebuilt lazily)."""
        _check(self.rast.lib.swrb_scene_touch(self._h, C.c_uint32(first), C.c_uint32(self.num_meshlets - first if count is None else count)))

    def destroy(self):
        if self._h:
            self.rast.lib.swrb_scene_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.destroy()
        except Exception:
            pass


class Batch:
    """swrb_batch: a prepared list of DrawMeshlets calls (one per glTF node, Main.cpp:216-240) living on the device."""

    def __init__(self, rast: "Rasterizer", scene: Scene, draws: list):
        self.rast, self.scene = rast, scene
        arr, n, keep = rast.make_batch(draws)
        self._h = C.c_void_p()
        _check(rast.lib.swrb_batch_create(scene._h, arr, C.c_uint32(n), C.byref(self._h)))
        rast._children.add(self)

    def destroy(self):
        if self._h:
            self.rast.lib.swrb_batch_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.destroy()
        except Exception:
            pass


class DepthPyramid:
    """The half-resolution R32f min pyramid of the previous frame's depth (Main.cpp:54-56, ImageHelpers.cpp:150-247)."""

    def __init__(self, rast: "Rasterizer", fb_width: int, fb_height: int):
        self.rast = rast
        self._h = C.c_void_p()
        _check(rast.lib.swrb_hiz_create(rast._h, C.c_uint32(fb_width), C.c_uint32(fb_height), C.byref(self._h)))
        d = TextureDesc()
        _check(rast.lib.swrb_hiz_info(self._h, C.byref(d)))
        self.width, self.height, self.mip_levels, self.row_shift = d.Width, d.Height, d.MipLevels, d.RowShift
        self.layer_stride = d.LayerStride
        self.mip_offsets = np.array(list(d.MipOffsets), dtype=np.uint32)
        rast._children.add(self)

    def build(self, fb: Framebuffer):
        """texutil::DownsampleDepth(fb, self)."""
        _check(self.rast.lib.swrb_hiz_build(self._h, fb._h))

    def download(self) -> np.ndarray:
        out = np.empty(self.layer_stride, dtype=np.float32)
        _check(self.rast.lib.swrb_hiz_download(self._h, _ptr(out)))
        return out

    def destroy(self):
        if self._h:
            self.rast.lib.swrb_hiz_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.destroy()
        except Exception:
            pass


class Rasterizer:
    """swr::Rasterizer (Rasterizer.h:202-340) bound to one CUDA device."""

    def __init__(self, cuda_device: int = 0, enable_binning: bool = True, enable_clipping: bool = True,
                 enable_guardband: bool = True, fused_frustum_cull: bool = False, resolve_cache: bool = True):
        self.lib = load_library()
        self._h = C.c_void_p()
        self._children = weakref.WeakSet()   # framebuffers / scenes must die before the device
        self._pinned = []
        _check(self.lib.swrb_device_create(C.c_int(cuda_device), C.byref(self._h)))
        self.set_flags(enable_binning, enable_clipping, enable_guardband, fused_frustum_cull, resolve_cache)

    # Rasterizer::EnableBinning / EnableClipping / EnableGuardband (Rasterizer.h:206-208)
    def set_flags(self, enable_binning=True, enable_clipping=True, enable_guardband=True, fused_frustum_cull=False,
                  resolve_cache=True):
        self.flags = ((FLAG_BINNING if enable_binning else 0) | (FLAG_CLIPPING if enable_clipping else 0) |
                      (FLAG_GUARDBAND if enable_guardband else 0) | (FLAG_FUSED_FRUSTUM_CULL if fused_frustum_cull else 0) |
                      (0 if resolve_cache else FLAG_NO_RESOLVE_CACHE))
        _check(self.lib.swrb_device_set_flags(self._h, C.c_uint32(self.flags)))

    def set_mesh_occupancy(self, blocks_per_sm: int):
        """Persistent-grid size of the mesh kernel (1..4 blocks per SM; 2 suits several contexts in flight)."""
        _check(self.lib.swrb_device_set_mesh_occupancy(self._h, C.c_uint32(blocks_per_sm)))

    def set_stream(self, cuda_stream: int | None):
        _check(self.lib.swrb_device_set_stream(self._h, C.c_void_p(cuda_stream or 0)))

    def reserve(self, max_triangles: int, max_bin_entries: int):
        _check(self.lib.swrb_device_reserve(self._h, C.c_uint64(max_triangles), C.c_uint64(max_bin_entries)))

    def create_framebuffer(self, width: int, height: int, layers: int = 2) -> Framebuffer:
        return Framebuffer(self, width, height, layers)

    def upload_scene(self, meshlets, materials=None, textures=None, lights=None) -> Scene:
        return Scene(self, meshlets, materials, textures, lights)

    # ShadingContext::CullMeshlets (Shading.cpp:775-869, frustum part)
    def cull_meshlets(self, scene: Scene, offset: int, count: int, proj, view, model, download: bool = True):
        bitmap = np.zeros((count + 15) // 16, dtype=np.uint16) if download else None
        vis = C.c_uint32(0)
        _check(self.lib.swrb_cull_meshlets(scene._h, C.c_uint32(offset), C.c_uint32(count), _ptr(_mat(proj)),
                                           _ptr(_mat(view)), _ptr(_mat(model)), _ptr(bitmap),
                                           C.byref(vis) if download else None))
        return bitmap, vis.value

    # HiZ occlusion culling: texutil::DownsampleDepth + the HiZ half of ShadingContext::CullMeshlets
    def create_hiz(self, fb_width: int, fb_height: int) -> "DepthPyramid":
        return DepthPyramid(self, fb_width, fb_height)

    def cull_meshlets_hiz(self, scene: Scene, offset: int, count: int, proj, view, model, prev_view, frame_w: int, frame_h: int,
                          hiz: "DepthPyramid | None" = None, download: bool = True):
        bitmap = np.zeros((count + 15) // 16, dtype=np.uint16) if download else None
        vis = C.c_uint32(0)
        _check(self.lib.swrb_cull_meshlets_hiz(scene._h, C.c_uint32(offset), C.c_uint32(count), _ptr(_mat(proj)), _ptr(_mat(view)),
                                               _ptr(_mat(model)), _ptr(_mat(prev_view)), C.c_float(frame_w), C.c_float(frame_h),
                                               hiz._h if hiz is not None else None, _ptr(bitmap), C.byref(vis) if download else None))
        return bitmap, vis.value

    def frustum_planes(self, proj, view, model) -> np.ndarray:
        out = np.zeros((5, 4), dtype=np.float32)
        _check(self.lib.swrb_frustum_planes(_ptr(_mat(proj)), _ptr(_mat(view)), _ptr(_mat(model)), _ptr(out)))
        return out

    @staticmethod
    def _desc(offset, count, object_to_clip, cull_bitmap=None, use_device_bitmap=False, planes=None, keep=None,
              object_to_world3=None) -> DrawDesc:
        d = DrawDesc()
        d.MeshletOffset, d.MeshletCount = offset, count
        d.ObjectToClip[:] = _mat(object_to_clip).tolist()
        d.ObjectToWorld[:] = [1.0, 0.0, 0.0, 0.0, 1.0, 0.0, 0.0, 0.0, 1.0] if object_to_world3 is None else _mat(object_to_world3).tolist()
        if cull_bitmap is not None:
            cb = np.ascontiguousarray(cull_bitmap, dtype=np.uint16)
            if keep is not None:
                keep.append(cb)
            d.CullBitmapHost = cb.ctypes.data
        d.UseDeviceCullBitmap = 1 if use_device_bitmap else 0
        if planes is not None:
            d.FrustumPlanes[:] = np.asarray(planes, dtype=np.float32).reshape(-1)[:20].tolist()
        return d

    # Rasterizer::DrawMeshlets(fb, count, {ShadingContext::VisBufferShader, &ctx}) (Rasterizer.cpp:493)
    def draw_meshlets(self, fb: Framebuffer, scene: Scene, meshlet_offset: int, count: int, object_to_clip,
                      cull_bitmap=None, use_device_bitmap=False, planes=None):
        keep = []
        d = self._desc(meshlet_offset, count, object_to_clip, cull_bitmap, use_device_bitmap, planes, keep)
        _check(self.lib.swrb_draw_meshlets(fb._h, scene._h, C.byref(d)))

    def draw_batch(self, fb: Framebuffer, scene: Scene, draws: list, program: int = PROGRAM_VISBUFFER):
        """draws: list of dicts(offset, count, object_to_clip[, cull_bitmap, use_device_bitmap, planes]).
        program: the shader table of the DrawMeshlets calls — PROGRAM_VISBUFFER, or PROGRAM_OVERDRAW (FS_Overdraw)."""
        keep = []
        arr = (DrawDesc * len(draws))()
        for i, dd in enumerate(draws):
            arr[i] = self._desc(dd["offset"], dd["count"], dd["object_to_clip"], dd.get("cull_bitmap"),
                                dd.get("use_device_bitmap", False), dd.get("planes"), keep, dd.get("object_to_world3"))
        if program == PROGRAM_VISBUFFER:
            _check(self.lib.swrb_draw_batch(fb._h, scene._h, arr, C.c_uint32(len(draws))))
        else:
            _check(self.lib.swrb_draw_batch_program(fb._h, scene._h, arr, C.c_uint32(len(draws)), C.c_uint32(program)))

    def make_batch(self, draws: list):
        """Pre-builds the descriptor array of draw_batch for repeated submission (bench loops)."""
        keep = []
        arr = (DrawDesc * len(draws))()
        for i, dd in enumerate(draws):
            arr[i] = self._desc(dd["offset"], dd["count"], dd["object_to_clip"], dd.get("cull_bitmap"),
                                dd.get("use_device_bitmap", False), dd.get("planes"), keep, dd.get("object_to_world3"))
        return arr, len(draws), keep

    def create_batch(self, scene: Scene, draws: list) -> "Batch":
        """swrb_batch_create: the frame's DrawMeshlets calls validated and uploaded once (same dict form as draw_batch)."""
        return Batch(self, scene, draws)

    def draw_prepared(self, fb: Framebuffer, batch: "Batch", program: int = PROGRAM_VISBUFFER):
        _check(self.lib.swrb_draw_prepared(fb._h, batch._h, C.c_uint32(program)))

    def make_frame(self, batch: "Batch", uniforms: ShadingUniforms | None, clear_color: int = 0xFF000000, clear_depth: float = 0.0,
                   pixels_device: int = 0, pixels_stream: int = 0, pixels_host: np.ndarray | None = None, peer_sync: PeerSync | None = None,
                   pixels_stride: int = 0) -> FrameDesc:
        """A swrb_frame_desc (Clear -> batch -> Resolve -> GetPixels); build once per view, submit every frame."""
        f = FrameDesc()
        f.ClearColor, f.ClearDepth, f.Batch = clear_color, clear_depth, batch._h
        f.Uniforms = C.cast(C.pointer(uniforms), C.c_void_p) if uniforms is not None else None
        f.PixelsDevice, f.PixelsStride, f.PixelsStream = pixels_device or None, pixels_stride, pixels_stream or None
        f.PeerSync = C.cast(C.pointer(peer_sync), C.c_void_p) if peer_sync is not None else None
        f.PixelsHost = pixels_host.ctypes.data if pixels_host is not None else None
        f._keep = (batch, uniforms, peer_sync, pixels_host)
        return f

    def submit_frame(self, fb: Framebuffer, frame: FrameDesc):
        """swrb_frame_submit: one iteration of the reference's frame loop (Main.cpp:213-252) in one call."""
        rc = self.lib.swrb_frame_submit(fb._h, C.addressof(frame))
        if rc:
            _check(rc)

    def draw_prebuilt(self, fb: Framebuffer, scene: Scene, batch):
        _check(self.lib.swrb_draw_batch(fb._h, scene._h, batch[0], C.c_uint32(batch[1])))

    def draw_meshlets_host(self, fb: Framebuffer, meshlets: np.ndarray, object_to_clip, cull_bitmap=None):
        """Literal drop-in form: ShadingContext::Meshlets is a host pointer uploaded on every call."""
        cb = None if cull_bitmap is None else np.ascontiguousarray(cull_bitmap, dtype=np.uint16)
        m = _mat(object_to_clip)
        _check(self.lib.swrb_draw_meshlets_host(fb._h, _ptr(meshlets), C.c_uint32(len(meshlets)), _ptr(m), _ptr(cb)))

    # ShadingContext::Resolve (Shading.cpp:658-689)
    @staticmethod
    def make_uniforms(world_to_clip, object_to_clip, object_to_world3, inv_screen_proj, view_pos, exposure: float = 1.0) -> ShadingUniforms:
        """The ShadingContext uniform block Resolve reads, as the C struct (build once, reuse per frame)."""
        u = ShadingUniforms()
        u.WorldToClip[:] = _mat(world_to_clip).tolist()
        u.ObjectToClip[:] = _mat(object_to_clip).tolist()
        u.ObjectToWorld[:] = _mat(object_to_world3).tolist()
        u.InvScreenProj[:] = _mat(inv_screen_proj).tolist()
        u.ViewPos[:] = [float(v) for v in view_pos]
        u.Exposure = float(exposure)
        return u

    def resolve(self, fb: Framebuffer, scene: Scene, world_to_clip, object_to_clip, object_to_world3, inv_screen_proj,
                view_pos, exposure: float = 1.0):
        u = self.make_uniforms(world_to_clip, object_to_clip, object_to_world3, inv_screen_proj, view_pos, exposure)
        _check(self.lib.swrb_resolve(fb._h, scene._h, C.byref(u)))

    # ShadingContext::ResolveDebug (Shading.cpp:734-773)
    def resolve_debug(self, fb: Framebuffer, scene: Scene, layer, world_to_clip, object_to_clip, object_to_world3,
                      inv_screen_proj, view_pos=(0.0, 0.0, 0.0), exposure: float = 1.0):
        layer = DEBUG_LAYERS.index(layer) if isinstance(layer, str) else int(layer)
        u = self.make_uniforms(world_to_clip, object_to_clip, object_to_world3, inv_screen_proj, view_pos, exposure)
        _check(self.lib.swrb_resolve_debug(fb._h, scene._h, C.byref(u), C.c_uint32(layer)))

    def resolve_prebuilt(self, fb: Framebuffer, scene: Scene, uniforms: ShadingUniforms):
        _check(self.lib.swrb_resolve(fb._h, scene._h, C.byref(uniforms)))

    def peer_collect(self, cuda_stream: int, ready_flags_ptr: int, n: int, expected: int, ack_flag_ptrs, ack_value: int):
        """Consumer side of the composite exchange (swrb_peer_collect): wait for n ready flags, then ack every producer."""
        arr = (C.c_void_p * n)(*[int(p) for p in ack_flag_ptrs])
        _check(self.lib.swrb_peer_collect(self._h, C.c_void_p(cuda_stream), C.c_void_p(ready_flags_ptr), C.c_uint32(n),
                                          C.c_uint64(expected), arr, C.c_uint64(ack_value)))

    def sync(self):
        _check(self.lib.swrb_sync(self._h))

    # perf::GetCurrent / perf::Reset
    def counters(self) -> dict:
        out = (C.c_uint64 * 8)()
        _check(self.lib.swrb_get_counters(self._h, out))
        return dict(zip(PERF_NAMES, [int(v) for v in out]))

    def reset_counters(self):
        _check(self.lib.swrb_reset_counters(self._h))

    # timing helpers
    def timer_begin(self):
        _check(self.lib.swrb_timer_begin(self._h))

    def timer_end(self) -> float:
        ms = C.c_float(0)
        _check(self.lib.swrb_timer_end(self._h, C.byref(ms)))
        return ms.value

    def flush_l2(self):
        _check(self.lib.swrb_flush_l2(self._h))

    def enable_stage_timing(self, on: bool = True):
        _check(self.lib.swrb_device_enable_stage_timing(self._h, C.c_int(1 if on else 0)))

    def stage_times_us(self) -> dict:
        t = (C.c_float * 6)()
        n = (C.c_uint32 * 6)()
        _check(self.lib.swrb_get_stage_times(self._h, t, n))
        return {k: (float(t[i]), int(n[i])) for i, k in enumerate(STAGE_NAMES)}

    def alloc_pinned(self, shape, dtype) -> np.ndarray:
        """A numpy array backed by cudaMallocHost memory (freed with the device)."""
        dtype = np.dtype(dtype)
        count = int(np.prod(shape))
        ptr = C.c_void_p()
        _check(self.lib.swrb_alloc_pinned(self._h, C.c_uint64(max(count * dtype.itemsize, 1)), C.byref(ptr)))
        self._pinned.append(ptr)
        buf = (C.c_uint8 * (count * dtype.itemsize)).from_address(ptr.value)
        return np.frombuffer(buf, dtype=dtype, count=count).reshape(shape)

    def launch_count(self) -> int:
        v = C.c_uint64(0)
        _check(self.lib.swrb_get_launch_count(self._h, C.byref(v)))
        return int(v.value)

    def draw_stats(self) -> dict:
        """Work-list sizes of the last draw: triangle records, big-list entries, tile-list entries."""
        out = (C.c_uint32 * 4)()
        _check(self.lib.swrb_get_draw_stats(self._h, out))
        return {"records": int(out[0]), "big": int(out[1]), "bin_entries": int(out[2])}

    def destroy(self):
        if self._h:
            for child in sorted(self._children, key=lambda c: 0 if isinstance(c, Batch) else 1):   # batches before their scenes
                child.destroy()
            for ptr in self._pinned:
                self.lib.swrb_free_pinned(self._h, ptr)
            self._pinned = []
            self.lib.swrb_device_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.destroy()
        except Exception:
            pass
