"""glimpsw_b200 — B200-native meshlet raster path behind GLimpSW's swr:: API surface.

The product is libswrb.so (hand-written sm_100a CUDA behind the C ABI in include/swrb.h).
This package is its host-side mirror for Python callers: `api` binds the C ABI with ctypes and
re-exposes the reference's names (Rasterizer, Framebuffer, ShadingContext); `scenes`/`camera`
generate the benchmark inputs. There is no CPU fallback: every call fails loudly without the
CUDA library and a GPU.
"""
from . import layout, camera  # noqa: F401

__all__ = ["layout", "camera", "scenes", "api", "build"]
