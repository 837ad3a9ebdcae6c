"""The C-ABI library loads on a CPU-only box, exports every symbol include/swrb.h declares, and the
product fails loudly (no CPU fallback) when there is no CUDA device. No compute calls here."""
import ctypes
import os
import re

import pytest

from glimpsw_b200 import api

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "swrb.h")).read()
    return sorted(set(re.findall(r"SWRB_API\s+[\w\s\*]+?\b(swrb_\w+)\s*\(", text)))


def test_header_and_wrapper_agree():
    assert declared_symbols() == sorted(api.EXPORTS)


def test_library_exports_every_declared_symbol():
    lib = api.load_library()
    for name in declared_symbols():
        assert hasattr(lib, name), f"libswrb.so does not export {name}"
    assert b"sm_100a" in lib.swrb_version()


def test_struct_layouts_match_header():
    # swrb_draw_desc: u32,u32,float[16],ptr,i32,float[20],float[9] -> 8-byte aligned pointer at offset 72
    assert api.DrawDesc.CullBitmapHost.offset == 72
    assert api.DrawDesc.ObjectToWorld.offset == 164
    assert ctypes.sizeof(api.DrawDesc) == 200
    # swrb_frame_desc: u32,f32,ptr,ptr,ptr,u32,(pad),ptr,ptr,ptr
    assert api.FrameDesc.PixelsStream.offset == 40 and ctypes.sizeof(api.FrameDesc) == 64
    assert ctypes.sizeof(api.ShadingUniforms) == (16 + 16 + 9 + 16 + 3 + 1) * 4
    assert ctypes.sizeof(api.TextureDesc) == 6 * 4 + 16 * 4 + 8


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present; the no-device failure path cannot be exercised")
    with pytest.raises(api.SwrbError) as e:
        api.Rasterizer(0)
    assert "no CPU fallback" in str(e.value) or "CUDA" in str(e.value)


def test_product_does_not_import_the_oracle():
    pkg = os.path.join(ROOT, "glimpsw_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                for pat in (r"(from|import)\s+oracle", r"liboracle", r"\borc_\w+\(", r"oracle/", r"\borc\."):
                    assert not re.search(pat, text), f"{f} uses the oracle ({pat}); the product must not"
