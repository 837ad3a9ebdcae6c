"""Compiles the REFERENCE's own raster-path translation units with g++ into oracle/_ref/libswr_ref.so.

TEST INFRASTRUCTURE. The reference (dubiousconst282/GLimpSW, mounted at /root/reference) is written in the Clang dialect and pulls
glm / tracy / stb through CPM; neither clang nor those packages exist in this image. This recipe builds the
three translation units of the hot path anyway — src/SwRast/Rasterizer.cpp, Shading.cpp, ImageHelpers.cpp with their headers
Rasterizer.h, Shading.h, Texture.h, Scene.h — from the sources WHERE THEY LIE:

  1. the files are read from /root/reference/src/SwRast and written, with the edits listed in EDITS, into a scratch directory
     outside the repository (nothing of the reference is copied into the repo; the edits name a line and the few tokens they
     replace, and the recipe refuses to run if a line does not contain what the edit expects);
  2. `SIMD.h`, `Camera.h`, `<glm/glm.hpp>`, `<tracy/Tracy.hpp>`, `<stb_image*.h>` resolve to the stand-ins under oracle/compat/
     (a from-scratch g++ implementation of the same vector API; a sliver of GLM; empty profiler macros; declarations only);
  3. oracle/ref_api.cpp — a C wrapper that drives swr::Rasterizer / ShadingContext through their public interface — is
     compiled with them into oracle/_ref/libswr_ref.so (canonical arithmetic: no -ffast-math, no contraction, IEEE 1/x for
     approx_rcp) and oracle/_ref/libswr_ref_fast.so (-Ofast -mrecip + the real vrcp14ps / vrsqrt14ps: what an upstream build's
     flags license, for the sensitivity comparison).

Every edit is one of five kinds, none of which touches arithmetic:
  T  `mask ? a : b` on vectors (a Clang extension)                     -> simd::select(mask, a, b)
  V  Clang-only vector type spellings (`uint8_t [[clang::ext_vector_type(16)]]`, bool vectors)
  A  inline asm with a vector-register constraint on a class type       -> the intrinsic the source itself names in a comment
  I  `#include "Camera.h"`                                              -> the stand-in (the one function used is replaced, see compat_camera.h)
  X  code outside the raster path that the stand-ins do not cover (IBL precomputation, image file loaders) -> `#if 0`

    python oracle/ref_build.py            # build if stale
    python oracle/ref_build.py --force -v
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference/src/SwRast"
OUT_DIR = os.path.join(HERE, "_ref")
LIB = os.path.join(OUT_DIR, "libswr_ref.so")
LIB_FAST = os.path.join(OUT_DIR, "libswr_ref_fast.so")
FILES = ["Rasterizer.h", "Rasterizer.cpp", "Shading.h", "Shading.cpp", "Texture.h", "Scene.h", "ImageHelpers.cpp"]

# (file, line, text the line must contain, replacement for that text). `line` is 1-based in the reference file.
EDITS: list[tuple[str, int, str, str]] = []

# (file, first line, last line): the range is wrapped in `#if 0 ... #endif` (kind X)
DISABLED: list[tuple[str, int, int]] = []


def available() -> bool:
    return os.path.isdir(REF)


def cpu_ok() -> bool:
    """The library uses AVX-512 F/BW/DQ/VL/VBMI + F16C like the reference itself."""
    try:
        flags = open("/proc/cpuinfo").read()
        return all(f in flags for f in ("avx512f", "avx512bw", "avx512dq", "avx512vl", "avx512_vbmi", "f16c"))
    except Exception:
        return False


def _load_edit_tables():
    from oracle import ref_edits
    return ref_edits.EDITS, ref_edits.DISABLED


def stage_sources(dst: str) -> None:
    edits, disabled = _load_edit_tables()
    for name in FILES:
        lines = open(os.path.join(REF, name), encoding="utf-8").read().split("\n")
        for f, ln, find, repl in edits:
            if f != name:
                continue
            if find not in lines[ln - 1]:
                raise RuntimeError(f"{name}:{ln} does not contain {find!r}: the reference differs from the snapshot this recipe was written for")
            lines[ln - 1] = lines[ln - 1].replace(find, repl, 1)
        for f, a, b in disabled:
            if f == name:
                lines[a - 1] = "#if 0\n" + lines[a - 1]
                lines[b - 1] = lines[b - 1] + "\n#endif"
        open(os.path.join(dst, name), "w", encoding="utf-8").write("\n".join(lines))
    open(os.path.join(dst, "SIMD.h"), "w").write('#pragma once\n#include "simd_gxx.h"\n')
    open(os.path.join(dst, "Camera.h"), "w").write('#pragma once\n#include "compat_camera.h"\n')


def is_stale() -> bool:
    if not (os.path.exists(LIB) and os.path.exists(LIB_FAST)):
        return True
    t = min(os.path.getmtime(LIB), os.path.getmtime(LIB_FAST))
    deps = [os.path.join(HERE, "ref_api.cpp"), os.path.join(HERE, "ref_build.py"), os.path.join(HERE, "ref_edits.py")]
    for root, _, files in os.walk(os.path.join(HERE, "compat")):
        deps += [os.path.join(root, f) for f in files]
    if available():
        deps += [os.path.join(REF, f) for f in FILES]
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def build(force: bool = False, verbose: bool = False) -> str | None:
    """Returns the path of libswr_ref.so, or None when the reference sources are not on this machine and no prebuilt
    library travelled with the repository."""
    if not available():
        return LIB if os.path.exists(LIB) else None
    if not force and not is_stale():
        return LIB
    os.makedirs(OUT_DIR, exist_ok=True)
    scratch = tempfile.mkdtemp(prefix="swr_ref_build_")
    try:
        stage_sources(scratch)
        common = ["g++", "-std=c++20", "-shared", "-fPIC", "-march=x86-64-v4", "-mavx512vbmi", "-mf16c", "-fwrapv", "-pthread", "-DNDEBUG",
                  "-w", "-I", scratch, "-I", os.path.join(HERE, "compat"), "-I", os.path.join(HERE, "..", "include")]
        srcs = [os.path.join(scratch, f) for f in ("Rasterizer.cpp", "Shading.cpp", "ImageHelpers.cpp")] + [os.path.join(HERE, "ref_api.cpp")]
        variants = [(LIB, ["-O2", "-fno-fast-math", "-ffp-contract=off"]),
                    (LIB_FAST, ["-Ofast", "-mrecip=all", "-DSWR_COMPAT_RCP14"])]
        for out, flags in variants:
            cmd = common + flags + ["-o", out] + srcs
            r = subprocess.run(cmd, capture_output=True, text=True)
            if verbose or r.returncode != 0:
                sys.stderr.write(r.stdout + r.stderr)
            if r.returncode != 0:
                raise RuntimeError(f"g++ failed building {os.path.basename(out)} (scratch kept at {scratch})")
    except Exception:
        raise
    else:
        shutil.rmtree(scratch, ignore_errors=True)
    return LIB


if __name__ == "__main__":
    sys.path.insert(0, os.path.dirname(HERE))
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
