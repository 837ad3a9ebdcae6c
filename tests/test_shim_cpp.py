"""include/swr_b200.hpp — the header-only C++ shim that keeps the reference's names over the C ABI.
CPU: it compiles against the headers, links libswrb.so and fails loudly without a device. GPU: the same program
renders a triangle through Rasterizer::CullMeshlets / DrawMeshlets / Framebuffer::GetPixels and matches the oracle."""
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def build_demo(tmp_path):
    from glimpsw_b200 import build
    lib = build.build()
    exe = str(tmp_path / "shim_demo")
    cmd = ["g++", "-std=c++17", "-O1", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "shim_demo.cpp"),
           "-o", exe, "-L", os.path.dirname(lib), "-lswrb", "-Wl,-rpath," + os.path.dirname(lib)]
    subprocess.run(cmd, check=True)
    return exe


def test_shim_compiles_links_and_fails_loudly_without_device(tmp_path):
    exe = build_demo(tmp_path)
    out = subprocess.run([exe], check=True, capture_output=True, text=True).stdout.strip()
    assert out.startswith("OK ") or out.startswith("NO_DEVICE ")
    if out.startswith("NO_DEVICE"):
        assert "no CPU fallback" in out


@pytest.mark.gpu
def test_shim_renders_like_the_oracle(tmp_path, orc):
    from test_oracle_kat import meshlet_from_clip_tris, IDENT
    exe = build_demo(tmp_path)
    out = subprocess.run([exe], check=True, capture_output=True, text=True).stdout.split()
    assert out[0] == "OK", out
    fb = orc.Framebuffer(64, 64)
    fb.clear(0xFFFFFFFF, 0.0)
    c = orc.draw_meshlets(fb, meshlet_from_clip_tris([[(-0.5, -0.5, 0.5), (-0.5, 0.5, 0.5), (0.5, -0.5, 0.5)]]), 0, 1, IDENT)
    ids = fb.get_pixels(0)
    assert int(c[1]) == 1 and int((ids != 0xFFFFFFFF).sum()) > 0
    assert [int(out[1]), int(out[2]), int(out[3]), int(out[4])] == [int((ids != 0xFFFFFFFF).sum()), int(ids[28, 28]), 1, 1]
    # OverdrawShader drawn twice counts every covered pixel twice; ResolveDebug paints sky fragment (0,0) white
    assert int(out[5]) == int((ids != 0xFFFFFFFF).sum()) and out[6] == "ffffffff"
