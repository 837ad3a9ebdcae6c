"""Scene generators, meshletizer and texture storage (host side, CPU only)."""
import numpy as np

from glimpsw_b200 import scenes, textures as tx
from glimpsw_b200.layout import MESHLET_DTYPE


def test_config_sizes():
    s = scenes.grid_scene()
    assert s.num_triangles == 999600 and len(s.meshlets) == 10200        # BASELINE config C2
    assert s.meshlets.dtype == MESHLET_DTYPE and s.meshlets.dtype.itemsize == 1728


def test_meshletizer_invariants():
    v, t = scenes.icosphere(3)
    ms = scenes.meshletize(v, t)
    assert ms["NumTriangles"].astype(int).sum() == len(t)
    assert ms["NumVertices"].max() <= 64 and ms["NumTriangles"].max() <= 128
    # every local index refers to a valid vertex and reproduces the original positions
    tri = 0
    for m in ms:
        nt, nv = int(m["NumTriangles"]), int(m["NumVertices"])
        assert m["Indices"][:, :nt].max() < nv
        for k in range(nt):
            for c in range(3):
                p = m["Positions"][:, m["Indices"][c, k]]
                assert np.allclose(p, v[t[tri, c]].astype(np.float32))
            tri += 1
    # bounding spheres contain their vertices
    for m in ms[:20]:
        nv = int(m["NumVertices"])
        d = np.linalg.norm(m["Positions"][:, :nv].T - m["BoundCenter"], axis=1)
        assert d.max() <= m["BoundRadius"] * (1 + 1e-5)


def test_concat_keeps_layout():
    a = np.zeros(2, MESHLET_DTYPE)
    b = np.zeros(3, MESHLET_DTYPE)
    c = scenes.concat_meshlets([a, b])
    assert c.dtype == MESHLET_DTYPE and len(c) == 5


def test_texture_layout_and_mips_match_oracle(orc):
    t = tx.procedural_material_texture(64, seed=3)
    # CreateTexture2D (Texture.h:600-636): 64x64 with 8 requested levels stops at 4x4 -> 5 levels
    assert t.mip_levels == 5 and t.row_shift == 6
    assert list(t.mip_offsets[:5]) == [0, 4096, 4096 + 1024, 4096 + 1024 + 256, 4096 + 1024 + 256 + 64]
    assert t.layer_stride == 4096 + 1024 + 256 + 64 + 64
    # the numpy mip generator equals the oracle's C restatement of Texture2D::GenerateMip
    ref = t.data.copy()
    scratch = tx.TextureData(t.width, t.height, t.mip_levels, t.num_layers, t.row_shift, t.layer_stride, t.mip_offsets, ref.copy())
    for layer in range(t.num_layers):
        for level in range(1, t.mip_levels):
            scratch.data = orc.generate_mip(scratch, layer, level)
    assert np.array_equal(scratch.data, t.data)
    # TiledY8 round trip
    px = tx.get_pixels(t, 0, 0)
    t2 = tx.create_texture(64, 64, 8, 2)
    tx.set_pixels(t2, px, 0, 0)
    assert np.array_equal(tx.get_pixels(t2, 0, 0), px)
