"""The BASELINE.json configurations as named, reproducible workloads (scene + cameras + which passes run).

One definition shared by bench.py, the golden-vector generator (tests/golden/make_golden_bench.py) and the GPU parity
tests, so the thing that is timed is the thing that is checked. Names follow BASELINE.json `configs`:

  c1_knot     C1 stand-in: 72,000-triangle textured torus knot (two materials, 1024^2 two-layer textures), 1920x1080,
              vis-buffer + resolve                                   (DamagedHelmet is not among the reference's assets)
  c1_sponza   C1 glTF case: the reference's own Sponza_LowPoly.gltf (63,084 triangles, 2 nodes, no materials) from the camera
              pose RasterBench.cpp:68 hard-codes, 1920x1080, vis-buffer + resolve. The imported meshlets travel as a
              fixture (tests/golden/sponza_lowpoly_scene.npz, written by make_golden_bench.py where the asset exists).
  c2_grid     C2: 999,600-triangle heightfield grid, 1920x1080, vis-buffer only
  c3_knot     C3 stand-in: 259,200-triangle torus knot with a double-sided alpha-tested material, 512^2 textures, 2560x1440,
              vis-buffer + resolve                                   (Sponza.bin, the textured 262 K-triangle mesh, is missing)
  c4_views    C4 geometry as a batch of views: 9,994,240 triangles in 122 DrawMeshlets calls (instance transforms baked, see
              scenes.instanced_scene), frustum cull fused into the mesh kernel, 64 seeded orbit cameras, 1920x1080,
              vis-buffer + resolve — the bench headline; step = the whole batch
  c5_views    C5: the same batch at 2048x2048
"""
from __future__ import annotations

import os
from dataclasses import dataclass

import numpy as np

from . import camera as cam
from . import scenes, textures as tx
from .layout import MATERIAL_DTYPE, MESHLET_DTYPE

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SPONZA_FIXTURE = os.path.join(ROOT, "tests", "golden", "sponza_lowpoly_scene.npz")
GOLDEN = os.path.join(ROOT, "tests", "golden", "bench_configs.json")
NUM_VIEWS = 64


@dataclass
class Workload:
    name: str
    description: str
    scene: scenes.SceneData
    resolve: bool              # vis-buffer + resolve, or vis-buffer only
    fused_cull: bool = False
    cameras: list | None = None    # view batch (None: the scene's own camera)


def save_sponza_fixture(scene: scenes.SceneData, path: str = SPONZA_FIXTURE) -> None:
    c = scene.camera
    np.savez_compressed(path, meshlets=np.frombuffer(np.ascontiguousarray(scene.meshlets).tobytes(), dtype=np.uint8),
                        nodes=np.array([[n.meshlet_offset, n.meshlet_count] for n in scene.nodes], dtype=np.int64),
                        models=np.stack([n.model for n in scene.nodes]).astype(np.float32),
                        camera=np.array([*c.position, c.euler[0], c.euler[1], c.fov, c.near_z], dtype=np.float64))


def load_sponza_fixture(width: int = 1920, height: int = 1080, path: str = SPONZA_FIXTURE) -> scenes.SceneData:
    z = np.load(path)
    meshlets = np.frombuffer(z["meshlets"].tobytes(), dtype=MESHLET_DTYPE).copy()
    nodes = [scenes.DrawNode(int(a), int(b), np.ascontiguousarray(m)) for (a, b), m in zip(z["nodes"], z["models"])]
    c = z["camera"]
    camera = cam.Camera(position=c[0:3], euler=(c[3], c[4]), fov_deg=c[5], aspect=width / height, near_z=c[6])
    scene = scenes.SceneData("Sponza_LowPoly", meshlets, nodes, camera, width, height)
    scene.lights = scenes.default_light()
    return scene


def _textured_c4(width: int, height: int) -> scenes.SceneData:
    scene = scenes.instanced_scene(width=width, height=height, bake_transforms=True)
    scene.meshlets["MaterialId"] = 0
    scene.materials = np.zeros(1, dtype=MATERIAL_DTYPE)
    scene.materials["TextureId"] = 0
    scene.materials["AlphaCutoff"] = 255
    scene.textures = [tx.procedural_material_texture(1024, seed=2)]
    scene.lights = scenes.default_light()
    return scene


def build(name: str) -> Workload:
    if name == "c1_knot":
        return Workload(name, "C1 stand-in: 72,000-triangle textured torus knot, 2 materials, 1920x1080, vis-buffer + resolve",
                        scenes.torus_knot_scene(), True)
    if name == "c1_sponza":
        return Workload(name, "C1 glTF: reference asset Sponza_LowPoly.gltf (63,084 triangles, 2 nodes), RasterBench.cpp:68 camera, 1920x1080, vis-buffer + resolve",
                        load_sponza_fixture(), True)
    if name == "c2_grid":
        return Workload(name, "C2: procedural 999,600-triangle meshlet grid, 1920x1080, depth + triangle id only", scenes.grid_scene(), False)
    if name == "c3_knot":
        return Workload(name, "C3 stand-in: 259,200 textured triangles incl. a double-sided alpha-tested material, 2560x1440, vis-buffer + resolve",
                        scenes.torus_knot_scene(600, 216, 2560, 1440, tex_size=512, alpha_material=True), True)
    if name in ("c4_views", "c5_views"):
        w, h = (1920, 1080) if name == "c4_views" else (2048, 2048)
        scene = _textured_c4(w, h)
        cams = scenes.orbit_cameras(scene, NUM_VIEWS)
        what = ("C4 geometry x 64 views" if name == "c4_views" else "C5")
        return Workload(name, f"{what}: batch of {NUM_VIEWS} orbit-camera views of the procedural {scene.num_triangles:,}-triangle instanced meshlet scene "
                              f"({len(scene.meshlets):,} meshlets, {len(scene.nodes)} DrawMeshlets calls per view, frustum cull fused into the mesh kernel), "
                              f"{w}x{h}, per view clear + cull + vis-buffer + resolve (1 material, 1024^2 2-layer texture, 1 directional light) + GetPixels",
                        scene, True, fused_cull=True, cameras=cams)
    raise KeyError(name)


def view_draws(rast, wl: Workload, view: int | None = None) -> list:
    """The DrawMeshlets calls of one frame of the workload (one per node), as dicts for Rasterizer.draw_batch / create_batch."""
    scene = wl.scene
    if view is not None and wl.cameras is not None:
        scene.camera = wl.cameras[view]
    proj, vm = scene.view_proj()
    draws = []
    for n in scene.nodes:
        d = dict(offset=n.meshlet_offset, count=n.meshlet_count, object_to_clip=scene.object_to_clip(n))
        if wl.fused_cull:
            d["planes"] = rast.frustum_planes(proj, vm, n.model)
        draws.append(d)
    return draws


def view_uniforms(wl: Workload, view: int | None = None) -> dict:
    scene = wl.scene
    if view is not None and wl.cameras is not None:
        scene.camera = wl.cameras[view]
    return scenes.resolve_uniforms(scene, scene.nodes[0])


def algorithmic_bytes(wl: Workload, meshlets_tested: int, meshlets_visible: int) -> dict:
    """compulsory DRAM bytes per frame of each stage (bound sphere per tested meshlet, the 1,216 hot bytes
    per surviving one; 8 B/px of vis-buffer; 12 B/px for the resolve pass)."""
    px = wl.scene.width * wl.scene.height
    return {"mesh": 16 * meshlets_tested + 1216 * meshlets_visible, "raster": 8 * px, "resolve": 12 * px}
