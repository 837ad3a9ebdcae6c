#!/usr/bin/env python
"""BASELINE config C5: a batch of 64 camera views of the 10 M-triangle instanced scene at 2048x2048 (vis-buffer + resolve),
dealt round-robin to the ranks, every finished composite gathered on rank 0 over NVLink peer memory.

    python tools/c5_views.py                                     # 1 GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29520 tools/c5_views.py

Not the contract bench (bench.py is); this measures the other multi-GPU workload BASELINE.json names and prints one
JSON line on rank 0: views/s, triangles/s, and a checksum of the gathered composites. Each view = fused frustum
cull + mesh/raster of the 122 instance draws + resolve + GetPixels into rank 0's memory; F contexts in flight per GPU."""
from __future__ import annotations

import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from glimpsw_b200 import api, scenes, sharding, textures as tx  # noqa: E402
from glimpsw_b200.layout import MATERIAL_DTYPE  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--views", type=int, default=64)
    ap.add_argument("--size", type=int, default=2048)
    ap.add_argument("--rounds", type=int, default=3, help="timed passes over the whole batch")
    ap.add_argument("--in-flight", type=int, default=3)
    ap.add_argument("--mesh-blocks", type=int, default=2)
    args = ap.parse_args()
    import torch
    import torch.distributed as dist

    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    sys.stdout.flush()
    saved = os.dup(1)
    os.dup2(2, 1)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    scene = scenes.instanced_scene(width=args.size, height=args.size)
    scene.meshlets["MaterialId"] = 0
    scene.materials = np.zeros(1, dtype=MATERIAL_DTYPE)
    scene.materials["AlphaCutoff"] = 255
    scene.textures = [tx.procedural_material_texture(1024, seed=2)]
    scene.lights = scenes.default_light()
    cams = scenes.orbit_cameras(scene, args.views)
    mine = sharding.views_for_rank(args.views, rank, world)
    tris = scene.num_triangles

    F = max(1, min(args.in_flight, len(mine)))
    ctxs = []
    for _ in range(F):
        r = api.Rasterizer(local, fused_frustum_cull=True)
        st = torch.cuda.Stream()
        r.set_stream(st.cuda_stream)
        r.set_mesh_occupancy(args.mesh_blocks if F > 1 else 4)
        ctxs.append(dict(rast=r, stream=st, fb=r.create_framebuffer(args.size, args.size),
                         scene=r.upload_scene(scene.meshlets, scene.materials, scene.textures, scene.lights),
                         resolved=torch.cuda.Event(), copied=None))
    # per view: the draw batch (122 nodes, each with its planes) and the resolve uniforms, built once
    per_view = []
    for v in mine:
        scene.camera = cams[v]
        proj, view = scene.view_proj()
        r0 = ctxs[0]["rast"]
        batch = r0.make_batch([dict(offset=n.meshlet_offset, count=n.meshlet_count, object_to_clip=scene.object_to_clip(n),
                                    planes=r0.frustum_planes(proj, view, n.model)) for n in scene.nodes])
        per_view.append((batch, api.Rasterizer.make_uniforms(**scenes.resolve_uniforms(scene, scene.nodes[0]))))

    comm, coll = torch.cuda.Stream(), torch.cuda.Stream()
    slots = 4
    peers = sharding.PeerComposites(args.size, args.size, rank, world, slots=slots) if world > 1 else None
    local_out = [torch.empty((args.size, args.size), dtype=torch.int32, device="cuda") for _ in range(slots)] if world == 1 else None
    done = [torch.cuda.Event() for _ in range(slots)]
    state = dict(k=0)

    def render(i):
        c = ctxs[i % F]
        batch, uni = per_view[i % len(per_view)]
        c["fb"].clear(0xFF000000, 0.0)
        c["rast"].draw_prebuilt(c["fb"], c["scene"], batch)
        if c["copied"] is not None:
            c["stream"].wait_event(c["copied"])
        c["rast"].resolve_prebuilt(c["fb"], c["scene"], uni)
        slot = state["k"] % slots
        state["k"] += 1
        c["resolved"].record(c["stream"])
        comm.wait_event(c["resolved"])
        if peers is not None:
            peers.send(c["fb"], slot, comm)
        else:
            comm.wait_event(done[slot])
            c["fb"].get_pixels_device(0, local_out[slot].data_ptr(), cuda_stream=comm.cuda_stream)
        if c["copied"] is None:
            c["copied"] = torch.cuda.Event()
        c["copied"].record(comm)
        if peers is not None:
            peers.collect(c["rast"], slot, coll)
            done[slot].record(coll if rank == 0 else comm)
        else:
            done[slot].record(comm)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    for i in range(len(mine)):                      # warm-up: one pass over this rank's views
        render(i)
    barrier()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record(ctxs[0]["stream"])
    for c in ctxs[1:]:
        c["stream"].wait_event(t0)
    n_frames = args.rounds * len(mine)
    for i in range(n_frames):
        render(i)
    for c in ctxs[1:]:
        ev = torch.cuda.Event()
        ev.record(c["stream"])
        ctxs[0]["stream"].wait_event(ev)
    for ev in done:
        ctxs[0]["stream"].wait_event(ev)
    t1.record(ctxs[0]["stream"])
    barrier()
    ms = t0.elapsed_time(t1)
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    counters = ctxs[0]["rast"].counters()
    if rank == 0:
        total_views = args.rounds * args.views
        last = peers.buf[(state["k"] - 1) % slots] if peers is not None else local_out[(state["k"] - 1) % slots]
        line = {"workload": f"C5: {args.views} views of the {tris:,}-triangle instanced scene at {args.size}x{args.size}, vis-buffer + resolve, "
                            f"views dealt round-robin to {world} GPU(s), composites gathered on rank 0",
                "n_gpus": world, "views_per_s": round(total_views / (ms * 1e-3), 1), "ms_per_view_per_gpu": round(ms / n_frames, 4),
                "Gtri_per_s_submitted": round(total_views * tris / (ms * 1e-3) / 1e9, 2), "timed_views": total_views,
                "frames_in_flight": F, "composite_bytes_per_view": args.size * args.size * 4,
                "rank0_counters_ctx0": {k: counters[k] for k in ("TrianglesProcessed", "TrianglesRasterized")},
                "last_composite_xor": int(np.bitwise_xor.reduce(last.cpu().numpy().view(np.uint32).reshape(-1)))}
        sys.stdout.flush()
        os.dup2(saved, 1)
        print(json.dumps(line), flush=True)
    for c in ctxs:
        c["rast"].destroy()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
