#!/usr/bin/env python
"""bench.py — headline benchmark of the B200-native meshlet raster path (contract: see task prompt).

A "step" is one frame of the hot path over one batch of synthetic input: Framebuffer::Clear ->
Rasterizer::DrawMeshlets (vis-buffer) -> ShadingContext::Resolve, i.e. BASELINE.json's metric
"Mtri/s and frames/s @1080p vis-buffer+resolve". Workload at N=1: BASELINE config C2 geometry
(procedural 999,600-triangle meshlet grid, 1920x1080) bound to a procedural two-layer material so the
resolve pass samples textures. `value` = scene triangles submitted per second with the scene resident
in HBM; `e2e` = the same frame through the C ABI with HOST buffers (meshlets uploaded from pinned
memory and the resolved image read back every step).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--mode binned|direct]

N>1 (torchrun, one rank per GPU): views are independent units, so every rank renders its own camera of
the same scene (weak scaling, no data-path collective); the resolved 1080p composites are gathered to
rank 0 with NCCL inside the timed region (BASELINE config C5's composite gather).
--impl reference: the CPU restatement of the reference (oracle/baseline_mt.cpp, all host threads) runs
the same frames on rank 0; the upstream binary cannot be built in this image.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from glimpsw_b200 import scenes, textures as tx  # noqa: E402
from glimpsw_b200.layout import MATERIAL_DTYPE  # noqa: E402

METRIC = "Mtri/s @1080p vis-buffer+resolve"
UNIT = "Mtri/s"


def build_workload(rank: int = 0):
    """C2 geometry + one material (procedural 1024^2 two-layer texture) + the default directional light."""
    scene = scenes.grid_scene(material_id=0)
    scene.materials = np.zeros(1, dtype=MATERIAL_DTYPE)
    scene.materials["TextureId"] = 0
    scene.materials["AlphaCutoff"] = 255
    scene.textures = [tx.procedural_material_texture(1024, seed=2)]
    scene.lights = scenes.default_light()
    if rank:   # every rank renders its own view of the same scene: a seeded sub-millimetre camera offset, so the
        # views differ (different sub-pixel coverage) while the per-GPU work stays the same (weak scaling)
        r = scenes.rand01(100 + rank, 3)
        scene.camera.position = scene.camera.position + (r - 0.5) * 0.004
    return scene


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons while the timed region runs (profiling recipe)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu, self.rows, self.proc = gpu_index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self) -> dict:
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
            except Exception:
                continue
            for name, col in (("hw_slowdown", 5), ("hw_thermal_slowdown", 6), ("sw_thermal_slowdown", 7), ("sw_power_cap", 8)):
                if len(r) > col and r[col].lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def algorithmic_bytes(scene, counters):
    """compulsory DRAM bytes per frame of each stage."""
    m_tested = len(scene.meshlets)
    m_visible = m_tested          # C2: every meshlet is inside the frustum, no cull bitmap
    px = scene.width * scene.height
    return {"mesh": 16 * m_tested + 1216 * m_visible, "raster": 8 * px, "resolve": 12 * px}


def run_ours(args):
    import torch
    import torch.distributed as dist
    from glimpsw_b200 import api

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product has no CPU path (use --impl reference for the CPU baseline)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    scene = build_workload(rank)
    node = scene.nodes[0]
    tris = scene.num_triangles
    rast = api.Rasterizer(local_rank, enable_binning=(args.mode == "binned"))
    # One in-order, non-default stream for our kernels, the torch timing events and NCCL. (torch's default
    # stream has handle 0, which swrb_device_set_stream reads as "use your own stream": events recorded
    # there would not be ordered with the kernels.)
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    assert stream.cuda_stream != 0
    rast.set_stream(stream.cuda_stream)
    gscene = rast.upload_scene(scene.meshlets, scene.materials, scene.textures, scene.lights)
    fb = rast.create_framebuffer(scene.width, scene.height)
    batch = rast.make_batch([dict(offset=node.meshlet_offset, count=node.meshlet_count, object_to_clip=scene.object_to_clip(node))])
    uni = scenes.resolve_uniforms(scene, node)
    # N > 1: the resolved composite of every view is gathered to rank 0 over NVLink. Default: the de-tile
    # kernel of each rank stores straight into rank 0's buffer (peer memory, glimpsw_b200.sharding.PeerComposites)
    # with device-side ready/ack signals; fallback (--gather nccl): NCCL gather on a side stream. Either way
    # the exchange of frame k overlaps the render of frame k+1 (double-buffered) and the tail is timed.
    from glimpsw_b200 import sharding
    comm = torch.cuda.Stream() if world > 1 else None
    gather_kind = "none"
    peers = None
    if world > 1:
        gather_kind = args.gather
        if gather_kind == "p2p":
            try:
                peers = sharding.PeerComposites(scene.height, scene.width, rank, world)
            except Exception as exc:          # symmetric memory unavailable: use the collective
                if rank == 0:
                    print(f"bench.py: peer-memory gather unavailable ({exc!r}); using NCCL gather", file=sys.stderr)
                gather_kind = "nccl"
        flag = torch.tensor([1 if gather_kind == "p2p" else 0], device="cuda")
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)       # all ranks must agree
        if int(flag.item()) == 0 and gather_kind == "p2p":
            gather_kind, peers = "nccl", None
    composites = [torch.empty((scene.height, scene.width), dtype=torch.int32, device="cuda") for _ in range(2)]
    gathered = [[torch.empty_like(composites[0]) for _ in range(world)] for _ in range(2)] if (gather_kind == "nccl" and rank == 0) else [None, None]
    rendered = [torch.cuda.Event() for _ in range(2)]
    gather_done = [torch.cuda.Event() for _ in range(2)]
    state = {"k": 0}

    resolved = torch.cuda.Event()

    def frame():
        fb.clear(0xFF000000, 0.0)
        rast.draw_prebuilt(fb, gscene, batch)
        exchange = world > 1 and gather_kind != "none"
        if exchange and state["k"] > 0:
            stream.wait_event(gather_done[(state["k"] - 1) & 1])     # the previous frame's de-tile has read layer 0
        rast.resolve(fb, gscene, **uni)
        if exchange:
            slot = state["k"] & 1
            state["k"] += 1
            resolved.record(stream)
            with torch.cuda.stream(comm):                            # everything below runs beside the next frame's draw
                comm.wait_event(resolved)
                if peers is not None:
                    peers.before_write(slot, comm)
                    fb.get_pixels_device(0, peers.dst_ptr(slot), cuda_stream=comm.cuda_stream)   # GetPixels straight into rank 0's memory
                    peers.after_write(slot, comm)
                    peers.collect(slot, comm)
                else:
                    fb.get_pixels_device(0, composites[slot].data_ptr(), cuda_stream=comm.cuda_stream)
                    dist.gather(composites[slot], gathered[slot], dst=0)
                gather_done[slot].record(comm)

    def barrier():
        if world > 1:
            comm.synchronize()
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        frame()
    barrier()

    # ---- timed region: exactly K steps; L2 is flushed before each step, outside the step's events
    sampler = ClockSampler(local_rank)
    sampler.start()
    launches0 = rast.launch_count()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    tail = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
    barrier()
    t_wall0 = time.perf_counter()
    for b, e in ev:
        rast.flush_l2()
        b.record(stream)
        frame()
        e.record(stream)
    tail[0].record(stream)
    if world > 1:                                                 # the last gathers are part of the job
        stream.wait_event(gather_done[0])
        stream.wait_event(gather_done[1])
    tail[1].record(stream)
    barrier()
    t_wall = time.perf_counter() - t_wall0
    launches = rast.launch_count() - launches0
    clocks = sampler.stop()
    step_ms = [b.elapsed_time(e) for b, e in ev]
    total_ms = float(sum(step_ms)) + tail[0].elapsed_time(tail[1])
    if world > 1:
        t = torch.tensor([total_ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms = float(t.item())
    ms_per_step = total_ms / args.steps
    value = tris * world / (ms_per_step * 1e-3) / 1e6

    # ---- per-stage device times (events around every kernel group, one extra frame) for the roofline
    rast.enable_stage_timing(True)
    stage_acc = {}
    reps = 5
    for _ in range(reps):
        rast.flush_l2()
        frame()
        for k, (us, n) in rast.stage_times_us().items():
            a = stage_acc.setdefault(k, [0.0, 0])
            a[0] += us / reps
            a[1] = n
    rast.enable_stage_timing(False)
    rast.reset_counters()
    frame()
    counters = rast.counters()
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak_gbs = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
    abytes = algorithmic_bytes(scene, counters)
    stages = {}
    for k in ("mesh", "bin", "raster", "resolve"):
        us = stage_acc.get(k, [0.0, 0])[0]
        if us <= 0:
            continue
        bytes_k = abytes.get(k)
        stages[k] = {"us": round(us, 2), "launches": stage_acc[k][1]}
        if bytes_k:
            stages[k].update({"algorithmic_bytes": bytes_k, "GBs": round(bytes_k / us / 1e3, 1), "frac": round(bytes_k / us / 1e3 / peak_gbs, 4)})
    dom = max((k for k in stages if "GBs" in stages[k]), key=lambda k: stages[k]["us"])
    roofline = {"bound": "hbm", "kernel": {"mesh": "k_mesh_setup", "raster": "k_tile_raster" if args.mode == "binned" else "k_raster_direct",
                                            "resolve": "k_resolve"}[dom],
                "achieved": stages[dom]["GBs"], "peak": peak_gbs, "unit": "GB/s", "frac": stages[dom]["frac"],
                "traffic": None, "peak_source": peak_src, "algorithmic_bytes": stages[dom]["algorithmic_bytes"],
                "note": "stage time from CUDA events on the launching stream, L2 flushed before each frame"}

    # ---- e2e: the same frame through the C ABI with HOST buffers (rank-local; max over ranks)
    host_meshlets = rast.alloc_pinned(scene.meshlets.shape, scene.meshlets.dtype)
    host_meshlets[...] = scene.meshlets
    host_image = rast.alloc_pinned((scene.height, scene.width), np.uint32)

    def frame_e2e():
        gscene.update_meshlets(host_meshlets, 0)          # H2D: 1728 B x meshlets, from pinned memory
        fb.clear(0xFF000000, 0.0)
        rast.draw_prebuilt(fb, gscene, batch)
        rast.resolve(fb, gscene, **uni)
        fb.get_pixels(0, host_image)                      # D2H: resolved RGBA8 image (synchronises)

    for _ in range(3):
        frame_e2e()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        frame_e2e()
    barrier()
    e2e_s = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    e2e_value = tris * world * args.steps / e2e_s / 1e6
    checksum = int(np.bitwise_xor.reduce(host_image.reshape(-1)))

    # ---- CPU baseline (rank 0, N=1 only): the reference restatement on the host cores, bounded sample
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu = cpu_frames(scene, node, uni, budget_s=args.cpu_budget, max_frames=40)

    if rank == 0:
        line = {
            "metric": METRIC, "value": round(value, 2), "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": round(ms_per_step, 5), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32+i32 (28.4 fixed-point coverage, fp32 depth/shading)", "data": "synthetic",
            "config": {"workload": "C2: procedural 999,600-triangle meshlet grid (10,200 meshlets), 1920x1080, "
                                   "clear + vis-buffer (depth + triangle id) + resolve (1 material, 1024^2 2-layer texture, 1 directional light)",
                       "triangles_per_frame": tris, "meshlets": len(scene.meshlets), "mode": args.mode,
                       "parallelism": f"view-parallel x{world}" + ({"p2p": ", composites stored by each rank's de-tile kernel straight into rank 0's memory over NVLink (peer memory + device-side signals, double-buffered, tail included)", "nccl": ", composites gathered to rank 0 with NCCL on a side stream (double-buffered, tail included)", "none": ""}[gather_kind]),
                       "l2": "evicted (256 MB write + 256 MB read) before every timed step, outside the step's events",
                       "timing": "CUDA events per step on the launching stream, summed; max over ranks"},
            "frames_per_s": round(world / (ms_per_step * 1e-3), 1),
            "step_ms_min_median_max": [round(float(np.min(step_ms)), 5), round(float(np.median(step_ms)), 5), round(float(np.max(step_ms)), 5)],
            "wall_ms_per_step_incl_flush": round(t_wall / args.steps * 1e3, 4),
            "clocks": clocks, "gpu_launches": int(launches),
            "e2e": {"value": round(e2e_value, 2), "unit": UNIT, "h2d_bytes_per_step": int(scene.meshlets.nbytes + 512),
                    "d2h_bytes_per_step": int(scene.width * scene.height * 4), "ms_per_step": round(e2e_s / args.steps * 1e3, 4),
                    "note": "meshlets re-uploaded from pinned host memory and the resolved image read back every step; wall clock"},
            "roofline": roofline, "stages": stages,
            "counters": {k: counters[k] for k in ("TrianglesProcessed", "TrianglesRasterized", "TrianglesClipped", "BinQueueFlushes")},
            "image_xor": checksum,
        }
        if cpu:
            line["cpu_baseline"] = cpu
        print(json.dumps(line), flush=True)
    rast.destroy()
    if world > 1:
        dist.destroy_process_group()


def cpu_frames(scene, node, uni, budget_s: float, max_frames: int, threads: int = 0):
    """Times the CPU restatement of the reference on the same frame (clear + draw + resolve)."""
    from oracle import orc
    orc.build()
    base = orc.Baseline(threads)
    fb = orc.Framebuffer(scene.width, scene.height)
    m = scene.object_to_clip(node)
    times = []
    t_start = time.perf_counter()
    while len(times) < max_frames and (time.perf_counter() - t_start < budget_s or len(times) < 3):
        t0 = time.perf_counter()
        base.clear(fb, 0xFF000000, 0.0)
        base.draw_meshlets(fb, scene.meshlets, node.meshlet_offset, node.meshlet_count, m, materials=scene.materials)
        base.resolve(fb, scene.meshlets, scene.materials, scene.textures, scene.lights, **uni)
        times.append(time.perf_counter() - t0)
    timed = times[1:] if len(times) > 1 else times
    med = float(np.median(timed))
    out = {"value": round(scene.num_triangles / med / 1e6, 2), "unit": UNIT, "cores": base.threads, "kind": "port",
           "isa": "avx512 (4x4-fragment inner loop)" if base.avx512 else "scalar (no AVX-512 on this host)",
           "sample": f"{len(timed)} full frames of the same workload after 1 warm-up (median {med * 1e3:.1f} ms/frame)",
           "ms_per_step": round(med * 1e3, 3), "cpu": cpu_model()}
    base.close()
    return out


def cpu_model() -> str:
    try:
        for line in open("/proc/cpuinfo"):
            if line.startswith("model name"):
                return line.split(":", 1)[1].strip()
    except Exception:
        pass
    return "unknown"


def run_reference(args):
    """--impl reference: the CPU restatement of GLimpSW's path on the host cores (rank 0 only)."""
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if rank != 0:
        return
    scene = build_workload(0)
    node = scene.nodes[0]
    uni = scenes.resolve_uniforms(scene, node)
    from oracle import orc
    orc.build()
    base = orc.Baseline(0)
    fb = orc.Framebuffer(scene.width, scene.height)
    m = scene.object_to_clip(node)

    def frame():
        base.clear(fb, 0xFF000000, 0.0)
        base.draw_meshlets(fb, scene.meshlets, node.meshlet_offset, node.meshlet_count, m, materials=scene.materials)
        base.resolve(fb, scene.meshlets, scene.materials, scene.textures, scene.lights, **uni)

    for _ in range(args.warmup):
        frame()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        frame()
    dt = time.perf_counter() - t0
    ms = dt / args.steps * 1e3
    value = scene.num_triangles / (ms * 1e-3) / 1e6
    cpu = {"value": round(value, 2), "unit": UNIT, "cores": base.threads, "kind": "port",
           "isa": "avx512" if base.avx512 else "scalar",
           "sample": f"{args.steps} full frames of the same workload (clear + draw + resolve), all host threads", "cpu": cpu_model()}
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": round(value, 2), "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": round(ms, 4), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32+i32", "data": "synthetic",
        "config": {"workload": "C2: procedural 999,600-triangle meshlet grid, 1920x1080, clear + vis-buffer + resolve",
                   "note": "CPU restatement of GLimpSW's binned AVX-512 path (oracle/baseline_mt.cpp); the upstream binary needs clang + CPM deps and cannot be built here"},
        "cpu_baseline": cpu,
        "e2e": {"value": round(value, 2), "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }), flush=True)
    base.close()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--mode", default="binned", choices=["binned", "direct"])
    ap.add_argument("--gather", default="p2p", choices=["p2p", "nccl", "none"], help="N>1: how composites reach rank 0 (none = diagnostic: no exchange)")
    ap.add_argument("--cpu-budget", type=float, default=15.0, help="seconds of CPU-baseline frames at N=1")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
