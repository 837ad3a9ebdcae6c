#!/usr/bin/env python
"""bench.py — headline benchmark of the B200-native meshlet raster path (contract: see task prompt).

A "step" is one frame of the hot path over one batch of synthetic input: Framebuffer::Clear ->
Rasterizer::DrawMeshlets (vis-buffer) -> ShadingContext::Resolve, i.e. BASELINE.json's metric
"Mtri/s and frames/s @1080p vis-buffer+resolve". Workload at N=1: BASELINE config C2 geometry
(procedural 999,600-triangle meshlet grid, 1920x1080) bound to a procedural two-layer material so the
resolve pass samples textures.

  value   whole-job throughput, scene resident in HBM: K frames submitted round-robin to F (default 6)
          independent render contexts (own stream, framebuffer, work buffers; mesh kernel sized to one block per SM) so that the issue-bound resolve
          of one frame overlaps the latency-bound mesh/raster kernels of the next. Inputs are larger than L2:
          the contexts rotate over 8 copies of the 17.6 MB meshlet buffer (141 MB > 126 MB L2), so no frame
          finds its meshlets cached. Timed with CUDA events on the launching streams; max over ranks.
  latency the same frame strictly serialised on one stream with the L2 evicted before every step
          (`latency_ms_per_frame`, `stages`, `roofline` come from this mode).
  e2e     the same metric through the C ABI with HOST buffers: every step uploads the meshlets from pinned
          host memory (H2D), renders, and reads the resolved image back (D2H); the F contexts keep the copies
          and the kernels of different steps overlapped.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--mode binned|direct]

N>1 (torchrun, one rank per GPU): views are independent units, so every rank renders its own camera of the
same scene (weak scaling, no data-path collective); the resolved 1080p composites are collected on rank 0:
each rank's de-tile kernel stores straight into rank 0's memory over NVLink (--gather p2p, default) or NCCL
gather (--gather nccl), on a side stream, double-buffered; the tail is inside the timed region.
--impl reference: the CPU restatement of the reference (oracle/baseline_mt.cpp, all host threads) runs the same
frames on rank 0; the upstream binary cannot be built in this image.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from types import SimpleNamespace

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from glimpsw_b200 import scenes, textures as tx  # noqa: E402
from glimpsw_b200.layout import MATERIAL_DTYPE  # noqa: E402

METRIC = "Mtri/s @1080p vis-buffer+resolve"
UNIT = "Mtri/s"
WORKLOAD = ("C2: procedural 999,600-triangle meshlet grid (10,200 meshlets), 1920x1080, "
            "clear + vis-buffer (depth + triangle id) + resolve (1 material, 1024^2 2-layer texture, 1 directional light)")
SCENE_COPIES = 8          # x 17.6 MB of meshlets = 141 MB > 126 MB L2
SLOTS = 6                 # N > 1: composite buffers in flight per rank


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
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "50",
                                          "-i", str(self.gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
            time.sleep(0.12)          # let the first sample land before the (short) timed region starts
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self) -> dict:
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.12)
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


def algorithmic_bytes(scene):
    """compulsory DRAM bytes per frame of each stage."""
    m_tested = len(scene.meshlets)
    m_visible = m_tested          # C2: every meshlet is inside the frustum, no cull bitmap
    px = scene.width * scene.height
    return {"mesh": 16 * m_tested + 1216 * m_visible, "raster": 8 * px, "resolve": 12 * px}


def bind_to_gpu_numa_node(gpu_index: int) -> None:
    """Pins this rank (and with it the pinned host buffers it allocates) to the CPU socket its GPU hangs off, so that
    the e2e copies of 8 ranks do not all cross the inter-socket link. Best effort: silently does nothing when the
    topology cannot be read."""
    try:
        out = subprocess.run(["nvidia-smi", "--query-gpu=pci.bus_id", "--format=csv,noheader", "-i", str(gpu_index)],
                             capture_output=True, text=True, timeout=20).stdout.strip().lower()
        bus = out[-12:] if len(out) >= 12 else out                      # 00000000:1b:00.0 -> 0000:1b:00.0
        node = int(open(f"/sys/bus/pci/devices/{bus}/numa_node").read())
        if node < 0:
            return
        cpus = set()
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
    except Exception:
        pass


def run_ours(args):
    import torch
    import torch.distributed as dist
    from glimpsw_b200 import api, sharding

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product has no CPU path (use --impl reference for the CPU baseline)")
    full_affinity = os.sched_getaffinity(0)
    bind_to_gpu_numa_node(local_rank)
    torch.cuda.set_device(local_rank)
    # stdout carries exactly one JSON line: anything libraries print there meanwhile (NCCL's version banner under
    # NCCL_DEBUG=VERSION, ...) is routed to stderr until the line is written
    sys.stdout.flush()
    saved_stdout = os.dup(1)
    os.dup2(2, 1)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    scene = build_workload(rank)
    node = scene.nodes[0]
    tris = scene.num_triangles
    uni = scenes.resolve_uniforms(scene, node)
    uni_c = api.Rasterizer.make_uniforms(**uni)       # the C uniform block, built once
    F = max(1, args.in_flight)
    copies = (SCENE_COPIES + F - 1) // F

    # ---- F independent render contexts on this GPU. Streams are explicit non-default torch streams (torch's
    # default stream has handle 0, which swrb_device_set_stream reads as "use your own stream").
    ctxs = []
    for i in range(F):
        r = api.Rasterizer(local_rank, enable_binning=(args.mode == "binned"), resolve_cache=not args.no_resolve_cache)
        st = torch.cuda.Stream()
        assert st.cuda_stream != 0
        r.set_stream(st.cuda_stream)
        r.set_mesh_occupancy(args.mesh_blocks if F > 1 else 4)     # several contexts in flight: leave half of each SM to the others' resolve
        c = SimpleNamespace(rast=r, stream=st, fb=r.create_framebuffer(scene.width, scene.height),
                            scenes=[r.upload_scene(scene.meshlets, scene.materials, scene.textures, scene.lights) for _ in range(copies)],
                            batch=r.make_batch([dict(offset=node.meshlet_offset, count=node.meshlet_count,
                                                     object_to_clip=scene.object_to_clip(node))]),
                            uses=0, resolved=torch.cuda.Event(), copied=None)
        ctxs.append(c)

    # ---- N > 1: composites go to rank 0 over NVLink, on a side stream, double-buffered
    comm = torch.cuda.Stream() if world > 1 else None
    gather_kind, peers = "none", None
    if world > 1:
        gather_kind = args.gather
        if gather_kind == "p2p":
            try:
                with torch.cuda.stream(comm):
                    peers = sharding.PeerComposites(scene.height, scene.width, rank, world, slots=SLOTS)
            except Exception as exc:          # symmetric memory unavailable: use the collective
                if rank == 0:
                    print(f"bench.py: peer-memory gather unavailable ({exc!r}); using NCCL gather", file=sys.stderr)
                gather_kind = "nccl"
        flag = torch.tensor([1 if gather_kind == "p2p" else 0], device="cuda")
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)       # all ranks must agree
        if int(flag.item()) == 0 and gather_kind == "p2p":
            gather_kind, peers = "nccl", None
    exchange = world > 1 and gather_kind != "none"
    composites = [torch.empty((scene.height, scene.width), dtype=torch.int32, device="cuda") for _ in range(SLOTS)]
    gathered = [[torch.empty_like(composites[0]) for _ in range(world)] for _ in range(SLOTS)] if (gather_kind == "nccl" and rank == 0) else [None] * SLOTS
    gather_done = [torch.cuda.Event() for _ in range(SLOTS)]
    coll = torch.cuda.Stream() if world > 1 else None      # rank 0 waits for its peers here, not on the stream that sends its own view
    state = {"k": 0}

    def frame(c, gscene):
        """Framebuffer::Clear -> DrawMeshlets -> Resolve on context c (+ the composite exchange when N > 1)."""
        c.fb.clear(0xFF000000, 0.0)
        c.rast.draw_prebuilt(c.fb, gscene, c.batch)
        if exchange and c.copied is not None:
            c.stream.wait_event(c.copied)                       # the previous de-tile of this context has read layer 0
        c.rast.resolve_prebuilt(c.fb, gscene, uni_c)
        if exchange:
            slot = state["k"] % SLOTS
            state["k"] += 1
            c.resolved.record(c.stream)
            comm.wait_event(c.resolved)                         # the exchange runs beside the next frames' kernels
            if peers is not None:
                # (every call below names its stream: no torch.cuda.stream() context on this path — at ~70 us per frame
                # the host cost of each torch stream / event call shows up in the multi-GPU step time)
                peers.send(c.fb, slot, comm)                    # GetPixels straight into rank 0's memory (+ slot flow control)
                if c.copied is None:
                    c.copied = torch.cuda.Event()               # per context: a shared per-slot event would be re-recorded by
                c.copied.record(comm)                           # later frames and chain consecutive frames together
                peers.collect(c.rast, slot, coll)               # rank 0 waits for its peers on a third stream
                gather_done[slot].record(coll if rank == 0 else comm)
            else:
                with torch.cuda.stream(comm):
                    comm.wait_event(gather_done[slot])
                    c.fb.get_pixels_device(0, composites[slot].data_ptr(), cuda_stream=comm.cuda_stream)
                    if c.copied is None:
                        c.copied = torch.cuda.Event()
                    c.copied.record(comm)
                    dist.gather(composites[slot], gathered[slot], dst=0)
                    gather_done[slot].record(comm)

    def frame_rr(k):
        c = ctxs[k % F]
        frame(c, c.scenes[(k // F) % copies])

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    for k in range(max(args.warmup, 3) * F):
        frame_rr(k)
    barrier()

    # ---- timed region (throughput): exactly K frames, F in flight, inputs larger than L2
    sampler = ClockSampler(local_rank)
    if rank == 0:               # one nvidia-smi poller per job, on the rank that prints: NVML queries from 8 pollers perturb the
        sampler.start()         # very launches they are meant to watch
    launches0 = sum(c.rast.launch_count() for c in ctxs)
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    t_wall0 = time.perf_counter()
    t0.record(ctxs[0].stream)
    for c in ctxs[1:]:
        c.stream.wait_event(t0)
    burst = min(args.steps, 60)                                   # short enough not to fill the driver's launch queue:
    for k in range(burst):                                        # pure host cost of enqueueing a frame
        frame_rr(k)
    t_submit = (time.perf_counter() - t_wall0) / burst
    for k in range(burst, args.steps):
        frame_rr(k)
    for c in ctxs[1:]:
        ev = torch.cuda.Event()
        ev.record(c.stream)
        ctxs[0].stream.wait_event(ev)
    if exchange:                                                  # the last exchanges are part of the job
        for ev_done in gather_done:
            ctxs[0].stream.wait_event(ev_done)
    t1.record(ctxs[0].stream)
    barrier()
    t_wall = time.perf_counter() - t_wall0
    launches = sum(c.rast.launch_count() for c in ctxs) - launches0
    clocks = sampler.stop()
    total_ms = t0.elapsed_time(t1)
    if world > 1:
        t = torch.tensor([total_ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms = float(t.item())
    ms_per_step = total_ms / args.steps
    value = tris * world / (ms_per_step * 1e-3) / 1e6

    # ---- latency mode: one context, strictly serial, L2 evicted before every frame; per-stage times for the roofline
    c0 = ctxs[0]
    rast = c0.rast
    rast.set_mesh_occupancy(4)                                    # a lone frame gets the whole register file
    lat_steps = min(args.steps, 100)
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(lat_steps)]
    save_exchange, exchange = exchange, False
    for b, e in ev:
        rast.flush_l2()
        b.record(c0.stream)
        frame(c0, c0.scenes[0])
        e.record(c0.stream)
    torch.cuda.synchronize()
    lat_ms = [b.elapsed_time(e) for b, e in ev]
    rast.enable_stage_timing(True)
    stage_acc = {}
    reps = 5
    for _ in range(reps):
        rast.flush_l2()
        frame(c0, c0.scenes[0])
        for k, (us, n) in rast.stage_times_us().items():
            a = stage_acc.setdefault(k, [0.0, 0])
            a[0] += us / reps
            a[1] = n
    rast.enable_stage_timing(False)
    rast.reset_counters()
    frame(c0, c0.scenes[0])
    counters = rast.counters()
    draw_stats = rast.draw_stats()
    exchange = save_exchange

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak_gbs = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
    abytes = algorithmic_bytes(scene)
    stages = {}
    for k in ("clear", "mesh", "bin", "raster", "resolve"):
        us = stage_acc.get(k, [0.0, 0])[0]
        if us <= 0:
            continue
        stages[k] = {"us": round(us, 2)}
        bytes_k = abytes.get(k)
        if bytes_k:
            stages[k].update({"algorithmic_bytes": bytes_k, "GBs": round(bytes_k / us / 1e3, 1), "frac": round(bytes_k / us / 1e3 / peak_gbs, 4)})
    traffic = {}
    try:   # DRAM bytes per launch from the committed ncu --set full capture of this workload
        traffic = json.load(open(os.path.join(ROOT, "profiles", "r01_traffic.json")))
    except Exception:
        pass
    dom = max((k for k in stages if "GBs" in stages[k]), key=lambda k: stages[k]["us"])
    roofline = {"bound": "hbm", "kernel": {"mesh": "k_mesh_setup", "raster": "k_tile_raster" if args.mode == "binned" else "k_raster_direct",
                                            "resolve": "k_resolve"}[dom],
                "achieved": stages[dom]["GBs"], "peak": peak_gbs, "unit": "GB/s", "frac": stages[dom]["frac"],
                "traffic": None, "peak_source": peak_src, "algorithmic_bytes": stages[dom]["algorithmic_bytes"],
                "avg_launch_us": stages[dom]["us"],
                "note": "kernel time from CUDA events on the launching stream in latency mode (L2 evicted before each frame); the "
                        "kernel is instruction-issue / load-latency bound, not HBM bound: ~580 warp instructions per pixel for 12 algorithmic bytes (profiles/r01_summary.md)"}
    roofline["traffic"] = traffic.get(roofline["kernel"])
    if roofline["traffic"] is not None:
        roofline["traffic_source"] = "profiles/r01_traffic.json (ncu dram__bytes_read.sum + dram__bytes_write.sum, per launch)"

    # ---- e2e: the same frames through the C ABI with HOST buffers; PCIe-bound, and more than 3 steps in flight only
    # add copy-engine contention (tools/e2e_probe.py), so at most 3 of the contexts take part
    FE = min(F, 3)
    for c in ctxs[:FE]:
        c.host_meshlets = c.rast.alloc_pinned(scene.meshlets.shape, scene.meshlets.dtype)
        c.host_meshlets[...] = scene.meshlets
        c.host_image = c.rast.alloc_pinned((scene.height, scene.width), np.uint32)

    def frame_e2e(k):
        c = ctxs[k % FE]
        if c.uses >= 1:
            c.rast.sync()                                         # this context's previous step (its image is now on the host)
        c.uses += 1
        g = c.scenes[0]
        g.update_meshlets(c.host_meshlets, 0)                     # H2D: 1728 B x meshlets, from pinned memory
        c.fb.clear(0xFF000000, 0.0)
        c.rast.draw_prebuilt(c.fb, g, c.batch)
        c.rast.resolve_prebuilt(c.fb, g, uni_c)
        c.fb.get_pixels_async(0, c.host_image)                    # D2H: resolved RGBA8 image

    for k in range(3 * FE):
        frame_e2e(k)
    barrier()
    t0e = time.perf_counter()
    for k in range(args.steps):
        frame_e2e(k)
    barrier()
    e2e_s = time.perf_counter() - t0e
    if world > 1:
        t = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    e2e_value = tris * world * args.steps / e2e_s / 1e6
    checksum = int(np.bitwise_xor.reduce(ctxs[0].host_image.reshape(-1)))

    # ---- CPU baseline (rank 0, N=1 only): the reference restatement on the host cores, bounded sample
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        os.sched_setaffinity(0, full_affinity)          # the CPU baseline gets every host core, not just the GPU's NUMA node
        cpu = cpu_frames(scene, node, uni, budget_s=args.cpu_budget, max_frames=40)

    if rank == 0:
        line = {
            "metric": METRIC, "value": round(value, 2), "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": round(ms_per_step, 5), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32+i32 (28.4 fixed-point coverage, fp32 depth/shading)", "data": "synthetic",
            "config": {"workload": WORKLOAD,
                       "triangles_per_frame": tris, "meshlets": len(scene.meshlets), "mode": args.mode,
                       "frames_in_flight": F, "mesh_kernel_blocks_per_sm": args.mesh_blocks if F > 1 else 4,
                       "parallelism": f"view-parallel x{world}" + {
                           "p2p": ", composites stored by each rank's de-tile kernel straight into rank 0's memory over NVLink (peer memory + device-side signals, double-buffered, tail included)",
                           "nccl": ", composites gathered to rank 0 with NCCL on a side stream (double-buffered, tail included)", "none": ""}[gather_kind],
                       "l2": f"inputs larger than L2: frames rotate over {F * copies} copies of the {scene.meshlets.nbytes / 1e6:.1f} MB meshlet buffer "
                             f"({F * copies * scene.meshlets.nbytes / 1e6:.0f} MB > 126 MB); latency mode evicts L2 (256 MB write + 256 MB read) before every frame",
                       "timing": "one CUDA-event pair around the K frames on the launching streams (all contexts joined), max over ranks"},
            "frames_per_s": round(world / (ms_per_step * 1e-3), 1),
            # `value` counts submitted triangles; the same rate for the triangles that survive meshlet culling
            # (Rasterizer.cpp:545) and for those that are rasterized (:579)
            "processed_Mtri_s": round(counters["TrianglesProcessed"] * world / (ms_per_step * 1e-3) / 1e6, 2),
            "rasterized_Mtri_s": round(counters["TrianglesRasterized"] * world / (ms_per_step * 1e-3) / 1e6, 2),
            "latency_ms_per_frame": round(float(np.median(lat_ms)), 5),
            "latency_ms_min_max": [round(float(np.min(lat_ms)), 5), round(float(np.max(lat_ms)), 5)],
            "wall_ms_per_step": round(t_wall / args.steps * 1e3, 4),
            "host_submit_ms_per_step": round(t_submit * 1e3, 4),
            "clocks": clocks, "gpu_launches": int(launches),
            "e2e": {"value": round(e2e_value, 2), "unit": UNIT, "h2d_bytes_per_step": int(scene.meshlets.nbytes + 512),
                    "d2h_bytes_per_step": int(scene.width * scene.height * 4), "ms_per_step": round(e2e_s / args.steps * 1e3, 4),
                    "note": f"every step: meshlets H2D from pinned host memory, draw + resolve, resolved image D2H to pinned host memory; "
                            f"{FE} steps in flight on separate streams; wall clock"},
            "roofline": roofline, "stages": stages,
            "counters": {k: counters[k] for k in ("TrianglesProcessed", "TrianglesRasterized", "TrianglesClipped", "BinQueueFlushes")},
            "draw_stats": draw_stats, "image_xor": checksum,
        }
        if cpu:
            line["cpu_baseline"] = cpu
        sys.stdout.flush()
        os.dup2(saved_stdout, 1)
        print(json.dumps(line), flush=True)
    for c in ctxs:
        c.rast.destroy()
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
           "isa": "avx512 (16-lane vertex transform, 16-triangle packet classification + early setup, 4x4-fragment raster loop, resolve; per-triangle edge setup and binning scalar)" if base.avx512 else "scalar (no AVX-512 on this host)",
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
        "config": {"workload": WORKLOAD, "triangles_per_frame": scene.num_triangles, "meshlets": len(scene.meshlets),
                   "note": "CPU restatement of GLimpSW's binned AVX-512 path (oracle/baseline_mt.cpp); the upstream binary needs clang + CPM deps and cannot be built here"},
        "cpu_baseline": cpu,
        "e2e": {"value": round(value, 2), "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }), flush=True)
    base.close()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=300)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--mode", default="binned", choices=["binned", "direct"])
    ap.add_argument("--in-flight", type=int, default=6, help="independent render contexts (frames in flight) per GPU")
    ap.add_argument("--mesh-blocks", type=int, default=1, help="mesh-kernel blocks per SM in the sustained mode (swrb_device_set_mesh_occupancy)")
    ap.add_argument("--gather", default="p2p", choices=["p2p", "nccl", "none"], help="N>1: how composites reach rank 0 (none = diagnostic: no exchange)")
    ap.add_argument("--cpu-budget", type=float, default=15.0, help="seconds of CPU-baseline frames at N=1")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-resolve-cache", action="store_true", help="A/B switch: resolve re-transforms every pixel's corners (SWRB_FLAG_NO_RESOLVE_CACHE)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
