#!/usr/bin/env python
"""bench.py — headline benchmark of the B200-native meshlet raster path (contract: see the task prompt / DESIGN.md §5).

Workload (BASELINE.json configs C4/C5, `glimpsw_b200.workloads.build("c4_views")`): a fixed batch of 64 orbit-camera views
of the procedural 9,994,240-triangle instanced meshlet scene (126,880 meshlets, 122 DrawMeshlets calls per view) at
1920x1080. One view = Framebuffer::Clear -> frustum cull (fused into the mesh kernel) -> Rasterizer::DrawMeshlets per node
(vis-buffer) -> ShadingContext::Resolve -> Framebuffer::GetPixels, the frame loop of Main.cpp:213-275 / RasterBench.cpp:92-106.
A "step" is the WHOLE batch. With N GPUs the views are dealt v mod N (strong scaling): every rank holds the
scene, renders its views, and each finished composite goes to rank 0 (de-tile kernel storing straight into rank 0's memory
over NVLink, --gather p2p; or an NCCL gather, --gather nccl).

  value   submitted triangles of the whole job per second, scene resident in HBM, F render contexts in flight
          per GPU; `processed_` / `rasterized_Mtri_s` count the triangles that survive meshlet culling / reach the rasterizer.
          Inputs are larger than L2 (219 MB of meshlets per scene copy, one copy per context). CUDA events on the
          launching streams, barrier + synchronize on both sides, max over ranks.
  e2e     the same batch through the C ABI with HOST buffers: every step uploads the meshlets from pinned host memory (each
          rank 1/N of them, the rest arrives by NCCL all-gather over NVLink) and every view's resolved image is read back to
          pinned host memory; all render contexts take part (their scene copies refreshed device-to-device), two scene buffers each so the
          next step's upload overlaps this step's rendering.
  parity  before anything is timed every view this rank renders is compared with the committed oracle fixture
          (tests/golden/bench_configs.json: SHA-256 of depth and surface ids, exact) and, where a colour fixture exists,
          with the oracle's resolved image (<= 2/255).
  configs (N = 1) the other BASELINE configs — C1 (torus knot and the reference's Sponza_LowPoly), C2, C3, C5 — each checked
          against its fixture, then timed frame by frame with the L2 evicted: ms/frame, rates, stage table, rooflines.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--mode binned|direct]

--impl reference: the CPU restatement of the reference (oracle/baseline_mt.cpp, all host threads) renders a bounded sample
of the same batch on rank 0. (The reference's own sources build here with g++ only on a scalar stand-in for its Clang vector
types — oracle/_ref, used to PIN the restatement — which would be an unfairly slow denominator)
"""
from __future__ import annotations

import argparse
import hashlib
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

from glimpsw_b200 import workloads  # noqa: E402

METRIC = "Mtri/s @1080p vis-buffer+resolve"
UNIT = "Mtri/s"
SLOTS = 4                 # composite buffers in flight per rank
REF_SAMPLE_VIEWS = [0, 8, 16, 24, 32, 40, 48, 56]   # --impl reference / cpu_baseline: the bounded sample of the batch


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
            time.sleep(0.12)          # let the first sample land before the timed region starts
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


def bind_to_gpu_numa_node(gpu_index: int) -> None:
    """Pins this rank (and with it the pinned host buffers it allocates) to the CPU socket its GPU hangs off, so that
    the e2e copies of 8 ranks do not all cross the inter-socket link. Best effort."""
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


def cpu_model() -> str:
    try:
        for line in open("/proc/cpuinfo"):
            if line.startswith("model name"):
                return line.split(":", 1)[1].strip()
    except Exception:
        pass
    return "unknown"


def load_golden() -> dict:
    try:
        return json.load(open(workloads.GOLDEN))
    except Exception:
        return {}


def load_colour(name: str):
    """(rgb, stride): the oracle's resolved image at pixels [::stride, ::stride]."""
    try:
        z = np.load(os.path.join(ROOT, "tests", "golden", name))
        return z["rgb"], int(z["stride"]) if "stride" in z else 1
    except Exception:
        return None


def peak_hbm():
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    return (float(peaks.get("hbm_gbs", 6650.0)),
            "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)")


def check_view(rast, fb, batch, frame, want: dict | None, colour_name: str | None, resolve: bool) -> dict:
    """Renders one frame twice — vis-buffer only, then the whole frame — and compares with the oracle fixture."""
    out = {"visbuffer": "no fixture", "colour": "no fixture"}
    fb.clear(0xFF000000, 0.0)
    rast.reset_counters()
    rast.draw_prepared(fb, batch)
    depth, ids = fb.download_tiled(1), fb.download_tiled(0)
    c = rast.counters()
    if want:
        ok = (hashlib.sha256(depth.tobytes()).hexdigest() == want["depth_sha256"] and hashlib.sha256(ids.tobytes()).hexdigest() == want["id_sha256"]
              and [c["TrianglesProcessed"], c["TrianglesRasterized"], c["TrianglesClipped"]] == want["counters"])
        out["visbuffer"] = "exact" if ok else "MISMATCH"
    if resolve:
        rast.submit_frame(fb, frame)
        img = fb.get_pixels(0).view(np.uint8).reshape(fb.height, fb.width, 4)[..., :3]
        fixture = load_colour(colour_name) if colour_name else None
        if fixture is not None:
            ref, stride = fixture
            img = img[::stride, ::stride]
            err = int(np.abs(img.astype(np.int32) - ref.astype(np.int32)).max())
            mse = float(((img.astype(np.float64) - ref.astype(np.float64)) ** 2).mean())
            out["colour"] = {"max_abs_err": err, "psnr_db": round(10 * np.log10(255.0 ** 2 / mse), 1) if mse > 0 else "inf", "ok": err <= 2}
    return out


def stage_table(rast, render, abytes: dict, peak_gbs: float, reps: int = 5) -> dict:
    """Per-stage device time of `render()` (L2 evicted before each repetition) and the stage's algorithmic HBM rate."""
    rast.enable_stage_timing(True)
    acc = {}
    for _ in range(reps):
        rast.flush_l2()
        render()
        for k, (us, n) in rast.stage_times_us().items():
            a = acc.setdefault(k, [0.0, 0])
            a[0] += us / reps
            a[1] = n
    rast.enable_stage_timing(False)
    stages = {}
    for k in ("clear", "mesh", "bin", "raster", "resolve"):
        us = acc.get(k, [0.0, 0])[0]
        if us <= 0:
            continue
        stages[k] = {"us": round(us, 2), "launches": acc[k][1]}
        b = abytes.get(k)
        if b:
            stages[k].update({"algorithmic_bytes": int(b), "GBs": round(b / us / 1e3, 1), "frac": round(b / us / 1e3 / peak_gbs, 4)})
    return stages


def time_frames(rast, stream, render, frames: int, flush: bool) -> list:
    import torch
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(frames)]
    for b, e in ev:
        if flush:
            rast.flush_l2()
        b.record(stream)
        render()
        e.record(stream)
    torch.cuda.synchronize()
    return [b.elapsed_time(e) for b, e in ev]


def run_config(name: str, local_rank: int, mode: str, golden: dict, peak_gbs: float) -> dict:
    """One of the other BASELINE configs: parity against its fixture, then frame-by-frame timing with the L2 evicted."""
    import torch
    from glimpsw_b200 import api
    wl = workloads.build(name)
    scene = wl.scene
    g = golden.get(name, {})
    rast = api.Rasterizer(local_rank, enable_binning=(mode == "binned"), fused_frustum_cull=wl.fused_cull)
    st = torch.cuda.Stream()
    rast.set_stream(st.cuda_stream)
    gscene = rast.upload_scene(scene.meshlets, scene.materials if len(scene.materials) else None, scene.textures, scene.lights)
    fb = rast.create_framebuffer(scene.width, scene.height)
    views = [None] if wl.cameras is None else [0, 21]
    out = {"workload": wl.description, "triangles_per_frame": scene.num_triangles, "meshlets": len(scene.meshlets), "views": []}
    for v in views:
        batch = rast.create_batch(gscene, workloads.view_draws(rast, wl, v))
        uni = api.Rasterizer.make_uniforms(**workloads.view_uniforms(wl, v)) if wl.resolve else None
        frame = rast.make_frame(batch, uni)
        want = g if v is None else g.get("views", {}).get(str(v))
        parity = check_view(rast, fb, batch, frame, want if want and "depth_sha256" in want else None, (want or {}).get("colour"), wl.resolve)

        def render():
            rast.submit_frame(fb, frame)
        for _ in range(3):
            render()
        rast.set_mesh_occupancy(4)
        lat = time_frames(rast, st, render, 20, flush=True)
        pipe = time_frames(rast, st, render, 40, flush=False)
        rast.reset_counters()
        render()
        c = rast.counters()
        stats = rast.draw_stats()
        visible = (want or {}).get("meshlets_visible", len(scene.meshlets))
        stages = stage_table(rast, render, workloads.algorithmic_bytes(wl, len(scene.meshlets), visible), peak_gbs)
        ms = float(np.median(lat))
        entry = {"view": v, "parity": parity, "ms_per_frame": round(ms, 4), "ms_per_frame_back_to_back": round(float(np.median(pipe)), 4),
                 "frames_per_s": round(1e3 / ms, 1), "submitted_Mtri_s": round(scene.num_triangles / ms / 1e3, 1),
                 "processed_Mtri_s": round(c["TrianglesProcessed"] / ms / 1e3, 1), "rasterized_Mtri_s": round(c["TrianglesRasterized"] / ms / 1e3, 1),
                 "counters": {k: c[k] for k in ("TrianglesProcessed", "TrianglesRasterized", "TrianglesClipped", "BinQueueFlushes")},
                 "draw_stats": stats, "stages": stages}
        out["views"].append(entry)
        batch.destroy()
    rast.destroy()
    return out


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
    # stdout carries exactly one JSON line: anything libraries print there meanwhile is routed to stderr until it is written
    sys.stdout.flush()
    saved_stdout = os.dup(1)
    os.dup2(2, 1)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    wl = workloads.build("c4_views")
    scene = wl.scene
    golden = load_golden()
    gviews = golden.get("c4_views", {}).get("views", {})
    tris = scene.num_triangles
    num_views = len(wl.cameras)
    mine = sharding.views_for_rank(num_views, rank, world)
    peak_gbs, peak_src = peak_hbm()
    F = max(1, min(args.in_flight, len(mine)))
    binned = args.mode == "binned"

    # ---- F independent render contexts on this GPU (own stream, framebuffer, scene copy, work buffers)
    ctxs = []
    for i in range(F):
        r = api.Rasterizer(local_rank, enable_binning=binned, fused_frustum_cull=True)
        st = torch.cuda.Stream()
        r.set_stream(st.cuda_stream)
        r.set_mesh_occupancy(args.mesh_blocks if F > 1 else 4)
        ctxs.append(SimpleNamespace(rast=r, stream=st, fb=r.create_framebuffer(scene.width, scene.height),
                                    scene=r.upload_scene(scene.meshlets, scene.materials, scene.textures, scene.lights), frames={}))
    # ---- N > 1: deal the views by measured cost instead of v mod N. The views of the batch cost 191..283 us each; with 8 views per
    # GPU the round-robin deal leaves the slowest rank 7 % over the mean (tools/view_costs.py). Every rank times its round-robin
    # share once (frames back to back on one context), the 64 costs are all-gathered, and every rank computes the same
    # longest-first deal with equal counts.
    deal = {"policy": "v mod N"}
    if world > 1 and not args.round_robin:
        c0 = ctxs[0]
        costs = torch.zeros(num_views, dtype=torch.float64, device="cuda")
        for v in mine:
            batch = c0.rast.create_batch(c0.scene, workloads.view_draws(c0.rast, wl, v))
            frame = c0.rast.make_frame(batch, api.Rasterizer.make_uniforms(**workloads.view_uniforms(wl, v)))
            for _ in range(2):
                c0.rast.submit_frame(c0.fb, frame)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(c0.stream)
            for _ in range(4):
                c0.rast.submit_frame(c0.fb, frame)
            e1.record(c0.stream)
            e1.synchronize()
            costs[v] = e0.elapsed_time(e1) / 4
            batch.destroy()
        dist.all_reduce(costs)                                    # every view was timed by exactly one rank
        dist.broadcast(costs, src=0)                              # bit-identical on every rank by construction; cheap insurance
        cost_list = [float(x) for x in costs.cpu()]
        shares = sharding.deal_views_by_cost(cost_list, world)
        mean = sum(cost_list) / world
        deal = {"policy": "longest measured cost first, equal counts (sharding.deal_views_by_cost)",
                "view_cost_us_min_max": [round(min(cost_list) * 1e3, 1), round(max(cost_list) * 1e3, 1)],
                "balance_v_mod_N": round(mean / max(sum(cost_list[v] for v in sharding.views_for_rank(num_views, r, world)) for r in range(world)), 4),
                "balance_this_deal": round(mean / max(sum(cost_list[v] for v in sh) for sh in shares), 4)}
        mine = shares[rank]

    for i, v in enumerate(mine):
        c = ctxs[i % F]
        batch = c.rast.create_batch(c.scene, workloads.view_draws(c.rast, wl, v))
        uni = api.Rasterizer.make_uniforms(**workloads.view_uniforms(wl, v))
        c.frames[i] = (batch, c.rast.make_frame(batch, uni))

    # ---- parity: every view of this rank against the oracle fixture, before anything is timed
    parity = {"views_checked": 0, "visbuffer_exact": 0, "colour_checked": 0, "colour_ok": 0, "worst_colour_err": 0, "mismatches": []}
    for i, v in enumerate(mine):
        c = ctxs[i % F]
        batch, frame = c.frames[i]
        want = gviews.get(str(v))
        res = check_view(c.rast, c.fb, batch, frame, want, (want or {}).get("colour"), True)
        parity["views_checked"] += 1
        if res["visbuffer"] == "exact":
            parity["visbuffer_exact"] += 1
        else:
            parity["mismatches"].append({"view": v, "visbuffer": res["visbuffer"]})
        if isinstance(res["colour"], dict):
            parity["colour_checked"] += 1
            parity["colour_ok"] += 1 if res["colour"]["ok"] else 0
            parity["worst_colour_err"] = max(parity["worst_colour_err"], res["colour"]["max_abs_err"])
    if world > 1:
        t = torch.tensor([parity["views_checked"], parity["visbuffer_exact"], parity["colour_checked"], parity["colour_ok"]], device="cuda")
        dist.all_reduce(t)
        worst = torch.tensor([parity["worst_colour_err"]], device="cuda")
        dist.all_reduce(worst, op=dist.ReduceOp.MAX)
        parity.update(dict(zip(["views_checked", "visbuffer_exact", "colour_checked", "colour_ok"], [int(x) for x in t.tolist()])))
        parity["worst_colour_err"] = int(worst.item())
    parity["against"] = "tests/golden/bench_configs.json (CPU oracle: SHA-256 of depth + surface ids + counters, exact; colour fixtures <= 2/255)"
    if parity["visbuffer_exact"] != parity["views_checked"] and rank == 0:
        print(f"bench.py: PARITY MISMATCH {parity}", file=sys.stderr)

    # ---- composites: every finished view goes to rank 0 (N > 1: over NVLink), on a side stream, SLOTS in flight
    comm, coll = torch.cuda.Stream(), torch.cuda.Stream()
    gather_kind, peers = ("local", None) if world == 1 else (args.gather, None)
    if world > 1 and gather_kind == "p2p":
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
    ring = [torch.empty((scene.height, scene.width), dtype=torch.int32, device="cuda") for _ in range(SLOTS)]
    gathered = [[torch.empty_like(ring[0]) for _ in range(world)] for _ in range(SLOTS)] if (gather_kind == "nccl" and rank == 0) else [None] * SLOTS
    slot_done = [torch.cuda.Event() for _ in range(SLOTS)]
    state = {"k": 0}

    def render_view(i):
        """One view: Clear -> DrawMeshlets x 122 -> Resolve (one C call), then GetPixels to where the composites are collected."""
        c = ctxs[i % F]
        c.rast.submit_frame(c.fb, c.frames[i][1])
        slot = state["k"] % SLOTS
        state["k"] += 1
        if gather_kind == "none":
            return
        if peers is not None:
            peers.send(c.fb, slot, comm)                        # GetPixels straight into rank 0's memory (+ slot flow control)
            peers.collect(c.rast, slot, coll)                   # rank 0 waits for its peers on a third stream
            slot_done[slot].record(coll if rank == 0 else comm)
        elif gather_kind == "nccl":
            with torch.cuda.stream(comm):
                comm.wait_event(slot_done[slot])
                c.fb.get_pixels_device(0, ring[slot].data_ptr(), cuda_stream=comm.cuda_stream)
                dist.gather(ring[slot], gathered[slot], dst=0)
                slot_done[slot].record(comm)
        else:
            c.fb.get_pixels_device(0, ring[slot].data_ptr(), cuda_stream=comm.cuda_stream)
            slot_done[slot].record(comm)

    def step():
        for i in range(len(mine)):
            render_view(i)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        step()
    barrier()

    # ---- timed region: exactly K steps (K x 64 views), F contexts in flight per GPU
    sampler = ClockSampler(local_rank)
    if rank == 0:               # one nvidia-smi poller per job: NVML queries from 8 pollers perturb the launches they watch
        sampler.start()
    launches0 = sum(c.rast.launch_count() for c in ctxs)
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    nvl0 = nvlink_kib(local_rank) if (rank == 0 and world > 1) else None     # rank 0's NVLink byte counters around the timed region
    barrier()
    t_wall0 = time.perf_counter()
    t0.record(ctxs[0].stream)
    for c in ctxs[1:]:
        c.stream.wait_event(t0)
    t_sub0 = time.perf_counter()
    step()                                                        # host cost of enqueueing one batch (the queues are empty)
    t_submit = (time.perf_counter() - t_sub0) / max(1, len(mine))
    for _ in range(1, args.steps):
        step()
    for c in ctxs[1:]:
        ev = torch.cuda.Event()
        ev.record(c.stream)
        ctxs[0].stream.wait_event(ev)
    t_rendered = torch.cuda.Event(enable_timing=True)             # this rank's own views are rendered (the exchange may still be draining)
    t_rendered.record(ctxs[0].stream)
    if gather_kind != "none":                                     # the last composites are part of the job
        for ev_done in slot_done:
            ctxs[0].stream.wait_event(ev_done)
    t1.record(ctxs[0].stream)
    barrier()
    t_wall = time.perf_counter() - t_wall0
    launches = sum(c.rast.launch_count() for c in ctxs) - launches0
    clocks = sampler.stop()
    nvl1 = nvlink_kib(local_rank) if nvl0 else None
    total_ms = t0.elapsed_time(t1)
    per_rank = None
    if world > 1:
        mine_ms = torch.tensor([t0.elapsed_time(t_rendered), total_ms], dtype=torch.float64, device="cuda")
        all_ms = [torch.zeros_like(mine_ms) for _ in range(world)]
        dist.all_gather(all_ms, mine_ms)
        per_rank = {"rendered_ms_per_step": [round(float(x[0]) / args.steps, 4) for x in all_ms],
                    "done_ms_per_step": [round(float(x[1]) / args.steps, 4) for x in all_ms]}
        t = torch.tensor([total_ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms = float(t.item())
        lt = torch.tensor([launches], dtype=torch.int64, device="cuda")
        dist.all_reduce(lt)
        launches = int(lt.item())
    ms_per_step = total_ms / args.steps
    value = tris * num_views / (ms_per_step * 1e-3) / 1e6
    nvlink = {"unavailable": "nvidia-smi nvlink -gt d reports no data counters (N/A) on this box"} if (rank == 0 and world > 1) else None
    if nvl0 and nvl1:
        rx, tx = (nvl1[0] - nvl0[0]) * 1024, (nvl1[1] - nvl0[1]) * 1024
        expect = (num_views - len(mine)) * scene.width * scene.height * 4 * args.steps if gather_kind in ("p2p", "nccl") else 0
        nvlink = {"rank0_rx_bytes": int(rx), "rank0_tx_bytes": int(tx), "rank0_rx_GBs": round(rx / (total_ms * 1e-3) / 1e9, 1),
                  "expected_composite_bytes": int(expect), "links": nvl0[2], "link_peak_GBs_per_direction": 900.0,
                  "source": "nvidia-smi nvlink -gt d on rank 0's GPU, sampled before and after the timed region"}

    # ---- the gathered composites are the producers' images (N > 1): checksum of the last round on both sides
    gather_check = None
    if world > 1 and peers is not None:
        last_slot = (state["k"] - 1) % SLOTS
        c = ctxs[(len(mine) - 1) % F]
        local = torch.empty((scene.height, scene.width), dtype=torch.int32, device="cuda")
        c.fb.get_pixels_device(0, local.data_ptr())
        torch.cuda.synchronize()
        mine_sum = torch.tensor([int(local.to(torch.int64).sum().item())], dtype=torch.int64, device="cuda")
        sums = [torch.zeros_like(mine_sum) for _ in range(world)]
        dist.all_gather(sums, mine_sum)
        if rank == 0:
            got = [int(peers.buf[last_slot][r].to(torch.int64).sum().item()) for r in range(world)]
            gather_check = {"ranks": world, "matching": sum(1 for r in range(world) if got[r] == int(sums[r].item())),
                            "what": "sum over the pixels of each rank's last composite: producer's framebuffer vs rank 0's gathered copy"}

    # ---- counters of one whole batch (all ranks)
    for c in ctxs:
        c.rast.reset_counters()
    save_gather, gather_kind = gather_kind, "none"
    step()
    gather_kind = save_gather
    csum = np.zeros(4, dtype=np.int64)
    for c in ctxs:
        cc = c.rast.counters()
        csum += np.array([cc["TrianglesProcessed"], cc["TrianglesRasterized"], cc["TrianglesClipped"], cc["BinQueueFlushes"]], dtype=np.int64)
    if world > 1:
        t = torch.tensor(csum, device="cuda")
        dist.all_reduce(t)
        csum = t.cpu().numpy()

    # ---- latency mode (rank 0's first views, one context, strictly serial, L2 evicted before every frame) + stage table
    c0 = ctxs[0]
    c0.rast.set_mesh_occupancy(4)                                 # a lone frame gets the whole register file
    lat_views = [i for i in range(len(mine)) if i % F == 0][:4]
    lat_ms, stages_per_view, stats_per_view = [], [], []
    for i in lat_views:
        def render(i=i):
            c0.rast.submit_frame(c0.fb, c0.frames[i][1])
        render()
        lat_ms += time_frames(c0.rast, c0.stream, render, 8, flush=True)
        want = gviews.get(str(mine[i]), {})
        visible = want.get("meshlets_visible", len(scene.meshlets))
        stages_per_view.append(stage_table(c0.rast, render, workloads.algorithmic_bytes(wl, len(scene.meshlets), visible), peak_gbs, reps=3))
        stats_per_view.append(c0.rast.draw_stats())
    stages = {}
    for k in ("clear", "mesh", "bin", "raster", "resolve"):
        rows = [s[k] for s in stages_per_view if k in s]
        if not rows:
            continue
        us = float(np.mean([r["us"] for r in rows]))
        stages[k] = {"us": round(us, 2)}
        if "algorithmic_bytes" in rows[0]:
            b = float(np.mean([r["algorithmic_bytes"] for r in rows]))
            stages[k].update({"algorithmic_bytes": int(b), "GBs": round(b / us / 1e3, 1), "frac": round(b / us / 1e3 / peak_gbs, 4)})
    draw_stats = {k: int(np.mean([s[k] for s in stats_per_view])) for k in stats_per_view[0]} if stats_per_view else {}
    traffic = {}
    try:   # DRAM bytes per launch from the committed ncu --set full capture of this workload
        traffic = json.load(open(os.path.join(ROOT, "profiles", "r02_traffic.json")))
    except Exception:
        pass
    dom = max((k for k in stages if "GBs" in stages[k]), key=lambda k: stages[k]["us"])
    kernel = {"mesh": "k_mesh_setup", "raster": "k_tile_raster" if binned else "k_raster_direct", "resolve": "k_resolve"}[dom]
    roofline = {"bound": "hbm", "kernel": kernel, "achieved": stages[dom]["GBs"], "peak": peak_gbs, "unit": "GB/s", "frac": stages[dom]["frac"],
                "traffic": traffic.get(kernel), "peak_source": peak_src, "algorithmic_bytes": stages[dom]["algorithmic_bytes"],
                "avg_launch_us": stages[dom]["us"],
                "note": "dominant stage of the frame by device time; time = CUDA events around the stage's launches on the launching stream, "
                        f"one frame at a time with the L2 evicted before each, mean over views {[mine[i] for i in lat_views]}; algorithmic bytes per SURVEY §8(d) "
                        "(mesh: 16 B per tested meshlet + 1,216 B per meshlet that survives the frustum test); the kernels are instruction-issue / "
                        "latency bound, not HBM bound (profiles/r02_summary.md)"}
    if roofline["traffic"] is not None:
        roofline["traffic_source"] = "profiles/r02_traffic.json (ncu dram__bytes_read.sum + dram__bytes_write.sum, per launch)"
    c0.rast.set_mesh_occupancy(args.mesh_blocks if F > 1 else 4)

    # ---- e2e: the same batch through the C ABI with HOST buffers. Every step: the scene goes up from pinned host memory (this rank's
    # 1/N of it, all-gathered over NVLink when N > 1) into a loader context's scene (own stream), is copied on the device to the render contexts' scenes,
    # and every view's resolved image comes back to pinned host memory. Two scene buffers per context (A/B), so the next step's
    # upload overlaps this step's rendering; all F contexts render their share of every step.
    chunk = (len(scene.meshlets) + world - 1) // world
    lo, hi = min(rank * chunk, len(scene.meshlets)), min((rank + 1) * chunk, len(scene.meshlets))
    host_meshlets = ctxs[0].rast.alloc_pinned((hi - lo,), scene.meshlets.dtype)
    host_meshlets[...] = scene.meshlets[lo:hi]
    host_images = [ctxs[0].rast.alloc_pinned((scene.height, scene.width), np.uint32) for _ in range(min(len(mine), 8))]

    class _Raw:        # a scene's meshlet array as a torch tensor (zero copy)
        def __init__(self, ptr, n):
            self.__cuda_array_interface__ = {"shape": (n,), "typestr": "|u1", "data": (ptr, False), "version": 3}
    nbytes = len(scene.meshlets) * 1728
    # the loader: a context of its own (own stream) that only receives the scene from the host, so uploads never sit between frames
    loader = SimpleNamespace(rast=api.Rasterizer(local_rank), stream=torch.cuda.Stream())
    loader.rast.set_stream(loader.stream.cuda_stream)
    loader.scenes2 = [loader.rast.upload_scene(scene.meshlets, scene.materials, scene.textures, scene.lights) for _ in range(2)]
    loader.tensors2 = [torch.as_tensor(_Raw(sc.meshlets_device_ptr(), nbytes), device="cuda") for sc in loader.scenes2]
    for c in ctxs:
        c.scenes2 = [c.scene, c.rast.upload_scene(scene.meshlets, scene.materials, scene.textures, scene.lights)]
        c.tensors2 = [torch.as_tensor(_Raw(sc.meshlets_device_ptr(), nbytes), device="cuda") for sc in c.scenes2]
        c.frames2 = [dict(c.frames), {}]
        c.copied = [torch.cuda.Event(), torch.cuda.Event()]      # this context's copy of scene buffer b has been made
    for i, v in enumerate(mine):
        c = ctxs[i % F]
        batch = c.rast.create_batch(c.scenes2[1], workloads.view_draws(c.rast, wl, v))
        c.frames2[1][i] = (batch, c.rast.make_frame(batch, api.Rasterizer.make_uniforms(**workloads.view_uniforms(wl, v))))
    uploaded = [torch.cuda.Event(), torch.cuda.Event()]
    even = world > 1 and len(scene.meshlets) % world == 0
    e2e_state = {"steps": 0}
    skip = set()                                                  # --e2e-diag: the legs left out of a diagnostic pass

    def step_e2e(k):
        b = k % 2
        if e2e_state["steps"] >= 2:
            uploaded[b].synchronize()                             # pace the host: at most two steps ahead of the GPU
            for c in ctxs:
                loader.stream.wait_event(c.copied[b])             # nobody still reads the loader's buffer b (device copies of step k - 2)
        if "h2d" not in skip:
            loader.scenes2[b].update_meshlets(host_meshlets, lo)  # H2D on the loader's stream, beside the rendering of step k - 1
        if world > 1 and "gather" not in skip:
            with torch.cuda.stream(loader.stream):
                tsr = loader.tensors2[b]
                if even:
                    dist.all_gather_into_tensor(tsr, tsr[lo * 1728:hi * 1728])
                else:
                    parts = [tsr[min(r * chunk, len(scene.meshlets)) * 1728:min((r + 1) * chunk, len(scene.meshlets)) * 1728] for r in range(world)]
                    dist.all_gather(parts, tsr[lo * 1728:hi * 1728])
        uploaded[b].record(loader.stream)
        for c in ctxs:
            with torch.cuda.stream(c.stream):
                c.stream.wait_event(uploaded[b])
                if "d2d" not in skip:
                    c.tensors2[b].copy_(loader.tensors2[b], non_blocking=True)   # device-to-device, after this context's own frames of step k - 2
                c.copied[b].record(c.stream)
            if "d2d" not in skip:
                c.scenes2[b].touch()
        for i in range(len(mine) if "render" not in skip else 0):
            c = ctxs[i % F]
            batch, frame = c.frames2[b][i]
            frame.PixelsHost = host_images[i % len(host_images)].ctypes.data if "d2h" not in skip else None
            c.rast.submit_frame(c.fb, frame)                      # clear + draw + resolve + GetPixels (D2H, async, on the context's copy stream)
            frame.PixelsHost = None
        e2e_state["steps"] += 1

    for k in range(4):
        step_e2e(k)
    barrier()
    t0e = time.perf_counter()
    for k in range(args.steps):
        step_e2e(k)
    barrier()
    e2e_s = time.perf_counter() - t0e
    if world > 1:
        t = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    e2e_value = tris * num_views * args.steps / e2e_s / 1e6
    if args.e2e_diag:                                             # which leg of the e2e step costs what: the same loop with one leg left out
        for legs in ((), ("h2d",), ("gather",), ("d2d",), ("d2h",), ("render",), ("h2d", "gather", "d2d"), ("h2d", "gather", "d2d", "d2h")):
            skip.clear(); skip.update(legs)
            for k in range(4):
                step_e2e(k)
            barrier()
            t0d = time.perf_counter()
            for k in range(args.steps):
                step_e2e(k)
            barrier()
            if rank == 0:
                print(f"[e2e-diag] without {'+'.join(legs) or 'nothing'}: {(time.perf_counter() - t0d) / args.steps * 1e3:.3f} ms/step", file=sys.stderr, flush=True)
        skip.clear()
    checksum = int(np.bitwise_xor.reduce(host_images[0].reshape(-1)))

    # ---- N = 1 only: the other BASELINE configs, and the CPU baseline on the host cores
    configs, cpu = None, None
    if rank == 0 and world == 1:
        for c in ctxs:
            c.rast.destroy()
        ctxs = []
        torch.cuda.empty_cache()
        if not args.no_configs:
            configs = {}
            for name in ("c1_knot", "c1_sponza", "c2_grid", "c3_knot", "c5_views"):
                try:
                    configs[name] = run_config(name, local_rank, args.mode, golden, peak_gbs)
                except Exception as exc:
                    configs[name] = {"error": repr(exc)}
        if not args.no_cpu_baseline:
            os.sched_setaffinity(0, full_affinity)          # the CPU baseline gets every host core, not just the GPU's NUMA node
            cpu = cpu_views(wl, REF_SAMPLE_VIEWS, budget_s=args.cpu_budget)

    if rank == 0:
        line = {
            "metric": METRIC, "value": round(value, 2), "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": round(ms_per_step, 5), "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f32+i32 (28.4 fixed-point coverage, fp32 depth/shading)", "data": "synthetic",
            "config": {"workload": wl.description, "step": f"the whole batch of {num_views} views",
                       "triangles_per_view": tris, "triangles_per_step": tris * num_views, "meshlets": len(scene.meshlets),
                       "draws_per_view": len(scene.nodes), "mode": args.mode, "frames_in_flight": F,
                       "mesh_kernel_blocks_per_sm": args.mesh_blocks if F > 1 else 4,
                       "parallelism": ("view-parallel: views dealt v mod " + str(world) if deal["policy"] == "v mod N" else
                                       f"view-parallel: the {num_views} views dealt to the {world} GPUs by measured cost, longest first, {len(mine)} each") + {
                           "p2p": ", composites stored by each rank's de-tile kernel straight into rank 0's memory over NVLink (peer memory + device-side flags, 4 slots in flight, tail included)",
                           "nccl": ", composites gathered to rank 0 with NCCL on a side stream (tail included)", "none": ", no gather (diagnostic)",
                           "local": ", composites de-tiled into a device ring buffer on a side stream"}[gather_kind],
                       "l2": f"inputs larger than L2: {scene.meshlets.nbytes / 1e6:.0f} MB of meshlets per scene copy, one copy per context; the latency mode "
                             "additionally evicts L2 (256 MB write + 256 MB read) before every frame",
                       "timing": "one CUDA-event pair around the K steps on the launching streams (all contexts and the composite streams joined), max over ranks"},
            "views_per_s": round(num_views / (ms_per_step * 1e-3), 1), "frames_per_s": round(num_views / (ms_per_step * 1e-3), 1),
            "ms_per_view": round(ms_per_step / num_views * world, 5),
            "processed_Mtri_s": round(int(csum[0]) / (ms_per_step * 1e-3) / 1e6, 2),
            "rasterized_Mtri_s": round(int(csum[1]) / (ms_per_step * 1e-3) / 1e6, 2),
            "latency_ms_per_view": round(float(np.median(lat_ms)), 5),
            "latency_ms_min_max": [round(float(np.min(lat_ms)), 5), round(float(np.max(lat_ms)), 5)],
            "wall_ms_per_step": round(t_wall / args.steps * 1e3, 4),
            "host_submit_us_per_view": round(t_submit * 1e6, 2),
            "clocks": clocks, "gpu_launches": int(launches), "parity": parity,
            "e2e": {"value": round(e2e_value, 2), "unit": UNIT, "h2d_bytes_per_step": int(scene.meshlets.nbytes),
                    "d2h_bytes_per_step": int(num_views * scene.width * scene.height * 4), "ms_per_step": round(e2e_s / args.steps * 1e3, 4),
                    "views_per_s": round(num_views * args.steps / e2e_s, 1),
                    "note": f"every step: the {scene.meshlets.nbytes / 1e6:.0f} MB of meshlets H2D from pinned host memory (each rank 1/{world} of them"
                            + (", all-gathered over NVLink with NCCL" if world > 1 else "") + f"), {num_views} x (clear + draw + resolve + GetPixels D2H to pinned host memory, "
                            f"{scene.width * scene.height * 4 / 1e6:.1f} MB per view) spread over the {F} render contexts of each GPU, whose scene copies are refreshed device-to-device "
                            "from the uploaded one; two scene buffers per context so the next step's upload overlaps; wall clock; bytes are whole-job totals"},
            "roofline": roofline, "stages": stages,
            "counters_per_step": {"TrianglesProcessed": int(csum[0]), "TrianglesRasterized": int(csum[1]), "TrianglesClipped": int(csum[2]), "BinQueueFlushes": int(csum[3])},
            "draw_stats": draw_stats, "image_xor": checksum,
        }
        wi = traffic.get("warp_instructions")
        if wi and clocks.get("sm_mhz"):
            # the kernels are issue-bound, so the roof that explains `value` is the GPU's instruction issue rate: warp instructions
            # one view executes (ncu smsp__inst_executed.sum of the profiled view, profiles/r02_traffic.json) per second of the timed
            # region, against 4 schedulers x SMs x the SM clock sampled during that region
            import torch
            per_view = float(sum(wi.values()))
            peak = torch.cuda.get_device_properties(local_rank).multi_processor_count * 4 * clocks["sm_mhz"] * 1e6
            achieved = per_view / (ms_per_step / num_views * world * 1e-3)
            line["issue_slots"] = {"warp_instructions_per_view": int(per_view), "achieved_per_s": round(achieved, 1), "peak_per_s": round(peak, 1),
                                   "frac": round(achieved / peak, 4), "per_kernel": wi,
                                   "note": "per GPU; instruction counts from the ncu capture of view 0 (profiles/r02_traffic.json), time from this run"}
        if gather_check is not None:
            line["gather_check"] = gather_check
        if nvlink is not None:
            line["nvlink"] = nvlink
        if world > 1:
            line["deal"] = deal
            line["per_rank"] = per_rank
        if configs is not None:
            line["configs"] = configs
        if cpu:
            line["cpu_baseline"] = cpu
        sys.stdout.flush()
        os.dup2(saved_stdout, 1)
        print(json.dumps(line), flush=True)
    for c in ctxs:
        c.rast.destroy()
    if world > 1:
        dist.destroy_process_group()


def cpu_render_view(base, orc, wl, v, fb):
    """One view on the CPU restatement: clear + CullMeshlets + DrawMeshlets per node + Resolve (the same frame loop)."""
    scene = wl.scene
    scene.camera = wl.cameras[v]
    base.clear(fb, 0xFF000000, 0.0)
    proj, vm = scene.view_proj()
    for node in scene.nodes:
        planes = orc.frustum_planes(proj, vm, node.model)
        bitmap, _ = orc.cull_meshlets(scene.meshlets[node.meshlet_offset:node.meshlet_offset + node.meshlet_count], planes)
        base.draw_meshlets(fb, scene.meshlets, node.meshlet_offset, node.meshlet_count, scene.object_to_clip(node), cull_bitmap=bitmap,
                           materials=scene.materials)
    base.resolve(fb, scene.meshlets, scene.materials, scene.textures, scene.lights, **workloads.view_uniforms(wl, v))


def cpu_views(wl, views, budget_s: float, threads: int = 0):
    """Times the CPU restatement of the reference on a bounded sample of the batch (rank 0, N = 1)."""
    from oracle import orc
    orc.build()
    out = None
    hw = os.cpu_count() or 1
    for label, nthreads in (("all", threads), ("reference_default", max(hw // 2, 1))):     # Rasterizer.cpp:133-138: hardware_concurrency / 2
        base = orc.Baseline(nthreads)
        fb = orc.Framebuffer(wl.scene.width, wl.scene.height)
        cpu_render_view(base, orc, wl, views[0], fb)                # warm-up
        times, t_start = [], time.perf_counter()
        for v in views:
            t0 = time.perf_counter()
            cpu_render_view(base, orc, wl, v, fb)
            times.append(time.perf_counter() - t0)
            if time.perf_counter() - t_start > budget_s / 2 and len(times) >= 3:
                break
        med = float(np.median(times))
        if label == "all":
            out = {"value": round(wl.scene.num_triangles / med / 1e6, 2), "unit": UNIT, "cores": base.threads, "kind": "port",
                   "isa": "avx512 (16-lane vertex transform, 16-triangle packet classification + early setup, 4x4-fragment raster loop, resolve; "
                          "per-triangle edge setup and binning scalar)" if base.avx512 else "scalar (no AVX-512 on this host)",
                   "sample": f"views {views[:len(times)]} of the batch, one frame each after 1 warm-up (median {med * 1e3:.1f} ms/view)",
                   "ms_per_view": round(med * 1e3, 3), "cpu": cpu_model()}
        else:
            out["reference_default_threads"] = {"cores": base.threads, "value": round(wl.scene.num_triangles / med / 1e6, 2), "ms_per_view": round(med * 1e3, 3),
                                                "note": "Rasterizer::SetThreadCount(0) = hardware_concurrency / 2 (Rasterizer.cpp:133-138)"}
        base.close()
    return out


def run_reference(args):
    """--impl reference: the CPU restatement of GLimpSW's path on the host cores (rank 0 only)."""
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if rank != 0:
        return
    wl = workloads.build("c4_views")
    from oracle import orc
    orc.build()
    base = orc.Baseline(0)
    fb = orc.Framebuffer(wl.scene.width, wl.scene.height)
    views = REF_SAMPLE_VIEWS[:max(1, args.ref_views)]

    def step():
        for v in views:
            cpu_render_view(base, orc, wl, v, fb)

    for _ in range(min(args.warmup, 1)):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = time.perf_counter() - t0
    ms_view = dt / (args.steps * len(views)) * 1e3
    value = wl.scene.num_triangles / (ms_view * 1e-3) / 1e6
    num_views = len(wl.cameras)
    cpu = {"value": round(value, 2), "unit": UNIT, "cores": base.threads, "kind": "port",
           "isa": "avx512" if base.avx512 else "scalar",
           "sample": f"each step renders views {views} of the {num_views}-view batch (clear + cull + draw + resolve), all host threads; "
                     f"ms_per_step is scaled to the whole batch", "cpu": cpu_model()}
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": round(value, 2), "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": min(args.warmup, 1), "ms_per_step": round(ms_view * num_views, 3), "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f32+i32", "data": "synthetic",
        "config": {"workload": wl.description, "step": f"the whole batch of {num_views} views", "triangles_per_view": wl.scene.num_triangles,
                   "triangles_per_step": wl.scene.num_triangles * num_views, "meshlets": len(wl.scene.meshlets), "draws_per_view": len(wl.scene.nodes),
                   "note": "CPU restatement of GLimpSW's binned AVX-512 path (oracle/baseline_mt.cpp); bit-identical with the reference's own sources built by oracle/ref_build.py, which only compile on a slow lane-array stand-in for Clang's vector types and are therefore not the timed arm"},
        "cpu_baseline": cpu,
        "e2e": {"value": round(value, 2), "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }), flush=True)
    base.close()


def nvlink_kib(index):
    """(rx KiB, tx KiB, links) summed over the NVLinks of one GPU (`nvidia-smi nvlink -gt d`), or None when the counters are not exposed."""
    import re
    try:
        import torch
        uuid = str(torch.cuda.get_device_properties(index).uuid)
        out = subprocess.run(["nvidia-smi", "nvlink", "-gt", "d", "-i", uuid if uuid.startswith("GPU-") else "GPU-" + uuid],
                             capture_output=True, text=True, timeout=20).stdout
        rx = [int(x) for x in re.findall(r"Data Rx:\s*(\d+)\s*KiB", out)]
        tx = [int(x) for x in re.findall(r"Data Tx:\s*(\d+)\s*KiB", out)]
        return (sum(rx), sum(tx), len(rx)) if rx and tx else None
    except Exception:
        return None


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--mode", default="binned", choices=["binned", "direct"])
    ap.add_argument("--in-flight", type=int, default=4, help="independent render contexts (views in flight) per GPU")
    ap.add_argument("--mesh-blocks", type=int, default=2, help="mesh-kernel blocks per SM with several contexts in flight (swrb_device_set_mesh_occupancy)")
    ap.add_argument("--gather", default="p2p", choices=["p2p", "nccl", "none"], help="N>1: how composites reach rank 0 (none = diagnostic: no exchange)")
    ap.add_argument("--cpu-budget", type=float, default=20.0, help="seconds of CPU-baseline frames at N=1")
    ap.add_argument("--ref-views", type=int, default=4, help="--impl reference: views of the batch rendered per step")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--e2e-diag", action="store_true", help="after the e2e leg, time it again with one leg left out at a time (stderr)")
    ap.add_argument("--round-robin", action="store_true", help="N>1: deal the views v mod N instead of by measured cost")
    ap.add_argument("--no-configs", action="store_true", help="skip the per-config block (C1/C2/C3/C5) at N=1")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
