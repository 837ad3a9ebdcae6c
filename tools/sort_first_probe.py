"""Sort-first latency probe: ONE view of the C4 scene split into horizontal bands, one per rank (scissor
rows); every rank culls against its band, draws, resolves and de-tiles its band into its rows of the image; the bands are
all-gathered (no depth compare). Prints the per-view latency and its gather share, then the same view with every rank's
de-tile kernel storing its band straight into rank 0's image over NVLink peer memory (swrb_fb_send_pixels).

    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 --master-port 29530 tools/sort_first_probe.py [view]
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402
from glimpsw_b200 import api, sharding, workloads  # noqa: E402

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
view = int(sys.argv[1]) if len(sys.argv) > 1 else 0
wl = workloads.build("c4_views")
scene = wl.scene
stream = torch.cuda.Stream()
torch.cuda.set_stream(stream)
rast = api.Rasterizer(local, fused_frustum_cull=True)
rast.set_stream(stream.cuda_stream)
gscene = rast.upload_scene(scene.meshlets, scene.materials, scene.textures, scene.lights)
fb = rast.create_framebuffer(scene.width, scene.height)
y0, y1 = sharding.band_rows(scene.height, rank, world)
fb.set_scissor_rows(y0, y1)
batch = rast.create_batch(gscene, workloads.view_draws(rast, wl, view))
uni = workloads.view_uniforms(wl, view)
image = torch.zeros((scene.height, scene.width), dtype=torch.int32, device="cuda")
ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
times, processed = [], 0
for it in range(13):
    rast.flush_l2()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    rast.reset_counters()
    ev[0].record(stream)
    fb.clear(0xFF000000, 0.0)
    rast.draw_prepared(fb, batch)
    ev[1].record(stream)
    rast.resolve(fb, gscene, **uni)
    fb.get_pixels_device(0, image.data_ptr())
    ev[2].record(stream)
    sharding.gather_bands(image, rank, world)
    ev[3].record(stream)
    torch.cuda.synchronize()
    processed = rast.counters()["TrianglesProcessed"]
    if it >= 3:
        times.append([ev[k].elapsed_time(ev[k + 1]) * 1e3 for k in range(3)])
t = torch.tensor(np.median(np.array(times), axis=0), device="cuda")
p = torch.tensor([processed], device="cuda")
if world > 1:
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dist.all_reduce(p, op=dist.ReduceOp.MAX)
if rank == 0:
    draw, res, gat = [float(x) for x in t.cpu()]
    print(f"sort-first N={world} view {view}: draw {draw:.1f} us  resolve+detile {res:.1f} us  band gather {gat:.1f} us  total {draw + res + gat:.1f} us "
          f"(max over ranks of the per-rank medians; most triangles processed by one rank {int(p.item())}; "
          f"image checksum {int(image.cpu().numpy().view(np.uint32).astype(np.uint64).sum())})")

# ---- the same with the band exchange folded into the de-tile kernel: all ranks store into ONE image in rank 0's memory
if world > 1:
    peers = sharding.PeerComposites(scene.height, scene.width, rank, world, slots=2, shared_image=True)
    coll = torch.cuda.Stream()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    times = []
    for it in range(13):
        rast.flush_l2()
        dist.barrier()
        torch.cuda.synchronize()
        e0.record(stream)
        coll.wait_event(e0)
        fb.clear(0xFF000000, 0.0)
        rast.draw_prepared(fb, batch)
        rast.resolve(fb, gscene, **uni)
        slot = it % 2
        peers.send(fb, slot, stream)
        peers.collect(rast, slot, coll)
        e1.record(coll if rank == 0 else stream)
        torch.cuda.synchronize()
        if it >= 3:
            times.append(e0.elapsed_time(e1) * 1e3)
    t = torch.tensor([float(np.median(times))], device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        got = peers.buf[(13 - 1) % 2, 0]
        print(f"sort-first N={world} view {view}, bands stored into rank 0's image by the de-tile kernels: total {float(t.item()):.1f} us "
              f"(rank 0: clear -> all bands arrived; image checksum {int(got.cpu().numpy().view(np.uint32).astype(np.uint64).sum())})")
    dist.barrier()
rast.destroy()
if world > 1:
    dist.destroy_process_group()
