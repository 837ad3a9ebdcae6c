"""Sort-last latency probe: ONE view of the C4 scene split over the ranks by meshlets, key buffers composited
with an NCCL all-reduce(max), resolved on every rank. Prints the per-view latency and its composite share.

    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 --master-port 29530 tools/sort_last_probe.py [view]
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
draws = []
for d in workloads.view_draws(rast, wl, view):                        # this rank's share of every DrawMeshlets call
    first, count = sharding.meshlets_for_rank(d["count"], rank, world)
    if count:
        draws.append(dict(d, offset=d["offset"] + first, count=count))
batch = rast.create_batch(gscene, draws)
uni = workloads.view_uniforms(wl, view)
ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
times = []
for it in range(13):
    rast.flush_l2()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    ev[0].record(stream)
    fb.clear(0xFF000000, 0.0)
    rast.draw_prepared(fb, batch)
    ev[1].record(stream)
    if world > 1:
        sharding.composite_framebuffer(fb, stream=stream)
    ev[2].record(stream)
    rast.resolve(fb, gscene, **uni)
    ev[3].record(stream)
    torch.cuda.synchronize()
    if it >= 3:
        times.append([ev[k].elapsed_time(ev[k + 1]) * 1e3 for k in range(3)])
t = torch.tensor(np.median(np.array(times), axis=0), device="cuda")
if world > 1:
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
if rank == 0:
    draw, comp, res = [float(x) for x in t.cpu()]
    print(f"sort-last N={world} view {view}: draw {draw:.1f} us  composite {comp:.1f} us  resolve {res:.1f} us  total {draw + comp + res:.1f} us "
          f"(max over ranks of the per-rank medians; image checksum {int(fb.get_pixels(0).astype(np.uint64).sum())})")
rast.destroy()
if world > 1:
    dist.destroy_process_group()
