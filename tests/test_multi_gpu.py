"""Multi-GPU (world size 2, NCCL + NVLink peer memory) test of the view-sharded path. Needs 2 GPUs; the test
spawns its own ranks. Run on a multi-GPU box: pytest tests/test_multi_gpu.py -m gpu"""
import os
import socket
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    from glimpsw_b200 import api, scenes, sharding

    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", device_id=torch.device("cuda", rank))
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    comm = torch.cuda.Stream()
    scene = scenes.torus_knot_scene(60, 24, 640, 360, tex_size=64)
    scene.camera.position = scene.camera.position + np.array([0.2 * rank, 0.0, 0.1 * rank])
    node = scene.nodes[0]
    rast = api.Rasterizer(rank)
    rast.set_stream(stream.cuda_stream)
    gscene = rast.upload_scene(scene.meshlets, scene.materials, scene.textures, scene.lights)
    fb = rast.create_framebuffer(scene.width, scene.height)
    uni = scenes.resolve_uniforms(scene, node)
    peers = sharding.PeerComposites(scene.height, scene.width, rank, world)
    sums = []
    got = []
    for k in range(7):                         # more rounds than slots: exercises the ack / slot-reuse path
        slot = k % peers.slots
        fb.clear(0xFF000000 + k, 0.0)
        rast.draw_meshlets(fb, gscene, node.meshlet_offset, node.meshlet_count, scene.object_to_clip(node))
        rast.resolve(fb, gscene, **uni)
        peers.send(fb, slot, stream)               # rank 1: ONE kernel = wait for the slot's ack, de-tile over NVLink, raise ready
        # rank 0: wait for rank 1's ready flag, copy the views out ON THE COLLECT STREAM, only then ack the slot — rank 1
        # runs ahead freely (no barrier in this loop), so a slot released too early would show up as a wrong checksum
        kept = []
        def consume(views):
            with torch.cuda.stream(comm):
                kept.append(views.clone())
        peers.collect(rast, slot, comm, consume=consume if rank == 0 else None)
        local = fb.get_pixels(0)
        sums.append(int(local.astype(np.uint64).sum()))
        if rank == 0:
            comm.synchronize()
            got.append([int(kept[0][r].cpu().numpy().view(np.uint32).astype(np.uint64).sum()) for r in range(world)])
    all_sums = [None] * world
    dist.all_gather_object(all_sums, sums)
    torch.cuda.synchronize()
    dist.barrier()
    if rank == 0:
        q.put((got, all_sums))
    rast.destroy()
    dist.destroy_process_group()


def test_peer_memory_composite_gather_world2():
    import torch
    import torch.multiprocessing as mp
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got, all_sums = q.get(timeout=300)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    for k in range(7):
        assert got[k] == [all_sums[r][k] for r in range(2)], f"round {k}: gathered composites differ from what the ranks rendered"
    assert all_sums[0] != all_sums[1]          # the two ranks really rendered different views


# ---- sort-last composition over NCCL ---------------------------------------------------------------------
def _sort_last_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    from glimpsw_b200 import api, scenes, sharding

    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", device_id=torch.device("cuda", rank))
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    scene = scenes.torus_knot_scene(120, 48, 960, 540, tex_size=128)
    rast = api.Rasterizer(rank)
    rast.set_stream(stream.cuda_stream)
    gscene = rast.upload_scene(scene.meshlets, scene.materials, scene.textures, scene.lights)
    fb = rast.create_framebuffer(scene.width, scene.height)
    fb.clear(0xFF000000, 0.0)
    draws = []
    for node in scene.nodes:                                     # this rank's share of every DrawMeshlets call
        first, count = sharding.meshlets_for_rank(node.meshlet_count, rank, world)
        if count:
            draws.append(dict(offset=node.meshlet_offset + first, count=count, object_to_clip=scene.object_to_clip(node)))
    rast.draw_batch(fb, gscene, draws)
    sharding.composite_framebuffer(fb, dst=None, stream=stream)  # ncclAllReduce(max, int64) on the key buffers, in place
    counters = rast.counters()
    total = torch.tensor([counters["TrianglesProcessed"], counters["TrianglesRasterized"]], device="cuda")
    dist.all_reduce(total)
    if rank == 0:
        depth, ids = fb.download_tiled(1), fb.download_tiled(0)
        rast.resolve(fb, gscene, **scenes.resolve_uniforms(scene, scene.nodes[0]))
        q.put((depth, ids, fb.download_tiled(0), [int(x) for x in total.cpu()]))
    torch.cuda.synchronize()
    dist.barrier()
    rast.destroy()
    dist.destroy_process_group()


def test_sort_last_key_composite_world2():
    """Two GPUs draw disjoint halves of the scene's meshlets; after an NCCL max over their 64-bit key buffers rank 0 holds the
    vis-buffer of the whole scene bit for bit (depth, ids, summed counters), and resolves it within the colour tolerance."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from glimpsw_b200 import scenes
    from oracle import orc
    from helpers import oracle_render
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_sort_last_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    depth, ids, colour, total = q.get(timeout=300)
    for p in procs:
        p.join(timeout=60)
    scene = scenes.torus_knot_scene(120, 48, 960, 540, tex_size=128)
    orc.build()
    ofb, oc = oracle_render(orc, scene)
    n = scene.width * scene.height
    assert np.array_equal(depth, ofb.data[1, :n]) and np.array_equal(ids, ofb.data[0, :n])
    assert total == [int(oc[0]), int(oc[1])]
    orc.resolve(ofb, scene.meshlets, scene.materials, scene.textures, scene.lights, **scenes.resolve_uniforms(scene, scene.nodes[0]))
    err = np.abs(colour.view(np.uint8).astype(np.int32) - ofb.data[0, :n].view(np.uint8).astype(np.int32))
    assert err.max() <= 2


# ---- sort-first: one view split into bands over NCCL ---------------------------------------------------------
def _sort_first_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    from glimpsw_b200 import api, scenes, sharding

    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", device_id=torch.device("cuda", rank))
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    scene = scenes.torus_knot_scene(120, 48, 960, 544, tex_size=128)
    rast = api.Rasterizer(rank)
    rast.set_stream(stream.cuda_stream)
    gscene = rast.upload_scene(scene.meshlets, scene.materials, scene.textures, scene.lights)
    fb = rast.create_framebuffer(scene.width, scene.height)
    fb.set_scissor_rows(*sharding.band_rows(scene.height, rank, world))          # this rank's band of the ONE view
    node = scene.nodes[0]
    batch = rast.create_batch(gscene, [dict(offset=node.meshlet_offset, count=node.meshlet_count, object_to_clip=scene.object_to_clip(node))])
    image = torch.zeros((scene.height, scene.width), dtype=torch.int32, device="cuda")
    frame = rast.make_frame(batch, api.Rasterizer.make_uniforms(**scenes.resolve_uniforms(scene, node)), 0xFF000000, 0.0,
                            pixels_device=image.data_ptr())
    rast.submit_frame(fb, frame)                                  # Clear -> Draw -> Resolve -> GetPixels of the band, into its rows
    sharding.gather_bands(image, rank, world)                     # concatenation, no depth compare
    counters = rast.counters()
    total = torch.tensor([counters["TrianglesProcessed"]], device="cuda")
    dist.all_reduce(total)
    torch.cuda.synchronize()
    # the same exchange folded into the de-tile kernels: every rank stores its band into ONE image in rank 0's memory (NVLink)
    peers = sharding.PeerComposites(scene.height, scene.width, rank, world, slots=2, shared_image=True)
    coll = torch.cuda.Stream()
    p2p_ok = True
    for it in range(3):                                           # three rounds over two slots: the ack path is exercised too
        fb.clear(0xFF000000, 0.0)
        rast.draw_prepared(fb, batch)
        rast.resolve_prebuilt(fb, gscene, api.Rasterizer.make_uniforms(**scenes.resolve_uniforms(scene, node)))
        peers.send(fb, it % 2, stream)
        got = []

        def consume(views):
            with torch.cuda.stream(coll):                         # the reads belong on the collect stream, before the ack
                got.append(views[0].clone())
        peers.collect(rast, it % 2, coll, consume=consume)
        torch.cuda.synchronize()
        if rank == 0:
            p2p_ok = p2p_ok and bool(torch.equal(got[0], image))
    q.put((rank, image.cpu().numpy().view(np.uint32), counters["TrianglesProcessed"], int(total.item()), p2p_ok))
    dist.barrier()
    rast.destroy()
    dist.destroy_process_group()


def test_sort_first_bands_world2():
    """Two GPUs render the upper and the lower band of one view (scissor rows); after the band gather both hold the whole image,
    equal to the oracle's within the colour tolerance and bit-identical on both ranks; each GPU shaded fewer meshlets than the view has."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from glimpsw_b200 import scenes
    from oracle import orc
    from helpers import oracle_render
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_sort_first_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = sorted([q.get(timeout=300) for _ in range(2)], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
    scene = scenes.torus_knot_scene(120, 48, 960, 544, tex_size=128)
    orc.build()
    ofb, oc = oracle_render(orc, scene)
    orc.resolve(ofb, scene.meshlets, scene.materials, scene.textures, scene.lights, **scenes.resolve_uniforms(scene, scene.nodes[0]))
    want = ofb.get_pixels(0)
    assert np.array_equal(got[0][1], got[1][1])
    err = np.abs(got[0][1].view(np.uint8).astype(np.int32) - want.view(np.uint8).astype(np.int32))
    assert err.max() <= 2
    assert got[0][2] < int(oc[0]) and got[1][2] < int(oc[0]) and got[0][3] >= int(oc[0])
    assert got[0][4] and got[1][4]                               # the peer-memory band exchange delivered the same image, three times
