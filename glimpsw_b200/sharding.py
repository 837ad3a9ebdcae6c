"""Multi-GPU host logic: independent camera views sharded over ranks, composites gathered to rank 0.

The reference has no multi-process mode; the path shards naturally at view/frame granularity
(SURVEY.md §8e): the scene is replicated read-only on every GPU, rank r renders views
{v : v mod world == r} — or, when the views' costs are known, its share of a cost-aware deal
(deal_views_by_cost) — and the finished RGBA8 composites are gathered to rank 0. The only exchange
step is that gather (NCCL over NVLink on GPUs, gloo in the CPU tests); there is no collective on the
render path itself.
"""
from __future__ import annotations

from typing import Callable, List, Optional, Sequence


def views_for_rank(num_views: int, rank: int, world: int) -> List[int]:
    """Round-robin deal of view indices to ranks (config C5: 64 views over 1/2/4/8 GPUs)."""
    return list(range(rank, num_views, world))


def deal_views_by_cost(costs: Sequence[float], world: int) -> List[List[int]]:
    """Cost-aware deal of len(costs) views to `world` ranks: longest processing time first, every rank gets the same number of
    views (the first len % world ranks one more — the composite exchange runs in rounds of one view per rank), ties and equal loads
    go to the lower rank, so every rank computes the same deal from the same cost vector. Each rank's list is ascending.
    On the bench batch the views cost 191..283 us; `v mod N` leaves the slowest of 8 ranks 7 % over the mean, this deal 0.3 %."""
    n = len(costs)
    cap = [n // world + (1 if r < n % world else 0) for r in range(world)]
    load = [0.0] * world
    out: List[List[int]] = [[] for _ in range(world)]
    for v in sorted(range(n), key=lambda v: (-float(costs[v]), v)):
        r = min((r for r in range(world) if len(out[r]) < cap[r]), key=lambda r: (load[r], r))
        load[r] += float(costs[v])
        out[r].append(v)
    return [sorted(o) for o in out]


def rounds(num_views: int, world: int) -> int:
    """Number of gather rounds: every rank renders at most one view per round."""
    return (num_views + world - 1) // world


def render_views_sharded(num_views: int, render_view: Callable[[int], "torch.Tensor"], rank: int, world: int,
                         gather: bool = True) -> Optional[list]:
    """Renders this rank's views with `render_view(view_index) -> 2-D tensor` and gathers them to rank 0.

    Returns, on rank 0, the list of all `num_views` composites in view order (None elsewhere).
    Ranks that run out of views in the last round contribute a zero image that is dropped."""
    import torch
    import torch.distributed as dist

    mine = views_for_rank(num_views, rank, world)
    out = [None] * num_views if rank == 0 else None
    template = None
    for rnd in range(rounds(num_views, world)):
        img = render_view(mine[rnd]) if rnd < len(mine) else None
        if img is not None:
            template = img
        if img is None:
            img = torch.zeros_like(template) if template is not None else None
        if not gather:
            if rank == 0 and rnd < len(mine):
                out[mine[rnd]] = img
            continue
        if world == 1:
            out[mine[rnd]] = img
            continue
        bucket = [torch.empty_like(img) for _ in range(world)] if rank == 0 else None
        dist.gather(img, bucket, dst=0)
        if rank == 0:
            for r in range(world):
                v = rnd * world + r
                if v < num_views:
                    out[v] = bucket[r]
    return out


class SlotProtocol:
    """The flow control of the composite exchange, as data: which flag a kernel waits for and which it raises.

    Flags are use counts that only grow. Every rank holds the same array layout (only some entries are live on each):
        ready[slot, src]  (index slot * world + src)      lives on rank 0, raised by src after its last store
        ack[slot]         (index slots * world + slot)     lives on every producer, raised by rank 0's collect
        scratch           (last index)                     sink of the wait-only collect launch
    The n-th use of a slot by producer r: wait ack[slot] >= n - 1 (nothing for n = 1), store the view, raise
    ready[slot, r] = n. The n-th collect of the slot on rank 0: wait ready[slot, r] >= n for all r, [consume,] raise
    ack[slot] = n on every producer. tests/test_sharding_cpu.py runs this plan under random interleavings."""

    def __init__(self, world: int, slots: int):
        self.world, self.slots = world, slots
        self.num_flags = slots * world + slots + 1
        self.uses = [0] * slots

    def ready_index(self, slot: int, src: int) -> int:
        return slot * self.world + src

    def ack_index(self, slot: int) -> int:
        return self.slots * self.world + slot

    @property
    def scratch_index(self) -> int:
        return self.num_flags - 1

    def next_use(self, slot: int) -> int:
        self.uses[slot] += 1
        return self.uses[slot]

    def send_plan(self, rank: int, slot: int, n: int):
        """(wait_index | None, wait_value, signal_index, signal_value) for producer `rank`'s n-th use of `slot`."""
        return (self.ack_index(slot) if n > 1 else None, n - 1, self.ready_index(slot, rank), n)

    def collect_plan(self, slot: int, n: int):
        """(ready indices on rank 0, expected value, ack index on each producer, ack value) for the n-th collect."""
        return ([self.ready_index(slot, r) for r in range(1, self.world)], n, self.ack_index(slot), n)


class PeerComposites:
    """Composite gather without a collective call: every rank's de-tile kernel (Framebuffer::GetPixels) stores
    its finished view straight into rank 0's buffer through NVLink peer memory, and the flow control rides in the
    same kernels (swrb_fb_send_pixels / swrb_peer_collect) instead of separate signal launches.

    Two symmetric allocations (torch symmetric memory, identical on every rank): the image buffer
    [slots, world, H, W] int32 (only rank 0's copy is the gather target) and a flag array of uint64 use counts:
        ready[slot, src]  in rank 0's memory, raised by src's send kernel after its last store (system-scope fence)
        ack[slot]         in every producer's memory, raised by rank 0's collect kernel when the slot may be reused
    Per frame a producer launches ONE kernel (wait ack -> de-tile into root[slot, rank] -> raise ready) and rank 0
    launches its local de-tile plus ONE collect kernel (wait for all ready flags -> raise all acks); no host
    synchronisation and no kernel on the render stream. Counts only grow, so nothing is reset across the link."""

    def __init__(self, height: int, width: int, rank: int, world: int, slots: int = 2, shared_image: bool = False):
        """shared_image: every rank stores into the SAME image of a slot — the sort-first split, where each rank's framebuffer has
        scissor rows and its send moves only its own band (swrb_fb_set_scissor_rows); the flags work as for whole views."""
        import torch
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm_mem

        self.rank, self.world, self.slots = rank, world, slots
        self.proto = SlotProtocol(world, slots)
        group = dist.group.WORLD.group_name
        self.shared_image = shared_image
        self.buf = symm_mem.empty((slots, 1 if shared_image else world, height, width), dtype=torch.int32, device="cuda")
        self.hdl = symm_mem.rendezvous(self.buf, group)
        self.root = self.hdl.get_buffer(0, self.buf.shape, self.buf.dtype)   # rank 0's buffer, mapped into this process
        self.num_flags = self.proto.num_flags            # ready[slot, src] | ack[slot] | one scratch word (see SlotProtocol)
        self.flags = symm_mem.empty((self.num_flags,), dtype=torch.int64, device="cuda")
        self.flags.zero_()
        self.fhdl = symm_mem.rendezvous(self.flags, group)
        self.flag_views = [self.fhdl.get_buffer(r, self.flags.shape, self.flags.dtype) for r in range(world)]
        torch.cuda.synchronize()
        dist.barrier()                                   # every rank's flags are zero before anyone raises one
        self.uses = self.proto.uses
        self.local_ready = [torch.cuda.Event() for _ in range(slots)]
        self.collected = [torch.cuda.Event() for _ in range(slots)]

    def dst_ptr(self, slot: int) -> int:
        return self.root[slot, 0 if self.shared_image else self.rank].data_ptr()

    def ready_ptr(self, slot: int, src: int) -> int:
        """ready[slot, src] in rank 0's memory."""
        return self.flag_views[0].data_ptr() + 8 * self.proto.ready_index(slot, src)

    def ack_ptr(self, slot: int, owner: int) -> int:
        """ack[slot] in `owner`'s memory."""
        return self.flag_views[owner].data_ptr() + 8 * self.proto.ack_index(slot)

    def send(self, fb, slot: int, comm_stream, layer: int = 0):
        """This rank's finished view -> root[slot, rank], on `comm_stream`. Returns the use count of the slot."""
        n = self.proto.next_use(slot)
        if self.rank == 0:
            if n > 1:
                comm_stream.wait_event(self.collected[slot])
            fb.get_pixels_device(layer, self.dst_ptr(slot), cuda_stream=comm_stream.cuda_stream)
            self.local_ready[slot].record(comm_stream)
        else:
            wait_index, wait_value, signal_index, signal_value = self.proto.send_plan(self.rank, slot, n)
            mine, root = self.flag_views[self.rank].data_ptr(), self.flag_views[0].data_ptr()
            fb.send_pixels(layer, self.dst_ptr(slot), comm_stream.cuda_stream,
                           wait_flag=0 if wait_index is None else mine + 8 * wait_index, wait_value=wait_value,
                           signal_flag=root + 8 * signal_index, signal_value=signal_value)
        return n

    def collect(self, rast, slot: int, coll_stream, consume=None):
        """Rank 0, on a side stream: wait for every view of this use of the slot, [consume them,] then release the slot.

        Without `consume` the slot is released by the same kernel that saw the last ready flag (one launch; what the
        benchmarks do — they only need the composites to have ARRIVED). A caller that reads the views passes
        `consume(views)`: it is called between a wait-only launch and an ack-only launch and must enqueue its reads on
        `coll_stream`, so no producer can overwrite the slot before they are done. Returns the [world, H, W] views."""
        if self.rank != 0:
            return None
        coll_stream.wait_event(self.local_ready[slot])
        n, views = self.uses[slot], self.buf[slot]
        acks = [self.ack_ptr(slot, r) for r in range(1, self.world)]
        if self.world > 1 and consume is None:
            rast.peer_collect(coll_stream.cuda_stream, self.ready_ptr(slot, 1), self.world - 1, n, acks, n)
        elif self.world > 1:
            scratch = self.flag_views[0].data_ptr() + 8 * self.proto.scratch_index
            rast.peer_collect(coll_stream.cuda_stream, self.ready_ptr(slot, 1), self.world - 1, n, [scratch] * (self.world - 1), n)   # wait only
            consume(views)
            rast.peer_collect(coll_stream.cuda_stream, self.ready_ptr(slot, 1), self.world - 1, 0, acks, n)                           # ack only
        elif consume is not None:
            consume(views)
        self.collected[slot].record(coll_stream)
        return views


# ---- sort-last composition -----------------------------------------------------------------------------
KEY_SEED_LOW = 0xFFFFFFFF
KEY_ID_BASE = 0xFFFFFFFE


def meshlets_for_rank(num_meshlets: int, rank: int, world: int):
    """Contiguous share [first, first + count) of a draw's meshlets for `rank` (sort-last: every GPU draws a subset of the scene)."""
    per = (num_meshlets + world - 1) // world
    first = min(rank * per, num_meshlets)
    return first, min(per, num_meshlets - first)


def key_rank(surface_id):
    """Draw-order rank of an accepted (unclipped) triangle from its surface id (csrc/common.cuh::key_rank):
    meshlet * 256 + packet * 32 + lane."""
    meshlet, prim = surface_id >> 7, surface_id & 127
    return (meshlet << 8) | ((prim >> 4) << 5) | (prim & 15)


def keys_from_layers(depth_bits, ids):
    """The key buffer a draw from a cleared (depth 0) framebuffer leaves behind, rebuilt from its depth / id layers
    (numpy uint32 arrays): depth << 32 | (0xFFFFFFFE - rank) where something was drawn, depth << 32 | 0xFFFFFFFF elsewhere."""
    import numpy as np
    depth_bits = np.asarray(depth_bits, dtype=np.uint64)
    low = np.where(depth_bits != 0, np.uint64(KEY_ID_BASE) - key_rank(np.asarray(ids, dtype=np.uint64)), np.uint64(KEY_SEED_LOW))
    return ((depth_bits << np.uint64(32)) | low).astype(np.uint64)


def layers_from_keys(keys, clear_color: int):
    """(depth bits, surface ids) a key buffer unpacks to (k_keys_unpack)."""
    import numpy as np
    keys = np.asarray(keys, dtype=np.uint64)
    low = (keys & np.uint64(0xFFFFFFFF)).astype(np.uint64)
    rank = np.uint64(KEY_ID_BASE) - low
    ids = ((rank >> np.uint64(8)) << np.uint64(7)) | ((rank >> np.uint64(1)) & np.uint64(0x70)) | (rank & np.uint64(15))
    seed = low == np.uint64(KEY_SEED_LOW)
    return (keys >> np.uint64(32)).astype(np.uint32), np.where(seed, np.uint64(clear_color), ids).astype(np.uint32)


def composite_keys(keys_tensor, dst=None):
    """Element-wise maximum of every rank's key buffer (an int64 view: keys of reverse-Z depths in (0, 1] are positive):
    all-reduce, or reduce to rank `dst`. In place."""
    import torch.distributed as dist
    if dist.get_world_size() == 1:
        return
    if dst is None:
        dist.all_reduce(keys_tensor, op=dist.ReduceOp.MAX)
    else:
        dist.reduce(keys_tensor, dst=dst, op=dist.ReduceOp.MAX)


def composite_framebuffer(fb, dst=None, stream=None):
    """Sort-last composite of a framebuffer whose vis-buffer draw is pending in its key buffer (api.Framebuffer): NCCL max over
    the ranks' keys, in place, on `stream` (a torch stream that is ordered after the draw; default: the current one)."""
    import torch

    class _Raw:
        def __init__(self, ptr, n):
            self.__cuda_array_interface__ = {"shape": (n,), "typestr": "<i8", "data": (ptr, False), "version": 3}
    ptr, n = fb.keys_device()
    t = torch.as_tensor(_Raw(ptr, n), device="cuda")
    if stream is not None:
        with torch.cuda.stream(stream):
            composite_keys(t, dst)
    else:
        composite_keys(t, dst)
    fb.keys_touched()


# ---- sort-first: one view split into horizontal bands, one per GPU ------------------------------------------
def band_rows(height: int, rank: int, world: int, align: int = 128):
    """Rows [y0, y1) of a `height`-row framebuffer that `rank` of `world` renders: `world` contiguous bands whose boundaries sit on
    multiples of `align` rows (128 = the reference's bin rows, Rasterizer.h BinSize) while that keeps the largest band within 13 % of
    height / world; the alignment is halved otherwise, down to 8 rows (what swrb_fb_set_scissor_rows accepts). The bands tile the
    framebuffer exactly."""
    if world < 1 or not 0 <= rank < world:
        raise ValueError("rank outside the world")
    while True:
        edges = [min(int(round(height * r / world / align)) * align, height) for r in range(world)] + [height]
        sizes = [b - a for a, b in zip(edges, edges[1:])]
        if min(sizes) > 0 and (max(sizes) * world <= 1.13 * height or align == 8):
            return edges[rank], edges[rank + 1]
        if align == 8:
            raise ValueError(f"{height} rows cannot be split into {world} bands of multiples of 8 rows")
        align //= 2


def gather_bands(image, rank: int, world: int, bands=None, dst=None):
    """Sort-first composite: every rank has written its own band of `image` (a [H, W] device tensor, row-major — GetPixels
    under a scissor fills exactly those rows); afterwards rank `dst` (or every rank when dst is None) holds all bands. No depth
    compare: the bands are disjoint. Bands of equal size go through one in-place all-gather, ragged ones through broadcasts."""
    import torch.distributed as dist
    if world == 1:
        return
    h = image.shape[0]
    bands = bands or [band_rows(h, r, world) for r in range(world)]
    sizes = {b[1] - b[0] for b in bands}
    if dst is None and len(sizes) == 1 and bands[0][0] == 0 and bands[-1][1] == h:
        dist.all_gather_into_tensor(image, image[bands[rank][0]:bands[rank][1]])
        return
    if dst is None:
        for r, (y0, y1) in enumerate(bands):
            dist.broadcast(image[y0:y1], src=r)
        return
    if rank == dst:
        reqs = [dist.irecv(image[y0:y1], src=r) for r, (y0, y1) in enumerate(bands) if r != dst]
        for q in reqs:
            q.wait()
    else:
        y0, y1 = bands[rank]
        dist.send(image[y0:y1], dst=dst)
