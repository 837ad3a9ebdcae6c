"""Multi-GPU host logic: independent camera views sharded over ranks, composites gathered to rank 0.

The reference has no multi-process mode; the path shards naturally at view/frame granularity
(SURVEY.md §8e): the scene is replicated read-only on every GPU, rank r renders views
{v : v mod world == r} and the finished RGBA8 composites are gathered to rank 0. The only exchange
step is that gather (NCCL over NVLink on GPUs, gloo in the CPU tests); there is no collective on the
render path itself.
"""
from __future__ import annotations

from typing import Callable, List, Optional


def views_for_rank(num_views: int, rank: int, world: int) -> List[int]:
    """Round-robin deal of view indices to ranks (config C5: 64 views over 1/2/4/8 GPUs)."""
    return list(range(rank, num_views, world))


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


class PeerComposites:
    """Composite gather without a collective call: every rank's de-tile kernel (Framebuffer::GetPixels) stores
    its finished view straight into rank 0's buffer through NVLink peer memory.

    The buffer is torch symmetric memory ([slots, world, H, W] int32 on every rank; only rank 0's copy is the
    gather target). Completion and slot reuse are stream-ordered device-side signals:
        sender:  wait ack(slot) -> de-tile into root[slot, rank] -> put ready(slot) to rank 0
        rank 0:  wait ready(slot) from every peer -> (consume) -> put ack(slot) to every peer
    so no host synchronisation and no NCCL kernel sits on the render stream."""

    def __init__(self, height: int, width: int, rank: int, world: int, slots: int = 2):
        import torch
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm_mem

        self.rank, self.world, self.slots = rank, world, slots
        self.buf = symm_mem.empty((slots, world, height, width), dtype=torch.int32, device="cuda")
        self.hdl = symm_mem.rendezvous(self.buf, dist.group.WORLD.group_name)
        self.root = self.hdl.get_buffer(0, self.buf.shape, self.buf.dtype)   # rank 0's buffer, mapped into this process
        self.uses = [0] * slots
        self.local_ready = [torch.cuda.Event() for _ in range(slots)]
        self.collected = [torch.cuda.Event() for _ in range(slots)]

    def dst_ptr(self, slot: int) -> int:
        return self.root[slot, self.rank].data_ptr()

    def before_write(self, slot: int, stream):
        """Sender side, on the render stream: the slot must have been consumed by rank 0."""
        if self.uses[slot] > 0:
            if self.rank == 0:
                stream.wait_event(self.collected[slot])
            else:
                self.hdl.wait_signal(0, channel=self.slots + slot)
        self.uses[slot] += 1

    def after_write(self, slot: int, stream):
        if self.rank == 0:
            self.local_ready[slot].record(stream)
        else:
            self.hdl.put_signal(0, channel=slot)

    def collect(self, slot: int, comm_stream):
        """Rank 0, on a side stream: wait for every view of this slot, then release the slot."""
        import torch
        if self.rank != 0:
            return None
        with torch.cuda.stream(comm_stream):
            comm_stream.wait_event(self.local_ready[slot])
            for src in range(1, self.world):
                self.hdl.wait_signal(src, channel=slot)
            views = self.buf[slot]                   # [world, H, W] — all composites of this round
            for src in range(1, self.world):
                self.hdl.put_signal(src, channel=self.slots + slot)
            self.collected[slot].record(comm_stream)
        return views
