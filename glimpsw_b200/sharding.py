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
