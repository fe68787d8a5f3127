"""1-D slab partitioning of ONE large volume (BASELINE.json configs[2]: "z-slab partitioned with halo exchange"): the
special case of ``block.py`` with one cut axis, kept as the entry point the north star names.

The list of window starts along one axis is cut into contiguous ranges, one per rank.  Rank r copies the slab of the input
its windows need, runs extract -> backbone -> accumulate on its own windows only, ships the planes it wrote beyond its
ownership (the halo) to rank r + 1 (NCCL send/recv), which adds them (``mss_halo_add_nd``), and finalises the planes it
owns.  Everything here delegates to ``block.py`` with ``dims = (1, 1, world)`` permuted onto the cut axis; the balanced
alternative for a box with peer access is ``flat.py`` (contiguous window ranges, 263 instead of 300 windows on the busiest
of 8 ranks).
"""
from __future__ import annotations

from typing import Any, Callable, Optional, Tuple

import torch

from . import block
from .block import SlabPartition, _region, cuda_halo_add, partition, split_counts  # noqa: F401  (re-exported)
from .grid import WindowGrid, make_grid


def _as_block(grid: WindowGrid, part: SlabPartition) -> block.BlockPartition:
    dims = [1, 1, 1]
    dims[part.axis] = part.world
    return block.block_partition(grid, part.world, dims)


def exchange_halos(acc: torch.Tensor, part: SlabPartition, rank: int, group: Any = None,
                   add_fn: Callable[[torch.Tensor, torch.Tensor], None] = cuda_halo_add, grid: Optional[WindowGrid] = None) -> int:
    """Nearest-neighbour halo reduction on ``acc[Nb, K, *buffer]`` along the slab axis.  Returns bytes received."""
    bp = block.BlockPartition(tuple(part.world if a == part.axis else 1 for a in range(3)),
                              [part if a == part.axis else _whole_axis(acc, a) for a in range(3)])
    return block.exchange_halos(acc, bp, rank, group, add_fn)


def _whole_axis(acc: torch.Tensor, a: int) -> SlabPartition:
    n = int(acc.shape[2 + a])
    return SlabPartition(a, 1, [0], [1], [0], [n], [0], [n])


def local_pass(volume: torch.Tensor, model: Callable[..., torch.Tensor], grid: WindowGrid, part: SlabPartition, rank: int,
               mode: Any = "gaussian", *, volume_is_slab: bool = False, **kw: Any):
    """Everything rank `rank` does before the exchange (``block.local_pass`` with one cut axis)."""
    return block.local_pass(volume, model, grid, _as_block(grid, part), rank, mode, volume_is_block=volume_is_slab, **kw)


def finalize_owned(st: Any, part: SlabPartition, rank: int, tie_tol: float = 1e-5,
                   logits_out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Normalise (weight count of the GLOBAL grid) + argmax of the planes rank `rank` owns -> uint8 ``[Nb, own box]``."""
    return block.finalize_owned(st, _as_block(st.plan.grid, part), rank, tie_tol, logits_out)


def sliding_window_infer_slab(volume: torch.Tensor, model: Callable[..., torch.Tensor], roi: Any = 96, overlap: float = 0.5,
                              mode: Any = "gaussian", *, group: Any = None, axis: Optional[int] = None,
                              **kw: Any) -> Tuple[torch.Tensor, Tuple[int, int], SlabPartition]:
    """Slab-partitioned ``sliding_window_infer`` over the ranks of ``group``: ``block.sliding_window_infer_blocks`` with the
    world on one axis (default: the axis with the most window starts).  Returns ``(labels, (own_lo, own_hi), partition)``."""
    import torch.distributed as dist

    world = dist.get_world_size(group)
    grid = make_grid(tuple(volume.shape[2:]), roi, overlap)
    if grid.padded:
        raise ValueError("slab partitioning expects a volume at least one window large on every axis")
    part = partition(grid, world, axis)
    dims = [1, 1, 1]
    dims[part.axis] = world
    labels, _own_box, _bp = block.sliding_window_infer_blocks(volume, model, roi, overlap, mode, group=group, dims=dims, **kw)
    rank = dist.get_rank(group)
    return labels, (part.own_lo[rank], part.own_hi[rank]), part
