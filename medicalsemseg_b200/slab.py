"""Multi-GPU slab partitioning of ONE large volume (BASELINE.json configs[2]: whole-body CT on 1/2/4/8 B200).

The reference is single-process (run_test.py never initialises torch.distributed); this is the B200 form of its
scaling mechanism.  The list of window starts along one axis is cut into contiguous ranges, one per rank.  Rank r

* copies the slab of the input its windows need (``[start_first, start_last + roi)`` along the axis),
* runs extract -> backbone -> accumulate on its own windows only (no communication),
* ships the planes it wrote beyond its ownership (the halo) to rank r+1 with NCCL send/recv, which adds them to its
  own partial sums (``mss_halo_add``) - the only data-path exchange, ~0.7 GB per boundary for 512x512x48 x K=14,
* finalises (weight count from the GLOBAL window grid, argmax) the planes it owns.

Ownership: rank r owns the planes from its first window start up to the next rank's first window start (rank 0
from plane 0, the last rank to the end of the image).  Because window starts increase strictly, rank r's halo
always lies inside rank r+1's buffer; when it reaches past rank r+1's ownership the contribution travels on with
rank r+1's own halo (receive -> add -> send), otherwise all ranks exchange concurrently.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Any, Callable, List, Optional, Sequence, Tuple

import torch

from . import _lib
from .grid import WindowGrid, make_grid


@dataclass
class SlabPartition:
    axis: int
    world: int
    win_lo: List[int]   # first window-start index of each rank along `axis`
    win_hi: List[int]   # one past the last
    buf_lo: List[int]   # global coordinate where each rank's buffer begins
    buf_hi: List[int]
    own_lo: List[int]   # planes each rank finalises
    own_hi: List[int]

    def halo(self, rank: int) -> Tuple[int, int]:
        """Global plane range rank `rank` wrote but does not own (empty for the last rank)."""
        return (self.own_hi[rank], self.buf_hi[rank]) if rank + 1 < self.world else (0, 0)

    def halo_depends_on_previous(self, rank: int) -> bool:
        """Does rank-1's halo reach into the planes `rank` itself has to forward?"""
        if rank == 0 or rank + 1 >= self.world:
            return False
        return self.halo(rank - 1)[1] > self.own_hi[rank]


def split_counts(n: int, world: int) -> List[int]:
    """n window starts over `world` ranks, larger shares first (21 over 8 -> 3,3,3,3,3,2,2,2)."""
    base, extra = divmod(n, world)
    return [base + (1 if r < extra else 0) for r in range(world)]


def partition(grid: WindowGrid, world: int, axis: Optional[int] = None) -> SlabPartition:
    ns = grid.n_starts
    if axis is None:
        axis = max(range(3), key=lambda a: ns[a])
    if ns[axis] < world:
        raise ValueError(f"axis {axis} has {ns[axis]} window starts: cannot partition over {world} ranks")
    counts = split_counts(ns[axis], world)
    lo, hi, acc = [], [], 0
    for c in counts:
        lo.append(acc)
        acc += c
        hi.append(acc)
    st = grid.starts[axis]
    roi = grid.roi[axis]
    buf_lo = [st[l] for l in lo]
    buf_hi = [st[h - 1] + roi for h in hi]
    own_lo = [0] + [st[l] for l in lo[1:]]
    own_hi = own_lo[1:] + [grid.image_size[axis]]
    return SlabPartition(axis, world, lo, hi, buf_lo, buf_hi, own_lo, own_hi)


def _region(t: torch.Tensor, axis: int, lo: int, hi: int) -> torch.Tensor:
    """View of planes [lo, hi) (buffer-local) along spatial `axis` of a [Nb, K, D, H, W] tensor."""
    idx = [slice(None)] * 5
    idx[2 + axis] = slice(lo, hi)
    return t[tuple(idx)]


def cuda_halo_add(dst_view: torch.Tensor, src: torch.Tensor) -> None:
    """``dst_view += src`` with the library kernel, one launch.  Both are (strided) views of up to 5 dimensions whose
    innermost dimension is contiguous - a box of an accumulator; ``src`` may be a peer GPU's memory."""
    if dst_view.shape != src.shape:
        raise ValueError(f"halo shapes differ: {tuple(dst_view.shape)} vs {tuple(src.shape)}")
    if dst_view.dim() > 5 or dst_view.dim() < 1:
        raise ValueError("halo boxes have 1..5 dimensions")
    if dst_view.numel() == 0:
        return
    if dst_view.stride(-1) != 1 or src.stride(-1) != 1:
        raise ValueError("the innermost halo dimension must be contiguous")
    outer = dst_view.dim() - 1
    dims = [1] * (4 - outer) + list(dst_view.shape[:-1])
    dst_s = [0] * (4 - outer) + list(dst_view.stride()[:-1])
    src_s = [0] * (4 - outer) + list(src.stride()[:-1])
    I4 = _lib.c_i64 * 4
    rc = _lib.load().mss_halo_add_nd(dst_view.data_ptr(), I4(*dst_s), src.data_ptr(), I4(*src_s), I4(*dims),
                                     dst_view.shape[-1], torch.cuda.current_stream().cuda_stream)
    _lib.check(rc, "mss_halo_add_nd")


def exchange_halos(acc: torch.Tensor, part: SlabPartition, rank: int, group: Any = None,
                   add_fn: Callable[[torch.Tensor, torch.Tensor], None] = cuda_halo_add) -> int:
    """Nearest-neighbour halo reduction on ``acc[Nb, K, *buffer]`` (buffer-local coordinates).  Returns bytes received.

    ``add_fn(dst_view, src)`` performs ``dst_view += src``; the default is the CUDA kernel, the gloo tests on CPU pass
    a torch stand-in to exercise the schedule."""
    import torch.distributed as dist

    ax, w = part.axis, part.world
    recv_bytes = 0
    send_lo, send_hi = part.halo(rank)
    send_buf = None

    def do_send():
        nonlocal send_buf
        if rank + 1 < w and send_hi > send_lo:
            send_buf = _region(acc, ax, send_lo - part.buf_lo[rank], send_hi - part.buf_lo[rank]).contiguous()
            return dist.isend(send_buf, rank + 1, group=group)
        return None

    reqs = []
    early = not part.halo_depends_on_previous(rank)
    if early:
        r = do_send()
        if r is not None:
            reqs.append(r)
    if rank > 0:
        lo, hi = part.halo(rank - 1)
        if hi > lo:
            view = _region(acc, ax, lo - part.buf_lo[rank], hi - part.buf_lo[rank])
            tmp = torch.empty(view.shape, dtype=acc.dtype, device=acc.device)
            dist.recv(tmp, rank - 1, group=group)
            add_fn(view, tmp)
            recv_bytes = tmp.numel() * tmp.element_size()
    if not early:
        r = do_send()
        if r is not None:
            reqs.append(r)
    for r in reqs:
        r.wait()
    return recv_bytes


def local_pass(volume: torch.Tensor, model: Callable[..., torch.Tensor], grid: WindowGrid, part: SlabPartition,
               rank: int, mode: Any = "gaussian", *, sw_batch_size: int = 4, sigma_scale: Any = 0.125, cval: float = 0.0,
               affine: Optional[torch.Tensor] = None, tuple_input: bool = False, tie_tol: float = 1e-5,
               stats: Any = None, time_kernels: bool = False, group_bytes: Optional[int] = None,
               volume_is_slab: bool = False):
    """Everything rank `rank` does before the exchange: slab copy, extract -> backbone -> accumulate of its own
    windows into raw weighted sums over its buffer box.  Returns the Stitcher (``.acc`` is the buffer).
    ``volume`` is the full volume, or only this rank's slab ``[buf_lo, buf_hi)`` when ``volume_is_slab``."""
    from .importance import importance_map as build_imp
    from .inferer import StitchPlan, Stitcher, _tma_ready

    dev = torch.device("cuda", torch.cuda.current_device())
    ax, nb = part.axis, volume.shape[0]
    origin = [0, 0, 0]
    extent = list(grid.image_size)
    origin[ax] = part.buf_lo[rank]
    extent[ax] = part.buf_hi[rank] - part.buf_lo[rank]
    win_lo, win_hi = [0, 0, 0], list(grid.n_starts)
    win_lo[ax], win_hi[ax] = part.win_lo[rank], part.win_hi[rank]
    plan = StitchPlan(grid, dev, nb, win_lo, win_hi, origin, extent)
    src = volume if volume_is_slab else _region(volume, ax, origin[ax], origin[ax] + extent[ax])
    if tuple(src.shape[2:]) != tuple(extent):
        raise ValueError(f"slab has spatial shape {tuple(src.shape[2:])}, expected {tuple(extent)}")
    slab = _tma_ready(src.to(device=dev, dtype=torch.float32, non_blocking=True).contiguous(), grid, cval)
    imp = build_imp(grid.roi, mode, sigma_scale, dev)
    st = Stitcher(plan, imp, fuse=_lib.FUSE_NONE, sw_batch=sw_batch_size, tie_tol=tie_tol, group_bytes=group_bytes,
                  stats=stats, time_kernels=time_kernels)
    if stats is not None:
        stats.n_windows = st.total
        stats._near_ties = st.near
    if affine is not None:
        affine = affine.to(dev)
    for _first, n, patches, centers in st.batches(slab, cval, vol_origin=origin):
        if sw_batch_size == 1:
            centers = centers.unsqueeze(0)
        with st.timer("predictor"):
            logits = model((patches, centers, affine) if tuple_input else patches)
        st.push(logits, n)
        if stats is not None:
            stats.n_predictor_calls += 1
    st.flush()
    return st


def finalize_owned(st: Any, part: SlabPartition, rank: int, tie_tol: float = 1e-5,
                   logits_out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Normalise (weight count of the GLOBAL grid) + argmax of the planes rank `rank` owns -> uint8 ``[Nb, own box]``."""
    from .inferer import labels_from_logits

    ax = part.axis
    origin, extent = st.plan.origin, st.plan.extent
    lo = [0, 0, 0]
    hi = list(extent)
    lo[ax] = part.own_lo[rank] - origin[ax]
    hi[ax] = part.own_hi[rank] - origin[ax]
    w_shift = lo[2] % 4  # the kernel wants a box start that is a multiple of 4 along W: recompute a few voxels, crop
    lo[2] -= w_shift
    buf = labels_from_logits(st.acc, st, tie_tol=tie_tol, normalise=True, box=(lo, hi), logits_out=logits_out)
    lo[2] += w_shift
    return buf[:, lo[0]:hi[0], lo[1]:hi[1], lo[2]:hi[2]]


def sliding_window_infer_slab(
    volume: torch.Tensor,
    model: Callable[..., torch.Tensor],
    roi: Any = 96,
    overlap: float = 0.5,
    mode: Any = "gaussian",
    *,
    group: Any = None,
    axis: Optional[int] = None,
    sw_batch_size: int = 4,
    sigma_scale: Any = 0.125,
    cval: float = 0.0,
    affine: Optional[torch.Tensor] = None,
    tuple_input: Optional[bool] = None,
    tie_tol: float = 1e-5,
    gather: bool = False,
    stats: Any = None,
    time_kernels: bool = False,
    group_bytes: Optional[int] = None,
) -> Tuple[torch.Tensor, Tuple[int, int], SlabPartition]:
    """Slab-partitioned ``sliding_window_infer`` over the ranks of ``group``.

    ``volume`` is the FULL ``[Nb, C, D, H, W]`` volume (host or device, identical on every rank); each rank copies
    only its slab to its GPU.  Returns ``(labels, (own_lo, own_hi), partition)`` where ``labels`` holds this rank's
    owned planes ``[Nb, ...]`` - or the whole label map on every rank when ``gather=True``.
    """
    import torch.distributed as dist

    rank, world = dist.get_rank(group), dist.get_world_size(group)
    dev = torch.device("cuda", torch.cuda.current_device())
    if volume.dim() != 5:
        raise ValueError("volume must be [N, C, D, H, W]")
    nb = volume.shape[0]
    grid = make_grid(tuple(volume.shape[2:]), roi, overlap)
    if grid.padded:
        raise ValueError("slab partitioning expects a volume at least one window large on every axis")
    part = partition(grid, world, axis)
    ax = part.axis
    if tuple_input is None:
        tuple_input = affine is not None
    st = local_pass(volume, model, grid, part, rank, mode, sw_batch_size=sw_batch_size, sigma_scale=sigma_scale, cval=cval,
                    affine=affine, tuple_input=tuple_input, tie_tol=tie_tol, stats=stats, time_kernels=time_kernels,
                    group_bytes=group_bytes)
    with st.timer("halo"):
        halo_bytes = exchange_halos(st.acc, part, rank, group)
    if stats is not None:
        stats.halo_bytes = halo_bytes
    own = finalize_owned(st, part, rank, tie_tol)
    if not gather:
        return own, (part.own_lo[rank], part.own_hi[rank]), part
    full = torch.empty((nb,) + grid.image_size, dtype=torch.uint8, device=dev)
    mine = own.contiguous()
    for r in range(world):  # variable-size slabs: one broadcast per owner
        shape = [nb] + list(grid.image_size)
        shape[1 + ax] = part.own_hi[r] - part.own_lo[r]
        buf = mine if r == rank else torch.empty(shape, dtype=torch.uint8, device=dev)
        dist.broadcast(buf, dist.get_global_rank(group, r) if group is not None else r, group=group)
        idx = [slice(None)] * 4
        idx[1 + ax] = slice(part.own_lo[r], part.own_hi[r])
        full[tuple(idx)] = buf
    return full, (part.own_lo[rank], part.own_hi[rank]), part
