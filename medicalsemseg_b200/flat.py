"""Flat partition of ONE large volume over the GPUs of a box (BASELINE.json configs[2]: whole-body CT on 1/2/4/8 B200).

The reference is single-process (run_test.py never initialises torch.distributed); this is the B200 form of its scaling
mechanism.  The window list of engine/utils.py:120-125 (C order, D slowest) is cut into one CONTIGUOUS RANGE per rank,
balanced to one window: 2100 windows over 8 ranks are 263/262 each (7.98x of the single-GPU backbone time), where whole
z-layers give 300 (7.0x) and the best box-shaped blocks 275 (7.64x; medicalsemseg_b200/block.py).  Rank r

* copies the planes of the input its windows need (a contiguous D slab: the D layers its range touches),
* runs extract -> backbone -> accumulate on its own windows only, into a ZEROED fp32 buffer over that slab
  (``mss_accumulate_range``: the other windows of those layers do not exist for this buffer) - no communication,
* finishes the box of the volume it OWNS (an equal share of the D planes) with ``mss_finalize_gather``: the kernel
  reads every accumulator that holds an owned voxel - its own and its peers', mapped over NVLink through torch
  symmetric memory - adds them in ascending rank order (= ascending window order between ranks) and takes the argmax.
  The finalise step IS the exchange: no halo buffers, no send/recv, no separate add pass; two device-side barriers
  (all sums written / all peers done reading) are the only synchronisation.

Sums differ from the single-GPU ones only in association at the rank boundaries (partial sums of consecutive window
ranges are added); labels agree bit for bit outside counted near-ties.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Any, Callable, List, Optional, Sequence, Tuple

import torch

from . import _lib
from .grid import WindowGrid, make_grid


@dataclass
class FlatPartition:
    world: int
    ranges: List[Tuple[int, int]]      # windows [n0, n1) of every rank (C-order window index of the whole grid)
    layers: List[Tuple[int, int]]      # D layers [id_lo, id_hi) the range touches
    buf_lo: List[int]                  # D planes [buf_lo, buf_hi) of every rank's accumulator / input slab
    buf_hi: List[int]
    own_lo: List[int]                  # D planes every rank finalises
    own_hi: List[int]

    def n_windows(self, rank: int) -> int:
        return self.ranges[rank][1] - self.ranges[rank][0]

    def contributors(self, rank: int) -> List[int]:
        """Ranks (ascending) whose accumulator box meets the planes `rank` owns."""
        return [q for q in range(self.world) if self.buf_lo[q] < self.own_hi[rank] and self.own_lo[rank] < self.buf_hi[q]]


def flat_partition(grid: WindowGrid, world: int) -> FlatPartition:
    n = grid.n_windows
    nd, nh, nw = grid.n_starts
    d = grid.image_size[0]
    if n < world:
        raise ValueError(f"{n} windows cannot be partitioned over {world} ranks")
    if d < world:
        raise ValueError(f"{d} planes cannot be owned by {world} ranks")
    base, extra = divmod(n, world)
    ranges, acc = [], 0
    for r in range(world):
        c = base + (1 if r < extra else 0)
        ranges.append((acc, acc + c))
        acc += c
    per_layer = nh * nw
    layers = [(n0 // per_layer, (n1 - 1) // per_layer + 1) for n0, n1 in ranges]
    buf_lo = [grid.starts[0][lo] for lo, _hi in layers]
    buf_hi = [grid.starts[0][hi - 1] + grid.roi[0] for _lo, hi in layers]
    own_lo = [r * d // world for r in range(world)]
    own_hi = own_lo[1:] + [d]
    return FlatPartition(world, ranges, layers, buf_lo, buf_hi, own_lo, own_hi)


def acc_shape(grid: WindowGrid, part: FlatPartition, rank: int, n_classes: int) -> Tuple[int, int, int, int, int]:
    _d, h, w = grid.image_size
    return (1, n_classes, part.buf_hi[rank] - part.buf_lo[rank], h, (w + 3) // 4 * 4)


def local_pass(volume: torch.Tensor, model: Callable[..., torch.Tensor], grid: WindowGrid, part: FlatPartition, rank: int,
               mode: Any = "gaussian", *, sw_batch_size: int = 4, sigma_scale: Any = 0.125, cval: float = 0.0,
               affine: Optional[torch.Tensor] = None, tuple_input: bool = False, stats: Any = None,
               time_kernels: bool = False, group_bytes: Optional[int] = None, volume_is_slab: bool = False,
               acc_alloc: Optional[Callable[[Tuple[int, ...]], torch.Tensor]] = None):
    """Everything rank `rank` does before the finalise: slab copy, extract -> backbone -> accumulate of its own window range
    into raw weighted sums over its (zeroed) slab buffer.  Returns the Stitcher (``.acc`` is the buffer).
    ``volume`` is the full ``[1, C, D, H, W]`` volume, or only the planes ``[buf_lo, buf_hi)`` when ``volume_is_slab``."""
    from .importance import importance_map as build_imp
    from .inferer import StitchPlan, Stitcher, _tma_ready

    if volume.shape[0] != 1:
        raise ValueError("the flat partition takes one volume at a time (N = 1)")
    dev = torch.device("cuda", torch.cuda.current_device())
    nd, nh, nw = grid.n_starts
    id_lo, id_hi = part.layers[rank]
    origin = (part.buf_lo[rank], 0, 0)
    extent = (part.buf_hi[rank] - part.buf_lo[rank], grid.image_size[1], grid.image_size[2])
    plan = StitchPlan(grid, dev, 1, (id_lo, 0, 0), (id_hi, nh, nw), origin, extent)
    src = volume if volume_is_slab else volume[:, :, origin[0]:origin[0] + extent[0]]
    if tuple(src.shape[2:]) != tuple(extent):
        raise ValueError(f"slab has spatial shape {tuple(src.shape[2:])}, expected {tuple(extent)}")
    slab = _tma_ready(src.to(device=dev, dtype=torch.float32, non_blocking=True).contiguous(), grid, cval)
    imp = build_imp(grid.roi, mode, sigma_scale, dev)
    n0, n1 = part.ranges[rank]
    first = n0 - id_lo * nh * nw  # position of the range inside the box's own enumeration
    st = Stitcher(plan, imp, fuse=_lib.FUSE_NONE, sw_batch=sw_batch_size, group_bytes=group_bytes, stats=stats,
                  time_kernels=time_kernels, acc_alloc=acc_alloc, own_range=(first, n1 - n0))
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


def finalize_owned(grid: WindowGrid, part: FlatPartition, rank: int, accs: Sequence[Optional[torch.Tensor]], imp: torch.Tensor,
                   *, tie_tol: float = 1e-5, return_logits: bool = False, near: Optional[torch.Tensor] = None,
                   stats: Any = None) -> Tuple[torch.Tensor, Optional[torch.Tensor]]:
    """``mss_finalize_gather`` for the planes `rank` owns.  ``accs[q]`` is rank q's accumulator ``[1, K, slab_q, H, Wp]`` as a
    tensor of THIS process (its own buffer, or a peer's mapped through symmetric memory) for every q in
    ``part.contributors(rank)``; other entries are ignored.  Returns ``(labels uint8 [1, own planes, H, W], logits or None)``."""
    lib = _lib.load()
    src = part.contributors(rank)
    first = accs[src[0]]
    dev, k = first.device, int(first.shape[1])
    _d, h, w = grid.image_size
    own = (part.own_hi[rank] - part.own_lo[rank], h, w)
    wp = (w + 3) // 4 * 4
    lay = _lib.Layout()
    lay.image = _lib.I3(*grid.image_size)
    lay.roi = _lib.I3(*grid.roi)
    lay.n_starts = _lib.I3(*grid.n_starts)
    lay.win_lo = _lib.I3(0, 0, 0)
    lay.win_hi = _lib.I3(*grid.n_starts)
    lay.origin = _lib.I3(part.own_lo[rank], 0, 0)
    lay.extent = _lib.I3(*own)
    lay.pitch_w = wp
    lay.n_volumes = 1
    lay.n_classes = k
    table_host = grid.table
    table_dev = _device_table(grid, dev)
    lay.table_host = table_host.ctypes.data
    lay.table_dev = table_dev.data_ptr()
    labels = torch.empty((1,) + own, dtype=torch.uint8, device=dev)
    logits = torch.empty((1, k, own[0], own[1], wp), dtype=torch.float32, device=dev) if return_logits else None
    if near is None:
        near = torch.zeros(1, dtype=torch.int64, device=dev)
    ptrs = (C.c_void_p * len(src))(*[accs[q].data_ptr() for q in src])
    I3n = C.c_int32 * (3 * len(src))
    origins = I3n(*[v for q in src for v in (part.buf_lo[q], 0, 0)])
    extents = I3n(*[v for q in src for v in (part.buf_hi[q] - part.buf_lo[q], h, w)])
    pitches = (C.c_int32 * len(src))(*[int(accs[q].shape[-1]) for q in src])
    for q in src:
        if tuple(accs[q].shape) != acc_shape(grid, part, q, k) or not accs[q].is_contiguous():
            raise ValueError(f"accumulator of rank {q} has shape {tuple(accs[q].shape)}, expected {acc_shape(grid, part, q, k)}")
    rc = lib.mss_finalize_gather(C.byref(lay), len(src), ptrs, origins, extents, pitches, imp.data_ptr(), labels.data_ptr(), w,
                                 None if logits is None else logits.data_ptr(), float(tie_tol), near.data_ptr(),
                                 torch.cuda.current_stream().cuda_stream)
    _lib.check(rc, "mss_finalize_gather")
    if stats is not None:
        stats.gpu_launches += 1
    return labels, (None if logits is None else logits[..., :w])


_TABLES: dict = {}


def _device_table(grid: WindowGrid, dev: torch.device) -> torch.Tensor:
    key = (id(grid.table), str(dev))
    hit = _TABLES.get(key)
    if hit is None:
        if len(_TABLES) > 16:
            _TABLES.clear()
        hit = (torch.from_numpy(grid.table).to(dev), grid.table)  # keep the host table alive with its device copy
        _TABLES[key] = hit
    return hit[0]


def gather_slabs(own: torch.Tensor, part: FlatPartition, rank: int, group: Any = None) -> torch.Tensor:
    """The whole label map ``[1, D, H, W]`` on every rank from the owned slabs: ONE all_gather of equally sized (padded)
    slabs instead of a broadcast per owner.  Works on CUDA (NCCL) and CPU (gloo) tensors."""
    import torch.distributed as dist

    world = part.world
    dmax = max(hi - lo for lo, hi in zip(part.own_lo, part.own_hi))
    _n, _d, h, w = own.shape
    send = own.new_zeros((1, dmax, h, w))
    send[:, : own.shape[1]] = own
    out = own.new_empty((world, 1, dmax, h, w))
    dist.all_gather_into_tensor(out.view(-1), send.view(-1), group=group)
    return torch.cat([out[r, :, : part.own_hi[r] - part.own_lo[r]] for r in range(world)], dim=1)


def sliding_window_infer_flat(
    volume: torch.Tensor,
    model: Callable[..., torch.Tensor],
    roi: Any = 96,
    overlap: float = 0.5,
    mode: Any = "gaussian",
    *,
    group: Any = None,
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
    volume_is_slab: bool = False,
    spatial: Optional[Sequence[int]] = None,
) -> Tuple[torch.Tensor, Tuple[int, int], FlatPartition]:
    """``sliding_window_infer`` of ONE volume over the ranks of ``group`` (one process per GPU, NVLink peer memory).

    ``volume`` is the full ``[1, C, D, H, W]`` volume (host or device, identical on every rank; each rank copies only the
    planes it needs) - or, with ``volume_is_slab`` and ``spatial`` = the full (D, H, W), just this rank's planes
    ``[buf_lo, buf_hi)``.  Returns ``(labels, (own_lo, own_hi), partition)``: this rank's owned planes ``[1, own, H, W]``, or
    the whole label map on every rank when ``gather=True``."""
    import torch.distributed as dist

    from .block import PeerAccumulators
    from .importance import importance_map as build_imp

    rank, world = dist.get_rank(group), dist.get_world_size(group)
    dev = torch.device("cuda", torch.cuda.current_device())
    if volume.dim() != 5 or volume.shape[0] != 1:
        raise ValueError("volume must be [1, C, D, H, W]")
    full = tuple(spatial) if spatial is not None else tuple(volume.shape[2:])
    grid = make_grid(full, roi, overlap)
    if grid.padded:
        raise ValueError("the flat partition expects a volume at least one window large on every axis")
    part = flat_partition(grid, world)
    if tuple_input is None:
        tuple_input = affine is not None
    peers_box: List[Any] = []

    def alloc(shape):  # collective on first use: every rank asks for the largest accumulator of the partition
        k = shape[1]
        numel = max(int(torch.tensor(acc_shape(grid, part, r, k)).prod()) for r in range(world))
        peers = PeerAccumulators.get(numel, dev, group)
        peers_box.append(peers)
        return peers.local(shape)

    st = local_pass(volume, model, grid, part, rank, mode, sw_batch_size=sw_batch_size, sigma_scale=sigma_scale, cval=cval,
                    affine=affine, tuple_input=tuple_input, stats=stats, time_kernels=time_kernels, group_bytes=group_bytes,
                    volume_is_slab=volume_is_slab, acc_alloc=alloc)
    peers = peers_box[0]
    k = int(st.acc.shape[1])
    imp = build_imp(grid.roi, mode, sigma_scale, dev)
    with st.timer("exchange+finalize"):
        peers.barrier()  # every accumulator holds its rank's sums
        accs: List[Optional[torch.Tensor]] = [None] * world
        for q in part.contributors(rank):
            accs[q] = st.acc if q == rank else peers.remote(q, acc_shape(grid, part, q, k))
        own, _ = finalize_owned(grid, part, rank, accs, imp, tie_tol=tie_tol, near=st.near, stats=stats)
        peers.barrier()  # nobody clears an accumulator a peer may still be reading
    if stats is not None:
        stats.halo_bytes = sum(4 * k * (min(part.buf_hi[q], part.own_hi[rank]) - max(part.buf_lo[q], part.own_lo[rank])) *
                               grid.image_size[1] * grid.image_size[2] for q in part.contributors(rank) if q != rank)
    if gather:
        return gather_slabs(own, part, rank, group), (part.own_lo[rank], part.own_hi[rank]), part
    return own, (part.own_lo[rank], part.own_hi[rank]), part
