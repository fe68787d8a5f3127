"""Block (up to 3-D) partition of ONE large volume over the GPUs of a box; 1-D slabs (slab.py) are the special case of one
cut axis and share every line of this file.

The 1-D slab split of BASELINE.json configs[2] (21 window starts along the long axis over 8 ranks = 3/3/3/3/3/2/2/2)
caps the speed-up at 2100 / 300 = 7.0x because window starts only come in whole layers of 100 windows.  Cutting the
window-index box along several axes balances better: 2 x 2 x 2 blocks of the 10 x 10 x 21 grid hold at most
5 * 5 * 11 = 275 windows (7.64x), 1 x 2 x 2 blocks on 4 ranks 550 instead of 600.  ``choose_dims`` picks the
factorisation of the world size with the smallest maximum window count (ties: fewer cut axes, i.e. less halo).

Per axis everything is 1-D (``SlabPartition``: window-start ranges, buffer box, owned planes,
halo = planes written beyond the ownership).  Halos are reduced one axis after the other - W, then H, then D: a rank
ships the planes beyond its ownership along the axis over the FULL extent of its buffer in the axes not yet reduced
(and only its owned range in the axes already done) to its +1 neighbour along that axis, which adds them
(``mss_halo_add``).  Contributions for diagonal neighbours travel in two or three hops.  Ranks that differ only in
their coordinate along the axis have identical buffer extents in the other axes, so the exchanged boxes match.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Any, Callable, List, Optional, Sequence, Tuple

import torch

from . import _lib
from .grid import WindowGrid, make_grid


# ---- one axis: contiguous ranges of window starts (the 1-D "z-slab" split of BASELINE.json configs[2]) -------------------

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


# ---- blocks ---------------------------------------------------------------------------------------------------------------

@dataclass
class BlockPartition:
    dims: Tuple[int, int, int]       # ranks along D, H, W; world = product
    axes: List[SlabPartition]        # 1-D partition of every axis (world = dims[a])

    @property
    def world(self) -> int:
        return self.dims[0] * self.dims[1] * self.dims[2]

    def coords(self, rank: int) -> Tuple[int, int, int]:
        pd, ph, pw = self.dims
        return rank // (ph * pw), (rank // pw) % ph, rank % pw

    def rank_of(self, coords: Sequence[int]) -> int:
        return (coords[0] * self.dims[1] + coords[1]) * self.dims[2] + coords[2]

    def n_windows(self, rank: int) -> int:
        c = self.coords(rank)
        n = 1
        for a in range(3):
            n *= self.axes[a].win_hi[c[a]] - self.axes[a].win_lo[c[a]]
        return n

    def box(self, rank: int, what: str) -> Tuple[List[int], List[int]]:
        """Global (lo, hi) per axis of rank's ``"buf"`` (accumulator) or ``"own"`` (finalised) box."""
        c = self.coords(rank)
        lo = [getattr(self.axes[a], what + "_lo")[c[a]] for a in range(3)]
        hi = [getattr(self.axes[a], what + "_hi")[c[a]] for a in range(3)]
        return lo, hi

    def win_box(self, rank: int) -> Tuple[List[int], List[int]]:
        c = self.coords(rank)
        return [self.axes[a].win_lo[c[a]] for a in range(3)], [self.axes[a].win_hi[c[a]] for a in range(3)]


def _factorisations(world: int) -> List[Tuple[int, int, int]]:
    out = []
    for pd in range(1, world + 1):
        if world % pd:
            continue
        for ph in range(1, world // pd + 1):
            if (world // pd) % ph:
                continue
            out.append((pd, ph, world // pd // ph))
    return out


def choose_dims(grid: WindowGrid, world: int) -> Tuple[int, int, int]:
    """Factorisation of ``world`` over (D, H, W) with the smallest maximum window count per rank; ties go to fewer
    cut axes, then to cutting the longest axes."""
    ns = grid.n_starts
    best, best_key = None, None
    for dims in _factorisations(world):
        if any(dims[a] > ns[a] for a in range(3)):
            continue
        load = 1
        for a in range(3):
            load *= -(-ns[a] // dims[a])
        key = (load, sum(1 for p in dims if p > 1), tuple(-dims[a] * ns[a] for a in range(3)))
        if best_key is None or key < best_key:
            best, best_key = dims, key
    if best is None:
        raise ValueError(f"a {ns} window grid cannot be partitioned over {world} ranks")
    return best


def block_partition(grid: WindowGrid, world: int, dims: Optional[Sequence[int]] = None) -> BlockPartition:
    dims = tuple(dims) if dims is not None else choose_dims(grid, world)
    if len(dims) != 3 or dims[0] * dims[1] * dims[2] != world:
        raise ValueError(f"dims {dims} do not multiply to the world size {world}")
    return BlockPartition(dims, [partition(grid, dims[a], axis=a) for a in range(3)])


class PeerAccumulators:
    """Per-rank fp32 accumulators in torch symmetric memory (CUDA peer mappings over NVLink / NVSwitch): a rank's
    ``mss_halo_add_nd`` then reads its neighbour's halo planes straight out of the neighbour's accumulator - the add
    is the transfer, no staging copy, no send/recv, no receive buffer.  One symmetric buffer (sized for the largest
    block of the partition) is allocated and rendezvoused once per (group, size) and reused for every volume."""

    _cache: dict = {}

    def __init__(self, numel: int, device: torch.device, group: Any = None) -> None:
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm

        self.group = group if group is not None else dist.group.WORLD
        self.numel = int(numel)
        self.buf = symm.empty(self.numel, dtype=torch.float32, device=device)
        try:
            self.hdl = symm.rendezvous(self.buf, self.group)
        except Exception:  # noqa: BLE001 - older torch wants the group enabled explicitly first
            symm.enable_symm_mem_for_group(self.group.group_name)
            self.hdl = symm.rendezvous(self.buf, self.group)
        self.rank = dist.get_rank(self.group)

    @classmethod
    def get(cls, numel: int, device: torch.device, group: Any = None) -> "PeerAccumulators":
        key = (id(group), str(device))
        cur = cls._cache.get(key)
        if cur is None or cur.numel < numel:
            cur = cls(numel, device, group)
            cls._cache[key] = cur
        return cur

    def local(self, shape: Sequence[int]) -> torch.Tensor:
        n = 1
        for v in shape:
            n *= int(v)
        return self.buf[:n].view(*shape)

    def remote(self, group_rank: int, shape: Sequence[int]) -> torch.Tensor:
        """The accumulator of ``group_rank`` (its ``local(shape)``) as a tensor of THIS process, backed by peer memory."""
        return self.hdl.get_buffer(group_rank, tuple(int(v) for v in shape), torch.float32, 0)

    def barrier(self) -> None:
        self.hdl.barrier(0)


def acc_shape(part: "BlockPartition", rank: int, n_volumes: int, n_classes: int) -> Tuple[int, int, int, int, int]:
    lo, hi = part.box(rank, "buf")
    ext = [h - l for l, h in zip(lo, hi)]
    return (n_volumes, n_classes, ext[0], ext[1], (ext[2] + 3) // 4 * 4)


def _box_view(t: torch.Tensor, lo: Sequence[int], hi: Sequence[int]) -> torch.Tensor:
    """View of the buffer-local box [lo, hi) of a ``[..., D, H, W]`` tensor."""
    return t[(Ellipsis, slice(lo[0], hi[0]), slice(lo[1], hi[1]), slice(lo[2], hi[2]))]


def local_pass(volume: torch.Tensor, model: Callable[..., torch.Tensor], grid: WindowGrid, part: BlockPartition,
               rank: int, mode: Any = "gaussian", *, sw_batch_size: int = 4, sigma_scale: Any = 0.125, cval: float = 0.0,
               affine: Optional[torch.Tensor] = None, tuple_input: bool = False, tie_tol: float = 1e-5,
               stats: Any = None, time_kernels: bool = False, group_bytes: Optional[int] = None,
               volume_is_block: bool = False, peer_group: Any = False):
    """Everything rank `rank` does before the exchange: block copy, extract -> backbone -> accumulate of its own
    windows into raw weighted sums over its buffer box.  Returns the Stitcher (``.acc`` is the buffer).
    ``volume`` is the full volume, or only this rank's buffer box of it when ``volume_is_block``.  ``peer_group``
    (a process group, or None for the world) places the accumulator in symmetric memory for ``exchange_halos_p2p``."""
    from .importance import importance_map as build_imp
    from .inferer import StitchPlan, Stitcher, _tma_ready

    dev = torch.device("cuda", torch.cuda.current_device())
    nb = volume.shape[0]
    origin, hi = part.box(rank, "buf")
    extent = [h - l for l, h in zip(origin, hi)]
    win_lo, win_hi = part.win_box(rank)
    plan = StitchPlan(grid, dev, nb, win_lo, win_hi, origin, extent)
    src = volume if volume_is_block else _box_view(volume, origin, hi)
    if tuple(src.shape[2:]) != tuple(extent):
        raise ValueError(f"block has spatial shape {tuple(src.shape[2:])}, expected {tuple(extent)}")
    block = _tma_ready(src.to(device=dev, dtype=torch.float32, non_blocking=True).contiguous(), grid, cval)
    imp = build_imp(grid.roi, mode, sigma_scale, dev)
    alloc = None
    if peer_group is not False:
        def alloc(shape):  # collective: every rank reaches its first predictor batch and asks for the same maximum size
            k = shape[1]
            numel = max(int(torch.tensor(acc_shape(part, r, nb, k)).prod()) for r in range(part.world))
            return PeerAccumulators.get(numel, dev, peer_group).local(shape)
    st = Stitcher(plan, imp, fuse=_lib.FUSE_NONE, sw_batch=sw_batch_size, tie_tol=tie_tol, group_bytes=group_bytes,
                  stats=stats, time_kernels=time_kernels, acc_alloc=alloc)
    if stats is not None:
        stats.n_windows = st.total
        stats._near_ties = st.near
    if affine is not None:
        affine = affine.to(dev)
    for _first, n, patches, centers in st.batches(block, cval, vol_origin=origin):
        if sw_batch_size == 1:
            centers = centers.unsqueeze(0)
        with st.timer("predictor"):
            logits = model((patches, centers, affine) if tuple_input else patches)
        st.push(logits, n)
        if stats is not None:
            stats.n_predictor_calls += 1
    st.flush()
    return st


def exchange_halos(acc: torch.Tensor, part: BlockPartition, rank: int, group: Any = None,
                   add_fn: Callable[[torch.Tensor, torch.Tensor], None] = cuda_halo_add,
                   order: Sequence[int] = (2, 1, 0)) -> int:
    """Axis-by-axis nearest-neighbour halo reduction on ``acc[Nb, K, *buffer]``.  Returns bytes received."""
    import torch.distributed as dist

    c = part.coords(rank)
    buf_lo, buf_hi = part.box(rank, "buf")
    own_lo, own_hi = part.box(rank, "own")
    lo = [0, 0, 0]                                    # buffer-local box still carrying live partial sums
    hi = [h - l for l, h in zip(buf_lo, buf_hi)]
    received = 0

    def peer(a: int, step: int) -> int:
        cc = list(c)
        cc[a] += step
        r = part.rank_of(cc)
        return dist.get_global_rank(group, r) if group is not None else r

    for a in order:
        p1, i = part.axes[a], c[a]
        if p1.world > 1:
            send_lo, send_hi = p1.halo(i)
            reqs, keep = [], []

            def do_send() -> None:
                if i + 1 < p1.world and send_hi > send_lo:
                    blo, bhi = list(lo), list(hi)
                    blo[a], bhi[a] = send_lo - buf_lo[a], send_hi - buf_lo[a]
                    buf = _box_view(acc, blo, bhi).contiguous()
                    keep.append(buf)
                    reqs.append(dist.isend(buf, peer(a, +1), group=group))

            early = not p1.halo_depends_on_previous(i)
            if early:
                do_send()
            if i > 0:
                rlo, rhi = p1.halo(i - 1)
                if rhi > rlo:
                    blo, bhi = list(lo), list(hi)
                    blo[a], bhi[a] = rlo - buf_lo[a], rhi - buf_lo[a]
                    view = _box_view(acc, blo, bhi)
                    tmp = torch.empty(view.shape, dtype=acc.dtype, device=acc.device)
                    dist.recv(tmp, peer(a, -1), group=group)
                    add_fn(view, tmp)
                    received += tmp.numel() * tmp.element_size()
            if not early:
                do_send()
            for r in reqs:
                r.wait()
        # from here on only the owned planes along `a` carry anything this rank still needs
        lo[a], hi[a] = own_lo[a] - buf_lo[a], own_hi[a] - buf_lo[a]
    return received


def can_exchange_p2p(part: BlockPartition) -> bool:
    """Peer reads need every halo to come from the immediate -1 neighbour only (no forwarding chain: overlap <= 0.5 of
    the window per cut axis); otherwise the NCCL schedule of ``exchange_halos`` applies."""
    return not any(part.axes[a].halo_depends_on_previous(i) for a in range(3) for i in range(part.axes[a].world))


def exchange_halos_p2p(acc: torch.Tensor, part: BlockPartition, rank: int, group: Any = None,
                       order: Sequence[int] = (2, 1, 0)) -> int:
    """The halo reduction of ``exchange_halos`` with peer memory instead of send/recv: after a device-side barrier every
    rank adds the planes its -1 neighbour wrote beyond its ownership by READING them from the neighbour's accumulator
    (``mss_halo_add_nd`` with a peer-mapped source).  ``acc`` must come from ``PeerAccumulators`` (``local_pass(...,
    peer_group=group)``).  Returns the bytes read over NVLink."""
    import torch.distributed as dist

    if not can_exchange_p2p(part):
        raise ValueError("this partition needs halo forwarding; use exchange_halos (NCCL)")
    peers = PeerAccumulators.get(acc.numel(), acc.device, group)
    nb, k = acc.shape[0], acc.shape[1]
    c = part.coords(rank)
    buf_lo, buf_hi = part.box(rank, "buf")
    own_lo, own_hi = part.box(rank, "own")
    lo = [0, 0, 0]
    hi = [h - l for l, h in zip(buf_lo, buf_hi)]
    moved = 0
    for a in order:
        p1, i = part.axes[a], c[a]
        if p1.world > 1:
            peers.barrier()  # every accumulator holds what its rank has summed so far
            if i > 0:
                rlo, rhi = p1.halo(i - 1)
                if rhi > rlo:
                    cc = list(c)
                    cc[a] -= 1
                    prev = part.rank_of(cc)
                    prev_lo, _ = part.box(prev, "buf")
                    theirs = peers.remote(prev, acc_shape(part, prev, nb, k))
                    slo, shi = list(lo), list(hi)   # same box in the axes the two ranks share ...
                    slo[a], shi[a] = rlo - prev_lo[a], rhi - prev_lo[a]   # ... the halo planes in the neighbour's frame
                    dlo, dhi = list(lo), list(hi)
                    dlo[a], dhi[a] = rlo - buf_lo[a], rhi - buf_lo[a]
                    src = _box_view(theirs, slo, shi)
                    cuda_halo_add(_box_view(acc, dlo, dhi), src)
                    moved += src.numel() * 4
        lo[a], hi[a] = own_lo[a] - buf_lo[a], own_hi[a] - buf_lo[a]
    peers.barrier()  # nobody overwrites an accumulator a neighbour may still be reading
    return moved


def finalize_owned(st: Any, part: BlockPartition, rank: int, tie_tol: float = 1e-5,
                   logits_out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Normalise (weight count of the GLOBAL grid) + argmax of the box rank `rank` owns -> uint8 ``[Nb, own box]``."""
    from .inferer import labels_from_logits

    origin = st.plan.origin
    own_lo, own_hi = part.box(rank, "own")
    lo = [own_lo[a] - origin[a] for a in range(3)]
    hi = [own_hi[a] - origin[a] for a in range(3)]
    w_shift = lo[2] % 4  # the kernel wants a box start that is a multiple of 4 along W: recompute a few voxels, crop
    lo[2] -= w_shift
    buf = labels_from_logits(st.acc, st, tie_tol=tie_tol, normalise=True, box=(lo, hi), logits_out=logits_out)
    lo[2] += w_shift
    return buf[:, lo[0]:hi[0], lo[1]:hi[1], lo[2]:hi[2]]


def sliding_window_infer_blocks(
    volume: torch.Tensor,
    model: Callable[..., torch.Tensor],
    roi: Any = 96,
    overlap: float = 0.5,
    mode: Any = "gaussian",
    *,
    group: Any = None,
    dims: Optional[Sequence[int]] = None,
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
    halo: str = "nccl",
) -> Tuple[torch.Tensor, Tuple[List[int], List[int]], BlockPartition]:
    """Block-partitioned ``sliding_window_infer`` over the ranks of ``group``.

    ``volume`` is the FULL ``[Nb, C, D, H, W]`` volume (host or device, identical on every rank); each rank copies
    only its block to its GPU.  Returns ``(labels, (own_lo, own_hi), partition)`` where ``labels`` holds this rank's
    owned box ``[Nb, ...]`` - or the whole label map on every rank when ``gather=True``.  ``halo``: ``"nccl"``
    (send/recv + add) or ``"p2p"`` (accumulators in symmetric memory, halos read from the neighbour over NVLink).
    """
    import torch.distributed as dist

    rank, world = dist.get_rank(group), dist.get_world_size(group)
    dev = torch.device("cuda", torch.cuda.current_device())
    if volume.dim() != 5:
        raise ValueError("volume must be [N, C, D, H, W]")
    nb = volume.shape[0]
    grid = make_grid(tuple(volume.shape[2:]), roi, overlap)
    if grid.padded:
        raise ValueError("block partitioning expects a volume at least one window large on every axis")
    part = block_partition(grid, world, dims)
    if tuple_input is None:
        tuple_input = affine is not None
    p2p = halo == "p2p" and can_exchange_p2p(part)
    st = local_pass(volume, model, grid, part, rank, mode, sw_batch_size=sw_batch_size, sigma_scale=sigma_scale, cval=cval,
                    affine=affine, tuple_input=tuple_input, tie_tol=tie_tol, stats=stats, time_kernels=time_kernels,
                    group_bytes=group_bytes, peer_group=group if p2p else False)
    with st.timer("halo"):
        halo_bytes = exchange_halos_p2p(st.acc, part, rank, group) if p2p else exchange_halos(st.acc, part, rank, group)
    if stats is not None:
        stats.halo_bytes = halo_bytes
    own = finalize_owned(st, part, rank, tie_tol)
    own_box = part.box(rank, "own")
    if not gather:
        return own, own_box, part
    full = torch.empty((nb,) + grid.image_size, dtype=torch.uint8, device=dev)
    mine = own.contiguous()
    for r in range(world):  # boxes differ in size: one broadcast per owner
        lo, hi = part.box(r, "own")
        buf = mine if r == rank else torch.empty([nb] + [h - l for l, h in zip(lo, hi)], dtype=torch.uint8, device=dev)
        dist.broadcast(buf, dist.get_global_rank(group, r) if group is not None else r, group=group)
        _box_view(full, lo, hi).copy_(buf)
    return full, own_box, part
