"""Host-side multi-rank logic of the slab partition on CPU: world_size-2/3 gloo process groups exchange halos of
NumPy-built partial accumulators and must reproduce the single-process weighted sums (the CUDA add kernel is replaced
by a torch stand-in; the schedule, ownership and halo ranges are what is tested)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from medicalsemseg_b200 import slab
from medicalsemseg_b200.grid import make_grid


def test_partition_known_answers():
    g = make_grid((512, 512, 1024), 96, 0.5)
    p = slab.partition(g, 8)
    assert p.axis == 2
    assert [h - l for l, h in zip(p.win_lo, p.win_hi)] == [3, 3, 3, 3, 3, 2, 2, 2]
    assert p.own_lo[0] == 0 and p.own_hi[-1] == 1024
    assert all(p.own_hi[r] == p.own_lo[r + 1] for r in range(7))
    assert [p.halo(r)[1] - p.halo(r)[0] for r in range(7)] == [48] * 7 and p.halo(7) == (0, 0)
    assert not any(p.halo_depends_on_previous(r) for r in range(8))
    for r in range(8):  # every rank's buffer holds all of its windows and its owned planes
        assert p.buf_lo[r] <= p.own_lo[r] or r == 0
        assert p.buf_hi[r] >= p.own_hi[r] or r == 7
    # high overlap: halos reach across more than one neighbour -> forwarding chain
    g2 = make_grid((64, 64, 200), 32, 0.75)
    p2 = slab.partition(g2, 8, axis=2)
    assert any(p2.halo_depends_on_previous(r) for r in range(1, 7))
    with pytest.raises(ValueError):
        slab.partition(make_grid((128, 128, 128), 96, 0.25), 4)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _windows_sum(g, k, rs_seed, win_range_axis, axis, buf_lo, buf_hi):
    """Weighted partial sums of the windows whose index along `axis` is in win_range_axis, on the buffer box."""
    rs = np.random.RandomState(rs_seed)
    ext = list(g.image_size)
    ext[axis] = buf_hi - buf_lo
    acc = np.zeros([1, k] + ext, dtype=np.float32)
    nd, nh, nw = g.n_starts
    for n in range(g.n_windows):
        idx = (n // (nh * nw), (n // nw) % nh, n % nw)
        logits = rs.standard_normal((k,) + g.roi).astype(np.float32)  # drawn for every window: same stream on all ranks
        if not (win_range_axis[0] <= idx[axis] < win_range_axis[1]):
            continue
        s = list(g.window_start(n))
        s[axis] -= buf_lo
        acc[0, :, s[0]:s[0] + g.roi[0], s[1]:s[1] + g.roi[1], s[2]:s[2] + g.roi[2]] += logits
    return acc


def _worker(rank, world, port, shape, roi, overlap, axis, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g = make_grid(shape, roi, overlap)
        p = slab.partition(g, world, axis)
        acc = torch.from_numpy(_windows_sum(g, 2, 7, (p.win_lo[rank], p.win_hi[rank]), p.axis, p.buf_lo[rank], p.buf_hi[rank]))
        slab.exchange_halos(acc, p, rank, None, add_fn=lambda dst, src: dst.add_(src))
        own = slab._region(acc, p.axis, p.own_lo[rank] - p.buf_lo[rank], p.own_hi[rank] - p.buf_lo[rank]).contiguous()
        full = _windows_sum(g, 2, 7, (0, g.n_starts[p.axis]), p.axis, 0, g.image_size[p.axis])
        want = slab._region(torch.from_numpy(full), p.axis, p.own_lo[rank], p.own_hi[rank])
        out[rank] = bool(torch.allclose(own, want, rtol=1e-5, atol=1e-5))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,shape,roi,overlap,axis", [
    (2, (20, 24, 60), 16, 0.5, None),     # the long axis is W
    (2, (60, 24, 20), 16, 0.5, 0),        # split along D
    (3, (16, 40, 24), 16, 0.75, 1),       # high overlap: forwarding chain along H
])
def test_halo_exchange_reproduces_single_process_sums(world, shape, roi, overlap, axis):
    port = _free_port()
    with mp.Manager() as mgr:
        out = mgr.dict()
        mp.spawn(_worker, args=(world, port, shape, roi, overlap, axis, out), nprocs=world, join=True)
        assert dict(out) == {r: True for r in range(world)}


def _gather_worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from medicalsemseg_b200.metrics import DiceMeter, gather_volume_counts
        rs = np.random.RandomState(rank)
        mine = rs.randint(0, 1000, size=(rank + 1, 3, 4)).astype(np.int64)   # rank r holds r + 1 volumes
        allc = gather_volume_counts(torch.from_numpy(mine))
        want = np.concatenate([np.random.RandomState(r).randint(0, 1000, size=(r + 1, 3, 4)).astype(np.int64) for r in range(world)])
        meter = DiceMeter(4)
        meter.add_counts(allc)
        means, m = meter.class_means()
        out[rank] = bool(np.array_equal(allc, want)) and len(meter.per_volume) == sum(r + 1 for r in range(world)) and np.isfinite(m)
    finally:
        dist.destroy_process_group()


def test_per_volume_dice_counts_gather_across_ranks():
    """cfg5: volumes sharded over ranks, per-volume counts gathered so the nan-mean rules of engine/test.py:59-69 apply."""
    port = _free_port()
    with mp.Manager() as mgr:
        out = mgr.dict()
        mp.spawn(_gather_worker, args=(3, port, out), nprocs=3, join=True)
        assert dict(out) == {0: True, 1: True, 2: True}
