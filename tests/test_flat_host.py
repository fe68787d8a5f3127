"""Host logic of the flat multi-GPU partition (medicalsemseg_b200/flat.py): runs without a GPU; the collective that
reassembles the label map is exercised with gloo, world_size 2 and 3."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from medicalsemseg_b200 import flat
from medicalsemseg_b200.grid import make_grid


def _covering(grid, d):
    """window indices (C order) covering plane d"""
    nd, nh, nw = grid.n_starts
    ids = [i for i, s in enumerate(grid.starts[0]) if s <= d < s + grid.roi[0]]
    return [(i * nh + j) * nw + k for i in ids for j in range(nh) for k in range(nw)]


@pytest.mark.parametrize("shape,roi,overlap,world", [
    ((512, 512, 1024), 96, 0.5, 8), ((512, 512, 1024), 96, 0.5, 4), ((512, 512, 1024), 96, 0.5, 2),
    ((240, 240, 155), 96, 0.5, 8), ((70, 40, 52), 16, 0.5, 3), ((64, 33, 47), (24, 16, 16), 0.25, 5),
    ((40, 36, 44), 16, 0.75, 7),
])
def test_flat_partition_properties(shape, roi, overlap, world):
    grid = make_grid(shape, roi, overlap)
    part = flat.flat_partition(grid, world)
    n = grid.n_windows
    # the ranges tile the window list, balanced to one window
    assert part.ranges[0][0] == 0 and part.ranges[-1][1] == n
    assert all(part.ranges[r][1] == part.ranges[r + 1][0] for r in range(world - 1))
    sizes = [part.n_windows(r) for r in range(world)]
    assert max(sizes) - min(sizes) <= 1 and max(sizes) == -(-n // world)
    # every rank's buffer holds the footprint of each of its windows
    for r in range(world):
        for w in (part.ranges[r][0], part.ranges[r][1] - 1):
            sd = grid.window_start(w)[0]
            assert part.buf_lo[r] <= sd and sd + grid.roi[0] <= part.buf_hi[r]
    # the owned planes tile the volume; whoever has a window over an owned plane is among the contributors
    assert part.own_lo[0] == 0 and part.own_hi[-1] == shape[0]
    assert all(part.own_hi[r] == part.own_lo[r + 1] and part.own_hi[r] > part.own_lo[r] for r in range(world - 1))
    owner_of = np.zeros(n, dtype=np.int64)
    for r, (a, b) in enumerate(part.ranges):
        owner_of[a:b] = r
    for r in range(world):
        need = set()
        for d in (part.own_lo[r], part.own_hi[r] - 1, (part.own_lo[r] + part.own_hi[r]) // 2):
            need |= {int(owner_of[w]) for w in _covering(grid, d)}
        assert need <= set(part.contributors(r)), (r, need, part.contributors(r))
    if shape == (512, 512, 1024) and world == 8:
        assert max(sizes) == 263  # BASELINE.json configs[2]: 2100 windows, 7.98x ceiling


def test_flat_partition_rejects_too_many_ranks():
    grid = make_grid((16, 16, 16), 16, 0.0)
    with pytest.raises(ValueError):
        flat.flat_partition(grid, 2)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _gather_worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        grid = make_grid((37, 20, 24), 16, 0.5)
        part = flat.flat_partition(grid, world)
        full = (torch.arange(37 * 20 * 24) % 251).to(torch.uint8).view(1, 37, 20, 24)
        own = full[:, part.own_lo[rank]:part.own_hi[rank]].contiguous()
        got = flat.gather_slabs(own, part, rank)
        out[rank] = bool(torch.equal(got, full))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_gather_slabs_one_collective(world):
    port = _free_port()
    with mp.Manager() as mgr:
        out = mgr.dict()
        mp.spawn(_gather_worker, args=(world, port, out), nprocs=world, join=True)
        assert dict(out) == {r: True for r in range(world)}
