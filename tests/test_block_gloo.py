"""Host-side logic of the 3-D block partition on CPU: gloo process groups (4 and 8 ranks) reduce halos axis by axis and
must reproduce the single-process weighted sums on every rank's owned box (the CUDA add kernel is replaced by a torch
stand-in; partition choice, ownership, halo boxes and the multi-hop schedule are what is tested)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from medicalsemseg_b200 import block
from medicalsemseg_b200.grid import make_grid


def test_choose_dims_known_answers():
    g = make_grid((512, 512, 1024), 96, 0.5)  # BASELINE.json configs[2]: 10 x 10 x 21 window starts
    assert block.choose_dims(g, 1) == (1, 1, 1)
    assert block.choose_dims(g, 2) == (2, 1, 1)      # 5 * 10 * 21 = 1050 = exactly half
    assert block.choose_dims(g, 4) == (2, 2, 1)      # 525 (a 1-D slab split would carry 600)
    assert block.choose_dims(g, 8) == (2, 2, 2)      # 275 (1-D: 300)
    p = block.block_partition(g, 8)
    loads = [p.n_windows(r) for r in range(8)]
    assert max(loads) == 275 and sum(loads) == g.n_windows
    # owned boxes tile the volume exactly once
    cover = np.zeros((512 // 16, 512 // 16, 1024 // 16), np.int32)
    for r in range(8):
        lo, hi = p.box(r, "own")
        assert all(v % 16 == 0 for v in lo + hi)
        cover[lo[0] // 16:hi[0] // 16, lo[1] // 16:hi[1] // 16, lo[2] // 16:hi[2] // 16] += 1
        blo, bhi = p.box(r, "buf")
        assert all(blo[a] <= lo[a] or p.coords(r)[a] == 0 for a in range(3))
    assert (cover == 1).all()
    with pytest.raises(ValueError):
        block.block_partition(make_grid((128, 128, 128), 96, 0.25), 16)
    with pytest.raises(ValueError):
        block.block_partition(g, 8, dims=(2, 2, 3))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _windows_sum(g, k, seed, win_lo, win_hi, buf_lo, buf_hi):
    """Partial sums of the windows of the index box [win_lo, win_hi) on the buffer box (unit weights)."""
    rs = np.random.RandomState(seed)
    ext = [h - l for l, h in zip(buf_lo, buf_hi)]
    acc = np.zeros([1, k] + ext, dtype=np.float32)
    nd, nh, nw = g.n_starts
    for n in range(g.n_windows):
        idx = (n // (nh * nw), (n // nw) % nh, n % nw)
        logits = rs.standard_normal((k,) + g.roi).astype(np.float32)  # drawn for every window: same stream on all ranks
        if not all(win_lo[a] <= idx[a] < win_hi[a] for a in range(3)):
            continue
        s = [g.window_start(n)[a] - buf_lo[a] for a in range(3)]
        acc[0, :, s[0]:s[0] + g.roi[0], s[1]:s[1] + g.roi[1], s[2]:s[2] + g.roi[2]] += logits
    return acc


def _worker(rank, world, port, shape, roi, overlap, dims, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.set_num_threads(1)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g = make_grid(shape, roi, overlap)
        p = block.block_partition(g, world, dims)
        wl, wh = p.win_box(rank)
        bl, bh = p.box(rank, "buf")
        acc = torch.from_numpy(_windows_sum(g, 2, 7, wl, wh, bl, bh))
        block.exchange_halos(acc, p, rank, None, add_fn=lambda dst, src: dst.add_(src))
        ol, oh = p.box(rank, "own")
        own = block._box_view(acc, [ol[a] - bl[a] for a in range(3)], [oh[a] - bl[a] for a in range(3)])
        full = _windows_sum(g, 2, 7, (0, 0, 0), g.n_starts, (0, 0, 0), g.image_size)
        want = block._box_view(torch.from_numpy(full), ol, oh)
        out[rank] = bool(torch.allclose(own, want, rtol=1e-5, atol=1e-5))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,shape,roi,overlap,dims", [
    (4, (40, 40, 24), 16, 0.5, (2, 2, 1)),
    (4, (24, 40, 56), 16, 0.5, None),           # chooses its own factorisation
    (8, (40, 40, 40), 16, 0.5, (2, 2, 2)),      # corners travel three hops
    (4, (16, 48, 40), 16, 0.75, (1, 2, 2)),     # high overlap: forwarding chains inside an axis step
])
def test_block_halo_exchange_reproduces_single_process_sums(world, shape, roi, overlap, dims):
    port = _free_port()
    with mp.Manager() as mgr:
        out = mgr.dict()
        mp.spawn(_worker, args=(world, port, shape, roi, overlap, dims, out), nprocs=world, join=True)
        assert dict(out) == {r: True for r in range(world)}


def test_partition_invariants_on_random_grids():
    """Every window belongs to exactly one rank, owned boxes tile the volume, a rank's buffer holds all its windows, and
    its halo along an axis is exactly what it wrote beyond its ownership - for random shapes, rois, overlaps and worlds."""
    rs = np.random.RandomState(0)
    done = 0
    while done < 60:
        shape = tuple(int(v) for v in rs.randint(20, 90, 3))
        roi = tuple(int(v) for v in rs.randint(8, 20, 3))
        overlap = float(rs.choice([0.0, 0.25, 0.5, 0.6, 0.75]))
        world = int(rs.choice([1, 2, 3, 4, 6, 8]))
        g = make_grid(shape, roi, overlap)
        try:
            p = block.block_partition(g, world)
        except ValueError:
            continue
        done += 1
        owner = np.full(g.n_starts, -1)
        cover = np.zeros(shape, np.int16)
        for r in range(world):
            wl, wh = p.win_box(r)
            assert np.all(owner[wl[0]:wh[0], wl[1]:wh[1], wl[2]:wh[2]] == -1)
            owner[wl[0]:wh[0], wl[1]:wh[1], wl[2]:wh[2]] = r
            bl, bh = p.box(r, "buf")
            ol, oh = p.box(r, "own")
            cover[ol[0]:oh[0], ol[1]:oh[1], ol[2]:oh[2]] += 1
            for a in range(3):
                starts = g.starts[a][wl[a]:wh[a]]
                assert bl[a] == starts[0] and bh[a] == starts[-1] + g.roi[a]          # buffer = union of its windows
                assert bl[a] <= ol[a] or p.coords(r)[a] == 0
                c = p.coords(r)[a]
                halo = p.axes[a].halo(c)
                assert halo == ((oh[a], bh[a]) if c + 1 < p.dims[a] else (0, 0))
        assert np.all(owner >= 0) and np.all(cover == 1)
        assert max(p.n_windows(r) for r in range(world)) * world >= g.n_windows
