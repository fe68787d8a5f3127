"""The bit-sliced vote / Dice chunk logic (medicalsemseg_b200/csrc/bitslice.cuh) compiled for the HOST and checked
against the oracle: the kernels inline exactly these functions, so this is CPU coverage of the device arithmetic."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from oracle import dice as odice
from oracle import vote as ovote

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "csrc", "host_chunks.cpp")
OUT = os.path.join(HERE, "_build", "libhost_chunks.so")


@pytest.fixture(scope="module")
def lib():
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    hdr = os.path.join(HERE, "..", "medicalsemseg_b200", "csrc", "bitslice.cuh")
    if not os.path.exists(OUT) or os.path.getmtime(OUT) < max(os.path.getmtime(SRC), os.path.getmtime(hdr)):
        subprocess.run(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-x", "c++", SRC, "-o", OUT], check=True)
    return C.CDLL(OUT)


def test_bitplanes_roundtrip(lib):
    rs = np.random.RandomState(0)
    for _ in range(50):
        v = rs.randint(0, 16, 32).astype(np.uint8)
        out = np.zeros(32, np.uint8)
        planes = np.zeros(4, np.uint32)
        lib.host_bitplanes_roundtrip(v.ctypes.data, out.ctypes.data, planes.ctypes.data)
        assert np.array_equal(out, v)
        for j in range(4):
            assert bin(int(planes[j])).count("1") == int(((v >> j) & 1).sum())


@pytest.mark.parametrize("m", [1, 2, 3, 4, 5, 6, 7, 8])
@pytest.mark.parametrize("k", [2, 3, 5, 14, 16])
def test_vote_chunks_match_oracle(lib, m, k):
    rs = np.random.RandomState(100 * m + k)
    n = 32 * 257
    base = rs.randint(0, k, n).astype(np.uint8)
    maps = []
    for i in range(m):
        # correlated maps (real ensembles agree mostly) with labels up to 15, some of them >= k
        noise = rs.randint(0, 16, n).astype(np.uint8)
        maps.append(np.ascontiguousarray(np.where(rs.random_sample(n) < 0.35, noise, base)))
    ptrs = (C.c_void_p * m)(*[a.ctypes.data for a in maps])
    out = np.zeros(n, np.uint8)
    assert lib.host_vote_chunks(ptrs, m, k, C.c_longlong(n), out.ctypes.data) == 0
    want = ovote.majority_vote(maps, k)
    assert np.array_equal(out, want)


@pytest.mark.parametrize("k", [1, 2, 3, 7, 14, 15, 16])
def test_dice_chunks_match_oracle(lib, k):
    rs = np.random.RandomState(k)
    n = 32 * 4099  # > 2040 chunks: exercises the 16-bit counter flush
    pred = rs.randint(0, 16, n).astype(np.uint8)
    lab = np.where(rs.random_sample(n) < 0.6, pred, rs.randint(0, 16, n)).astype(np.uint8)
    counts = np.zeros((3, 16), np.int64)
    assert lib.host_dice_chunks(pred.ctypes.data, lab.ctypes.data, k, C.c_longlong(n), counts.ctypes.data) == 0
    want = odice.dice_counts(pred, lab, k)
    assert np.array_equal(counts[:, :k], want)


def test_dice_chunks_uniform_volume(lib):
    # one class everywhere: the worst case for the packed 16-bit counters
    n = 32 * 5000
    pred = np.full(n, 3, np.uint8)
    counts = np.zeros((3, 16), np.int64)
    assert lib.host_dice_chunks(pred.ctypes.data, pred.ctypes.data, 14, C.c_longlong(n), counts.ctypes.data) == 0
    assert counts[0, 3] == n and counts[1, 3] == n and counts[2, 3] == n and counts.sum() == 3 * n


@pytest.mark.parametrize("shape,density", [((9, 11, 13), 0.05), ((20, 7, 15), 0.01), ((1, 16, 16), 0.1), ((12, 12, 1), 0.3),
                                           ((17, 19, 23), 0.002), ((8, 8, 8), 1.0), ((6, 5, 4), 0.0), ((30, 3, 31), 0.03),
                                           ((5, 6, 70), 0.02), ((40, 33, 37), 0.0005)])
def test_edt_line_routine_matches_scipy(lib, shape, density):
    """csrc/edt.cuh (the routine the GPU pass kernel runs per line) applied along the three axes on the host: exact
    squared distances, i.e. scipy.ndimage.distance_transform_edt squared."""
    from scipy import ndimage
    rs = np.random.RandomState(int(np.prod(shape)) + int(density * 1000))
    feat = (rs.random_sample(shape) < density).astype(np.uint8)
    out = np.zeros(shape, np.int32)
    lib.host_edt_squared(feat.ctypes.data, shape[0], shape[1], shape[2], out.ctypes.data)
    # round 2: row scan from the mask + the envelope with the register-cached stack top (what the GPU driver runs now)
    out2 = np.zeros(shape, np.int32)
    lib.host_edt_squared_v2(feat.ctypes.data, shape[0], shape[1], shape[2], 0, out2.ctypes.data)
    assert np.array_equal(out, out2)
    lib.host_edt_squared_v2(feat.ctypes.data, shape[0], shape[1], shape[2], 1, out2.ctypes.data)  # divisions by reciprocal table
    assert np.array_equal(out, out2)
    if feat.sum() == 0:
        assert np.all(out == 1 << 29)
        return
    want = ndimage.distance_transform_edt(feat == 0)
    assert np.array_equal(np.sqrt(out.astype(np.float64)), want)
    assert np.array_equal(out, np.rint(want ** 2).astype(np.int32))


def test_edt_reciprocal_division_is_exact(lib):
    """edt_div2k(a, k, ceil(2^31 / k)) == a // (2 k) over a strided sweep of a < 2^31 (plus both ends) for every k < 2048."""
    lib.host_edt_div_check.restype = C.c_longlong
    assert lib.host_edt_div_check(2048, 99991) == 0


def _blobby_labels(rs, shape, k):
    from scipy import ndimage
    f = ndimage.gaussian_filter(rs.standard_normal(shape), 2.5)
    return np.digitize(f, np.quantile(f, np.linspace(0, 1, k + 1)[1:-1])).astype(np.uint8)


@pytest.mark.parametrize("shape", [(20, 24, 18), (12, 30, 9), (1, 20, 20), (16, 1, 12), (10, 11, 1)])
def test_mask_edges_rule_matches_monai_restated(lib, shape):
    """The surface rule of csrc/edt.cuh::mask_edge_at vs oracle.hausdorff.get_mask_edges (binary_erosion XOR mask on
    the bounding box, squeezed) for every class of blobby random label maps, thin boxes included."""
    from oracle import hausdorff as oh
    rs = np.random.RandomState(sum(shape))
    gt = _blobby_labels(rs, shape, 4)
    pred = gt.copy()
    flip = rs.random_sample(shape) < 0.08
    pred[flip] = rs.randint(0, 4, int(flip.sum()))
    # a class confined to a single plane / line / voxel of the volume: the reference squeezes those axes away
    pred[pred == 3] = 2
    gt[gt == 3] = 2
    pred[shape[0] // 2, 2:6, 0:min(5, shape[2])] = 3
    gt[shape[0] // 2, 3:7, 0:min(4, shape[2])] = 3
    for c in range(4):
        union = (pred == c) | (gt == c)
        if not union.any():
            continue
        lo, hi = [], []
        for a in range(3):
            idx = np.nonzero(union.any(axis=tuple(x for x in range(3) if x != a)))[0]
            lo.append(int(idx[0])), hi.append(int(idx[-1]) + 1)
        n = [h - l for l, h in zip(lo, hi)]
        if all(v == 1 for v in n):
            continue  # 0-d erosion: not defined by the reference
        want_p, want_g = oh.get_mask_edges(pred == c, gt == c)
        for arr, want in ((pred, want_p), (gt, want_g)):
            got = np.zeros(n, np.uint8)
            lib.host_mask_edges(np.ascontiguousarray(arr).ctypes.data, (C.c_int * 3)(*shape), c, (C.c_int * 3)(*lo),
                                (C.c_int * 3)(*hi), got.ctypes.data)
            assert np.array_equal(np.squeeze(got).astype(bool), want), (c, lo, hi)
