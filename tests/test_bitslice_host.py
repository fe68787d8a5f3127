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
