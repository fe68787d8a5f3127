"""Host-side window geometry: the argument handling of engine/utils.py:81-110 without MONAI.

Mirrors the reference's observable behaviour (same error types and messages, same roi fall-back,
same clamped last window, same C-order enumeration); the arithmetic of the per-axis starts and the
per-coordinate cover table runs in the C library's host-only helpers (``mss_axis_starts``,
``mss_geom_table_build``) so kernels and host agree by construction.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field
from typing import Any, List, Optional, Sequence, Tuple

import numpy as np

from . import _lib

PAD_MODES = ("constant", "reflect", "replicate", "circular")  # PytorchPadMode, engine/utils.py:28
BLEND_MODES = ("constant", "gaussian")  # BlendMode, engine/utils.py:26


def _option(value: Any, supported: Sequence[str], what: str) -> str:
    """``look_up_option``: accept the string value or an Enum member carrying it (engine/utils.py:103)."""
    v = getattr(value, "value", value)
    if isinstance(v, str):
        v = v.strip()
    if v not in supported:
        raise ValueError(f"Unsupported option '{value}', Available options are {list(supported)}.")
    return v


def fall_back_roi(roi_size: Any, image_size: Sequence[int]) -> Tuple[int, ...]:
    """``fall_back_tuple(roi_size, image_size_)`` (engine/utils.py:95): scalars repeat, non-positive / None
    components fall back to the image dimension."""
    n = len(image_size)
    if isinstance(roi_size, (list, tuple, np.ndarray)):
        if len(roi_size) != n:
            raise ValueError(f"Sequence must have length {n}, got {len(roi_size)}.")
        user = tuple(roi_size)
    else:
        user = (roi_size,) * n
    return tuple(int(u) if (u is not None and u and u > 0) else int(d) for u, d in zip(user, image_size))


def scan_interval(image_size: Sequence[int], roi: Sequence[int], overlap: float) -> Tuple[int, ...]:
    """``_get_scan_interval`` (engine/utils.py:105)."""
    if len(image_size) != 3:
        raise ValueError("image coord different from spatial dims.")
    if len(roi) != 3:
        raise ValueError("roi coord different from spatial dims.")
    out = []
    for i in range(3):
        if roi[i] == image_size[i]:
            out.append(int(roi[i]))
        else:
            step = int(roi[i] * (1 - overlap))
            out.append(step if step > 0 else 1)
    return tuple(out)


def axis_starts(image: int, roi: int, interval: int) -> List[int]:
    lib = _lib.load()
    n = lib.mss_axis_starts(image, roi, interval, None, 0)
    if n <= 0:
        _lib.check(n, "mss_axis_starts")
    buf = (C.c_int32 * n)()
    n2 = lib.mss_axis_starts(image, roi, interval, buf, n)
    if n2 != n:
        _lib.check(n2 if n2 < 0 else -1, "mss_axis_starts")
    return list(buf)


@dataclass
class WindowGrid:
    """Everything engine/utils.py:81-110 derives from (input shape, roi_size, overlap)."""

    orig_size: Tuple[int, int, int]    # spatial size of the caller's volume
    roi: Tuple[int, int, int]
    image_size: Tuple[int, int, int]   # stitched size: max(orig, roi) per axis (engine/utils.py:97)
    pad_lo: Tuple[int, int, int]       # symmetric pad offsets, diff // 2 (engine/utils.py:98-102)
    interval: Tuple[int, int, int]
    starts: Tuple[Tuple[int, ...], ...]
    table: np.ndarray = field(repr=False, default=None)  # int32 geometry table (host copy)

    @property
    def n_starts(self) -> Tuple[int, int, int]:
        return tuple(len(s) for s in self.starts)

    @property
    def n_windows(self) -> int:
        a, b, c = self.n_starts
        return a * b * c

    @property
    def padded(self) -> bool:
        return self.image_size != self.orig_size

    def window_start(self, n: int) -> Tuple[int, int, int]:
        """Start of window ``n`` in C order (first axis slowest), as dense_patch_slices enumerates."""
        nd, nh, nw = self.n_starts
        return self.starts[0][n // (nh * nw)], self.starts[1][(n // nw) % nh], self.starts[2][n % nw]

    def centers(self, n: int) -> Tuple[float, float, float]:
        """engine/utils.py:126-128."""
        s = self.window_start(n)
        return tuple((s[a] + self.roi[a] - self.roi[a] // 2) / self.image_size[a] for a in range(3))


def build_table(image_size: Sequence[int], roi: Sequence[int], starts: Sequence[Sequence[int]]) -> np.ndarray:
    lib = _lib.load()
    img = _lib.I3(*image_size)
    r = _lib.I3(*roi)
    ns = _lib.I3(*(len(s) for s in starts))
    length = lib.mss_geom_table_len(img, ns)
    if length <= 0:
        raise _lib.MssError("mss_geom_table_len rejected the geometry")
    table = np.zeros(int(length), dtype=np.int32)
    arrs = [np.ascontiguousarray(s, dtype=np.int32) for s in starts]
    rc = lib.mss_geom_table_build(img, r, ns, arrs[0].ctypes.data, arrs[1].ctypes.data, arrs[2].ctypes.data,
                                  table.ctypes.data, length)
    _lib.check(rc, "mss_geom_table_build")
    return table


def make_grid(spatial_size: Sequence[int], roi_size: Any, overlap: float,
              starts: Optional[Sequence[Sequence[int]]] = None) -> WindowGrid:
    """Window grid of one volume.  ``starts`` overrides the per-axis starts (used by the slab partitioner)."""
    if len(spatial_size) != 3:
        raise ValueError("medicalsemseg_b200 stitches 3-D volumes (N, C, D, H, W) only.")
    if overlap < 0 or overlap >= 1:
        raise AssertionError("overlap must be >= 0 and < 1.")  # engine/utils.py:82-83
    orig = tuple(int(s) for s in spatial_size)
    roi = fall_back_roi(roi_size, orig)
    image = tuple(max(orig[i], roi[i]) for i in range(3))
    pad_lo = tuple(max(roi[i] - orig[i], 0) // 2 for i in range(3))
    interval = scan_interval(image, roi, overlap)
    if starts is None:
        starts = [axis_starts(image[a], roi[a], interval[a]) for a in range(3)]
    starts_t = tuple(tuple(int(x) for x in s) for s in starts)
    return WindowGrid(orig, roi, image, pad_lo, interval, starts_t, build_table(image, roi, starts_t))
