"""CPU oracle (TEST INFRASTRUCTURE) for the label-map resampling after the argmax.

Restates utils/misc.py:420-425 (``resample_3d``): the reference calls ``scipy.ndimage.zoom(img, ratio, order=0,
prefilter=False)`` with ``ratio = float(t) / float(n)`` per axis.  scipy is a third-party dependency of the
reference (``requirements.txt``: ``scipy``, unpinned); it is installed in this image (1.18.1), so ``resample_3d`` below
calls it exactly like the reference does, and ``zoom_index_rule`` restates the per-axis index rule it applies:
``k -> floor(k * zoom + 0.5)`` with ``zoom = (n_in - 1) / (n_out - 1)`` in float64 and the constant 0 wherever
``k * zoom`` leaves ``[0, n_in - 1]`` (a rounding artefact that zeroes the last plane for some size pairs - kept,
the drop-in must be bit-exact).  Pinned against outputs of the reference's own function in
tests/golden/resample_*.npz (tests/golden/make_golden.py AST-extracts and runs it).
"""
from __future__ import annotations

from typing import Sequence

import numpy as np


def resample_3d(img: np.ndarray, target_size: Sequence[int]) -> np.ndarray:
    """utils/misc.py:420-425, statement for statement."""
    from scipy import ndimage

    imx, imy, imz = img.shape
    tx, ty, tz = target_size
    zoom_ratio = (float(tx) / float(imx), float(ty) / float(imy), float(tz) / float(imz))
    return ndimage.zoom(img, zoom_ratio, order=0, prefilter=False)


def zoom_index_rule(n_in: int, n_out: int) -> np.ndarray:
    """Per-axis source index of scipy's order-0 zoom (``-1``: scipy writes the constant 0)."""
    zoom = np.float64(n_in - 1) / np.float64(n_out - 1) if n_out > 1 else np.float64(1.0)
    cc = np.arange(n_out, dtype=np.float64) * zoom
    idx = np.minimum(np.floor(cc + 0.5).astype(np.int64), n_in - 1)
    return np.where((cc >= 0) & (cc <= n_in - 1), idx, -1).astype(np.int32)


def resample_3d_rule(img: np.ndarray, target_size: Sequence[int]) -> np.ndarray:
    """The same result from the index rule alone (no scipy): what the CUDA kernel computes."""
    out_shape = tuple(int(round(n * (float(t) / float(n)))) for n, t in zip(img.shape, target_size))
    tabs = [zoom_index_rule(n, o) for n, o in zip(img.shape, out_shape)]
    out = img[np.ix_(*[np.maximum(t, 0) for t in tabs])].copy()
    for a, t in enumerate(tabs):
        sl = [slice(None)] * 3
        sl[a] = t < 0
        out[tuple(sl)] = 0
    return out
