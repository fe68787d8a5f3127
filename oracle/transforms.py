"""CPU oracle (TEST INFRASTRUCTURE) for the test-time intensity transforms (data/dataset_builder.py:322-370).

``scale_cubed_intensity_range`` restates the reference's own ``ScaleCubedIntensityRange`` (data/transforms.py:17-71)
in float32 NumPy; ``scale_intensity_range`` / ``normalize_intensity`` restate MONAI 0.8's ``ScaleIntensityRange`` and
``NormalizeIntensity`` (third-party, absent from /root/reference, unpinned in requirements.txt: parity unpinned for
those two beyond their published formulas).  The cubed transform is pinned against outputs of the reference class
itself (tests/golden/intensity_*.npz, produced by tests/golden/make_golden.py).

Precision note: the reference subtracts ``np.cbrt(a_min)`` - a NumPy float64 scalar - from a float32 array.  NumPy < 2
(the reference's vintage) keeps float32 there, NumPy >= 2 promotes to float64 and rounds once at the final cast; the
two differ by at most 1 ulp of the result.  ``dtype=`` selects the intermediate precision.
"""
from __future__ import annotations

from typing import Optional

import numpy as np


def scale_intensity_range(img: np.ndarray, a_min: float, a_max: float, b_min: Optional[float] = None,
                          b_max: Optional[float] = None, clip: bool = False, dtype=np.float32) -> np.ndarray:
    """MONAI 0.8 ScaleIntensityRange.__call__ == data/transforms.py:56-69 without the cube root."""
    img = img.astype(dtype)
    a_min, a_max = float(a_min), float(a_max)
    if a_max - a_min == 0.0:
        if b_min is None:
            return (img - dtype(a_min)).astype(np.float32)
        return (img - dtype(a_min) + dtype(b_min)).astype(np.float32)
    img = (img - dtype(a_min)) / dtype(a_max - a_min)
    if b_min is not None and b_max is not None:
        img = img * dtype(dtype(b_max) - dtype(b_min)) + dtype(b_min)
    if clip:
        img = np.clip(img, None if b_min is None else dtype(b_min), None if b_max is None else dtype(b_max))
    return img.astype(np.float32)


def scale_cubed_intensity_range(img: np.ndarray, a_min: float, a_max: float, b_min: Optional[float] = None,
                                b_max: Optional[float] = None, clip: bool = False, dtype=np.float32) -> np.ndarray:
    """data/transforms.py:45-46 (bounds) and :54 (data) take the cube root, then the range scaling above."""
    return scale_intensity_range(np.cbrt(img.astype(np.float32)), float(np.cbrt(a_min)), float(np.cbrt(a_max)), b_min, b_max,
                                 clip, dtype)


def normalize_intensity(img: np.ndarray, subtrahend: Optional[float] = None, divisor: Optional[float] = None,
                        nonzero: bool = False, channel_wise: bool = False) -> np.ndarray:
    """MONAI 0.8 NormalizeIntensity (``_normalize``): float32, population std, zero divisor -> 1."""
    img = img.astype(np.float32).copy()
    if channel_wise:
        for c in range(img.shape[0]):
            img[c] = normalize_intensity(img[c], subtrahend, divisor, nonzero, False)
        return img
    slices = (img != 0) if nonzero else np.ones(img.shape, dtype=bool)
    if not slices.any():
        return img
    sub = np.float32(subtrahend) if subtrahend is not None else np.float32(np.mean(img[slices]))
    div = np.float32(divisor) if divisor is not None else np.float32(np.std(img[slices]))
    if div == 0.0:
        div = np.float32(1.0)
    img[slices] = (img[slices] - sub) / div
    return img


def scale_intensity_range_percentiles(img: np.ndarray, lower: float, upper: float, b_min: Optional[float],
                                      b_max: Optional[float], clip: bool = False, relative: bool = False) -> np.ndarray:
    """MONAI 0.8 ScaleIntensityRangePercentiles._normalize."""
    a_min = float(np.percentile(img, lower))
    a_max = float(np.percentile(img, upper))
    bmin, bmax = b_min, b_max
    if relative:
        bmin = ((b_max - b_min) * (lower / 100.0)) + b_min
        bmax = ((b_max - b_min) * (upper / 100.0)) + b_min
    out = scale_intensity_range(img, a_min, a_max, bmin, bmax, clip=False)
    if clip:
        out = np.clip(out, b_min, b_max)
    return out.astype(np.float32)
