"""Test-time intensity transforms on the GPU, one fused elementwise pass over the volume: the chain
data/dataset_builder.py:322-370 builds for the test loader (ScaleIntensityRange, the reference's own
ScaleCubedIntensityRange data/transforms.py:17-71, ScaleIntensityRangePercentiles, NormalizeIntensity) - the step
right before sliding_window_inference.  Statistics (percentiles, non-zero mean/std) are torch reductions - plumbing;
the per-voxel arithmetic is ``mss_intensity_transform``."""
from __future__ import annotations

from typing import Any, Optional

import numpy as np
import torch

from . import _lib


def _launch(img: torch.Tensor, flags: int, a_min: float = 0.0, denom: float = 1.0, b_min: float = 0.0, b_max: float = 0.0,
            sub: float = 0.0, div: float = 1.0, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    if not img.is_cuda:
        raise _lib.MssError("intensity transforms need a CUDA tensor; there is no CPU fallback")
    src = img.to(torch.float32).contiguous()
    if out is None:
        out = torch.empty_like(src)
    elif out.dtype != torch.float32 or not out.is_contiguous() or out.numel() != src.numel():
        raise ValueError("out must be a contiguous float32 tensor of the input's size")
    with torch.cuda.device(src.device):
        rc = _lib.load().mss_intensity_transform(src.data_ptr(), out.data_ptr(), src.numel(), int(flags), float(a_min),
                                                 float(denom), float(b_min), float(b_max), float(sub), float(div),
                                                 torch.cuda.current_stream().cuda_stream)
    _lib.check(rc, "mss_intensity_transform")
    return out


def _scale_flags(a_min: float, a_max: float, b_min: Optional[float], b_max: Optional[float], clip: bool):
    """Flags and constants of MONAI ScaleIntensityRange.__call__ (same statements as data/transforms.py:56-67)."""
    a_min, a_max = float(a_min), float(a_max)
    if a_max - a_min == 0.0:
        # degenerate range: `img - a_min (+ b_min)`; x - (-b_min) is exactly x + b_min
        flags = _lib.INT_SCALE | (_lib.INT_NORM if b_min is not None else 0)
        return flags, a_min, 1.0, 0.0, 0.0, (-float(b_min) if b_min is not None else 0.0)
    flags = _lib.INT_SCALE
    if b_min is not None and b_max is not None:
        flags |= _lib.INT_RESCALE
    if clip:
        flags |= (_lib.INT_CLIP_LO if b_min is not None else 0) | (_lib.INT_CLIP_HI if b_max is not None else 0)
    return flags, a_min, a_max - a_min, (0.0 if b_min is None else float(b_min)), (0.0 if b_max is None else float(b_max)), 0.0


def _cubed_f64() -> bool:
    """data/transforms.py:62 subtracts the NumPy float64 scalar ``np.cbrt(a_min)`` from a float32 array: NumPy >= 2
    evaluates that in float64 (one rounding at the final cast), NumPy < 2 in float32.  Follow the installed NumPy, i.e.
    what the reference itself would produce in this environment."""
    return int(np.__version__.split(".")[0]) >= 2


def scale_intensity_range(img: torch.Tensor, a_min: float, a_max: float, b_min: Optional[float] = None,
                          b_max: Optional[float] = None, clip: bool = False, cubed: bool = False,
                          out: Optional[torch.Tensor] = None, float64: Optional[bool] = None) -> torch.Tensor:
    """``monai.transforms.ScaleIntensityRange`` (``cubed=False``) or the reference's ``ScaleCubedIntensityRange``
    (``cubed=True``: cube root of the data and of the bounds first, data/transforms.py:45-46,54).  ``float64`` selects
    the intermediate precision (default: float32, and for the cubed scaler whatever the installed NumPy does)."""
    if cubed:
        a_min, a_max = float(np.cbrt(a_min)), float(np.cbrt(a_max))
    flags, lo, denom, bmin, bmax, sub = _scale_flags(a_min, a_max, b_min, b_max, clip)
    if cubed:
        flags |= _lib.INT_CBRT
    if float64 if float64 is not None else (cubed and _cubed_f64()):
        flags |= _lib.INT_F64
    return _launch(img, flags, lo, denom, bmin, bmax, sub, 1.0, out)


def normalize_intensity(img: torch.Tensor, subtrahend: Optional[float] = None, divisor: Optional[float] = None,
                        nonzero: bool = False, channel_wise: bool = False, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """``monai.transforms.NormalizeIntensity``: ``(img - subtrahend) / divisor`` with the mean / population std of the
    (non-zero) voxels when not given, per channel of a channel-first ``[C, ...]`` image when ``channel_wise``."""
    src = img.to(torch.float32).contiguous()
    if out is None:
        out = torch.empty_like(src)
    if channel_wise:
        for c in range(src.shape[0]):
            normalize_intensity(src[c], subtrahend, divisor, nonzero, False, out=out[c])
        return out
    sel = src[src != 0] if nonzero else src.reshape(-1)
    if sel.numel() == 0:
        out.copy_(src)
        return out
    sub = float(subtrahend) if subtrahend is not None else float(sel.mean())
    div = float(divisor) if divisor is not None else float(sel.std(unbiased=False))
    if div == 0.0:
        div = 1.0
    flags = _lib.INT_NORM | (_lib.INT_NONZERO if nonzero else 0)
    return _launch(src, flags, sub=sub, div=div, out=out)


def percentile(img: torch.Tensor, q: float) -> float:
    """``numpy.percentile(img, q)`` (linear interpolation between order statistics) of a CUDA tensor: the two order
    statistics come from ``mss_select2`` (three radix-histogram passes over the float32 voxels, no sort)."""
    lib = _lib.load()
    flat = img.reshape(-1).to(torch.float32).contiguous()
    n = flat.numel()
    pos = (n - 1) * (float(q) / 100.0)
    lo = int(np.floor(pos))
    hi = min(lo + 1, n - 1)
    t = pos - lo
    with torch.cuda.device(flat.device):
        out = torch.zeros(4, dtype=torch.int64, device=flat.device)
        scratch = torch.empty(int(lib.mss_select_scratch_bytes()) // 8 + 1, dtype=torch.int64, device=flat.device)
        rc = lib.mss_select2(flat.data_ptr(), 1, None, n, 0, 0.0, lo, hi, scratch.data_ptr(), out.data_ptr(),
                             torch.cuda.current_stream().cuda_stream)
        _lib.check(rc, "mss_select2")
        keys = out.cpu().numpy()
    def unmap(k: int) -> float:  # inverse of the order-preserving float32 -> uint32 map
        u = np.uint32(k)
        u = np.uint32(u & np.uint32(0x7FFFFFFF)) if (u & np.uint32(0x80000000)) else np.uint32(~u)
        return float(np.frombuffer(np.uint32(u).tobytes(), dtype=np.float32)[0])
    a, b = unmap(int(keys[1])), unmap(int(keys[2]))
    d = b - a
    return float(b - d * (1.0 - t)) if t >= 0.5 else float(a + d * t)  # numpy's _lerp


def scale_intensity_range_percentiles(img: torch.Tensor, lower: float, upper: float, b_min: Optional[float],
                                      b_max: Optional[float], clip: bool = False, relative: bool = False,
                                      out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """``monai.transforms.ScaleIntensityRangePercentiles`` as configured at data/dataset_builder.py:344-353."""
    a_min, a_max = percentile(img, lower), percentile(img, upper)
    bmin, bmax = b_min, b_max
    if relative:
        bmin = ((b_max - b_min) * (lower / 100.0)) + b_min
        bmax = ((b_max - b_min) * (upper / 100.0)) + b_min
    res = scale_intensity_range(img, a_min, a_max, bmin, bmax, clip=False, out=out)
    if clip:
        res.clamp_(b_min, b_max)
    return res


def test_time_intensity(img: torch.Tensor, cfg: Any, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """The intensity part of ``build_test_transforms(cfg)`` (data/dataset_builder.py:322-370) for one channel-first
    image ``[C, D, H, W]``: range scaling (cubed / fixed / percentile) then normalisation.  With fixed statistics the
    whole chain is ONE launch (8 bytes per voxel instead of 8 per transform)."""
    g = lambda name, default=None: getattr(cfg, name, default)  # noqa: E731
    flags, a_min, denom, bmin, bmax = 0, 0.0, 1.0, 0.0, 0.0
    if g("t_cubed_ct_intensity", False):
        lo, hi = float(np.cbrt(g("t_ct_min"))), float(np.cbrt(g("t_ct_max")))
        flags, a_min, denom, bmin, bmax, _ = _scale_flags(lo, hi, 0.0, 1.0, True)
        flags |= _lib.INT_CBRT
        if g("t_normalize", False) and not g("t_normalize_channel_wise", False) and _cubed_f64():
            # the float64 scaler result is cast to float32 before NormalizeIntensity sees it: two launches keep that rounding
            img = _launch(img, flags | _lib.INT_F64, a_min, denom, bmin, bmax)
            flags = 0
        elif _cubed_f64():
            flags |= _lib.INT_F64
    elif g("t_fixed_ct_intensity", False):
        flags, a_min, denom, bmin, bmax, _ = _scale_flags(float(g("t_ct_min")), float(g("t_ct_max")), 0.0, 1.0, True)
    elif g("t_percentile_ct_intensity", False):
        img = scale_intensity_range_percentiles(img, 5, 95, 0.0, 1.0, clip=True, relative=False)
    if g("t_normalize", False):
        if g("t_normalize_channel_wise", False):
            if flags:
                img = _launch(img, flags, a_min, denom, bmin, bmax)
            return normalize_intensity(img, nonzero=True, channel_wise=True, out=out)
        div = float(g("t_norm_std"))
        return _launch(img, flags | _lib.INT_NORM, a_min, denom, bmin, bmax, float(g("t_norm_mean")), div if div != 0.0 else 1.0,
                       out)
    if flags:
        return _launch(img, flags, a_min, denom, bmin, bmax, out=out)
    return img
