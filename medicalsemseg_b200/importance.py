"""Importance-map construction (MONAI ``compute_importance_map`` at engine/utils.py:113-115).

The 3-D map is generated on the GPU (``mss_importance_map``: rounded outer product, /max, clamp).
Its three 1-D factors are ~100 floats each.  By default they are evaluated on the host with the very
torch ops MONAI uses (float32 ``erf``): float32 erf implementations differ by 1 ulp between CPU and
GPU, and the difference ``erf(a) - erf(b)`` near |erf| = 1 amplifies that to ~1e-3 relative in the
tails of the window, which would break parity with the reference's CPU path at the 1e-5 level.
``taps="device"`` evaluates them in the kernel instead (``mss_gaussian_profile``).
"""
from __future__ import annotations

from typing import Any, Dict, Sequence, Tuple

import torch

from . import _lib
from .grid import BLEND_MODES, _option


def _tuple3(x: Any) -> Tuple[float, float, float]:
    if isinstance(x, (list, tuple)):
        if len(x) != 3:
            raise ValueError(f"Sequence must have length 3, got {len(x)}.")
        return tuple(float(v) for v in x)
    return (float(x),) * 3


def monai08_profile(n: int, sigma: float) -> torch.Tensor:
    """Per-axis factor of the MONAI 0.8 gaussian map: ``gaussian_1d(sigma, truncated=4, approx='erf')``
    cross-correlated with a unit impulse at ``n // 2`` under zero padding -> ``profile[i] = tap[n//2 + tail - i]``."""
    sig = torch.as_tensor(sigma, dtype=torch.float)
    tail = int(max(float(sig) * 4.0, 0.5) + 0.5)
    x = torch.arange(-tail, tail + 1, dtype=torch.float)
    t = 0.70710678 / torch.abs(sig)
    taps = (0.5 * ((t * (x + 0.5)).erf() - (t * (x - 0.5)).erf())).clamp(min=0)
    idx = n // 2 + tail - torch.arange(n)
    ok = (idx >= 0) & (idx < taps.numel())
    return torch.where(ok, taps[idx.clamp(0, taps.numel() - 1)], torch.zeros((), dtype=torch.float))


def monai12_profile(n: int, sigma: float) -> torch.Tensor:
    """Per-axis factor of the MONAI >= 1.2 map: ``exp(x^2 / (-2 sigma^2))`` on ``-(n-1)/2 .. (n-1)/2``."""
    x = torch.arange(start=-(n - 1) / 2.0, end=(n - 1) / 2.0 + 1, dtype=torch.float)
    return torch.exp(x**2 / (-2 * sigma**2))


_CACHE: Dict[Any, torch.Tensor] = {}


def importance_map(roi: Sequence[int], mode: Any = "gaussian", sigma_scale: Any = 0.125,
                   device: Any = "cuda", variant: str = "monai08", taps: str = "host") -> torch.Tensor:
    """``[roi_d, roi_h, roi_w]`` float32 window weights on ``device`` (cached per argument set)."""
    mode = _option(mode, BLEND_MODES, "mode")
    roi = tuple(int(r) for r in roi)
    scales = _tuple3(sigma_scale)
    device = torch.device(device)
    if device.type != "cuda":
        raise _lib.MssError("importance_map: medicalsemseg_b200 runs on CUDA devices only")
    if device.index is None:
        device = torch.device("cuda", torch.cuda.current_device())
    key = (roi, mode, scales, device, variant, taps)
    hit = _CACHE.get(key)
    if hit is not None:
        return hit
    if variant not in ("monai08", "monai12"):
        raise ValueError(f"unknown importance-map variant '{variant}'")
    lib = _lib.load()
    with torch.cuda.device(device):
        stream = torch.cuda.current_stream().cuda_stream
        out = torch.empty(roi, dtype=torch.float32, device=device)
        roi_c = _lib.I3(*roi)
        if mode == "constant":
            _lib.check(lib.mss_importance_map(out.data_ptr(), roi_c, _lib.BLEND_CONSTANT, None, None, None, 0.0, None,
                                              stream), "mss_importance_map")
        else:
            sigmas = [r * s for r, s in zip(roi, scales)]
            profs = []
            for n, sg in zip(roi, sigmas):
                if taps == "host":
                    host = monai08_profile(n, sg) if variant == "monai08" else monai12_profile(n, sg)
                    profs.append(host.to(device))
                elif taps == "device":
                    p = torch.empty(n, dtype=torch.float32, device=device)
                    var = _lib.GAUSS_MONAI08_ERF if variant == "monai08" else _lib.GAUSS_MONAI12_EXP
                    _lib.check(lib.mss_gaussian_profile(p.data_ptr(), n, float(sg), var, stream), "mss_gaussian_profile")
                    profs.append(p)
                else:
                    raise ValueError("taps must be 'host' or 'device'")
            scratch = torch.empty(2, dtype=torch.int32, device=device)
            floor_abs = 1e-3 if variant == "monai12" else 0.0
            _lib.check(lib.mss_importance_map(out.data_ptr(), roi_c, _lib.BLEND_PROFILES, profs[0].data_ptr(),
                                              profs[1].data_ptr(), profs[2].data_ptr(), floor_abs, scratch.data_ptr(),
                                              stream), "mss_importance_map")
    _CACHE[key] = out
    return out
