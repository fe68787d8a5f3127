"""Restatement of the MONAI 0.8.x helpers used by the reference's hot path.

TEST INFRASTRUCTURE (see ``oracle/__init__.py``).  MONAI is a third-party
dependency of the reference that is neither vendored nor version-pinned
(``/root/reference/requirements.txt:1``); the 0.8.x vintage is inferred from
the API the reference uses.  These functions restate the
published MONAI 0.8.1 algorithms; the reference call site each one serves is
cited per function.  Parity for THIS file is unpinned by the reference (no
golden vectors exist); it is anchored on the reference's call sites and on the
known-answer tables in ``tests/test_oracle_kat.py``.
"""
from __future__ import annotations

import math
from enum import Enum
from typing import Any, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F


class BlendMode(Enum):
    """``monai.utils.BlendMode`` - imported at engine/utils.py:8, default arg at :26."""

    CONSTANT = "constant"
    GAUSSIAN = "gaussian"


class PytorchPadMode(Enum):
    """``monai.utils.PytorchPadMode`` - imported at engine/utils.py:9, used at :28,:103."""

    CONSTANT = "constant"
    REFLECT = "reflect"
    REPLICATE = "replicate"
    CIRCULAR = "circular"


def look_up_option(opt: Any, supported: Any, default: Any = "no_default") -> Any:
    """``monai.utils.look_up_option`` - engine/utils.py:103 (``look_up_option(padding_mode, PytorchPadMode).value``).

    Accepts an Enum member or its string value and returns the member.
    """
    if isinstance(supported, type) and issubclass(supported, Enum):
        if isinstance(opt, supported):
            return opt
        if isinstance(opt, str):
            opt = opt.strip()
        for member in supported:
            if member.value == opt:
                return member
        if default != "no_default":
            return default
        raise ValueError(f"Unsupported option '{opt}', Available options are {[m.value for m in supported]}.")
    if opt in supported:
        return supported[opt] if isinstance(supported, dict) else opt
    if default != "no_default":
        return default
    raise ValueError(f"Unsupported option '{opt}'.")


def optional_import(module: str, name: str = "", **_: Any) -> Tuple[Any, bool]:
    """``monai.utils.optional_import`` - engine/utils.py:15 (only for tqdm, never called later)."""
    try:
        mod = __import__(module, fromlist=[name] if name else [])
        return (getattr(mod, name) if name else mod), True
    except Exception:  # noqa: BLE001 - mirrors MONAI's lazy failure
        return None, False


def _issequence(x: Any) -> bool:
    return isinstance(x, (list, tuple, np.ndarray, torch.Size)) or (
        hasattr(x, "__iter__") and not isinstance(x, (str, bytes))
    )


def ensure_tuple_rep(tup: Any, dim: int) -> Tuple[Any, ...]:
    """``monai.utils.ensure_tuple_rep``: scalar -> repeated; sequence of len ``dim`` -> tuple; else ValueError."""
    if not _issequence(tup):
        return (tup,) * dim
    if len(tup) == dim:
        return tuple(tup)
    raise ValueError(f"Sequence must have length {dim}, got {len(tup)}.")


def ensure_tuple_size(tup: Any, dim: int, pad_val: Any = 0) -> Tuple[Any, ...]:
    """``monai.utils.ensure_tuple_size``: pad with ``pad_val`` / truncate to ``dim`` entries."""
    tup = tuple(tup) if _issequence(tup) else (tup,)
    return (tup + (pad_val,) * dim)[:dim]


def fall_back_tuple(user_provided: Any, default: Sequence[int]) -> Tuple[int, ...]:
    """``monai.utils.fall_back_tuple`` - engine/utils.py:95.

    Components of ``user_provided`` that are None or non-positive fall back to
    the matching component of ``default`` (the image size).
    """
    ndim = len(default)
    user = ensure_tuple_rep(user_provided, ndim)
    return tuple(u if (u is not None and u and u > 0) else d for u, d in zip(user, default))


def get_scan_interval(
    image_size: Sequence[int], roi_size: Sequence[int], num_spatial_dims: int, overlap: float
) -> Tuple[int, ...]:
    """``monai.inferers.utils._get_scan_interval`` - engine/utils.py:105.

    Per dim: the roi itself when it spans the whole image, else
    ``int(roi * (1 - overlap))`` floored at 1.
    """
    if len(image_size) != num_spatial_dims:
        raise ValueError("image coord different from spatial dims.")
    if len(roi_size) != num_spatial_dims:
        raise ValueError("roi coord different from spatial dims.")
    out = []
    for i in range(num_spatial_dims):
        if roi_size[i] == image_size[i]:
            out.append(int(roi_size[i]))
        else:
            step = int(roi_size[i] * (1 - overlap))
            out.append(step if step > 0 else 1)
    return tuple(out)


_get_scan_interval = get_scan_interval  # the name engine/utils.py:6 imports


def get_valid_patch_size(image_size: Sequence[int], patch_size: Any) -> Tuple[int, ...]:
    """``monai.data.utils.get_valid_patch_size`` - engine/utils.py:114: ``min(image, patch or image)`` per dim."""
    ndim = len(image_size)
    patch = ensure_tuple_size(patch_size, ndim)
    return tuple(min(ms, ps or ms) for ms, ps in zip(image_size, patch))


def axis_starts(image: int, patch: int, interval: int) -> list:
    """Window starts along one axis (the per-dim body of ``dense_patch_slices``).

    ``scan_num = 1 + first d in range(ceil(image/interval)) with d*interval + patch >= image``
    and the start of each window is pulled back so it ends inside the image
    (the LAST window is clamped, not padded - quirk Q1 of SURVEY.md section 8).
    """
    if interval == 0:
        num = 1
    else:
        upper = int(math.ceil(float(image) / interval))
        first = next((d for d in range(upper) if d * interval + patch >= image), None)
        num = first + 1 if first is not None else 1
    starts = []
    for idx in range(num):
        s = idx * interval
        s -= max(s + patch - image, 0)
        starts.append(s)
    return starts


def dense_patch_slices(
    image_size: Sequence[int], patch_size: Sequence[int], scan_interval: Sequence[int]
) -> list:
    """``monai.data.utils.dense_patch_slices`` - engine/utils.py:108.

    Cartesian product of the per-axis starts in C order (first spatial axis
    slowest, last fastest) as tuples of ``slice(start, start + patch)``.
    """
    ndim = len(image_size)
    patch = get_valid_patch_size(image_size, patch_size)
    interval = ensure_tuple_size(scan_interval, ndim)
    starts = [axis_starts(image_size[d], patch[d], interval[d]) for d in range(ndim)]
    grid = np.asarray([g.flatten() for g in np.meshgrid(*starts, indexing="ij")]).T
    return [tuple(slice(int(s), int(s) + patch[d]) for d, s in enumerate(row)) for row in grid]


def gaussian_1d(sigma: torch.Tensor, truncated: float = 4.0) -> torch.Tensor:
    """``monai.networks.layers.gaussian_1d(approx="erf")`` (not renormalised).

    ``tail = int(max(sigma*truncated, 0.5) + 0.5)``; taps ``x = -tail..tail``;
    ``0.5 * (erf(t(x+.5)) - erf(t(x-.5)))`` with ``t = 0.70710678/|sigma|``, clamped at 0.
    All arithmetic in float32, as in MONAI.
    """
    sigma = torch.as_tensor(sigma, dtype=torch.float)
    tail = int(max(float(sigma) * truncated, 0.5) + 0.5)
    x = torch.arange(-tail, tail + 1, dtype=torch.float)
    t = 0.70710678 / torch.abs(sigma)
    out = 0.5 * ((t * (x + 0.5)).erf() - (t * (x - 0.5)).erf())
    return out.clamp(min=0)


def _separable_zero_padded_filter(vol: torch.Tensor, kernels: Sequence[torch.Tensor]) -> torch.Tensor:
    """``monai.networks.layers.separable_filtering`` with zero padding: axis 0 first, last axis last."""
    ndim = vol.dim()
    x = vol[None, None]
    conv = [F.conv1d, F.conv2d, F.conv3d][ndim - 1]
    for d in range(ndim):
        k = kernels[d]
        shape = [1, 1] + [1] * ndim
        shape[d + 2] = -1
        pad = [0] * ndim
        pad[d] = (k.numel() - 1) // 2
        x = conv(x, k.reshape(shape), padding=tuple(pad))
    return x[0, 0]


def compute_importance_map(
    patch_size: Sequence[int],
    mode: Any = BlendMode.CONSTANT,
    sigma_scale: Any = 0.125,
    device: Any = "cpu",
) -> torch.Tensor:
    """``monai.data.utils.compute_importance_map`` (0.8.x) - engine/utils.py:113-115.

    gaussian: unit impulse at ``i // 2`` per axis, separable zero-padded
    gaussian filter with ``sigma = sigma_scale * patch`` per axis, divide by the
    maximum, cast to float32 and clamp to the smallest non-zero value.
    """
    mode = look_up_option(mode, BlendMode)
    patch_size = tuple(int(p) for p in patch_size)
    if mode == BlendMode.CONSTANT:
        return torch.ones(patch_size, dtype=torch.float, device=device)
    centre = tuple(p // 2 for p in patch_size)
    scales = ensure_tuple_rep(sigma_scale, len(patch_size))
    sigmas = [p * s for p, s in zip(patch_size, scales)]
    impulse = torch.zeros(patch_size, dtype=torch.float)
    impulse[centre] = 1
    kernels = [gaussian_1d(torch.as_tensor(s, dtype=torch.float)) for s in sigmas]
    imap = _separable_zero_padded_filter(impulse, kernels)
    imap = (imap / torch.max(imap)).float()
    min_non_zero = imap[imap != 0].min().item()
    return torch.clamp(imap, min=min_non_zero).to(device)


def compute_importance_map_v12(patch_size: Sequence[int], sigma_scale: Any = 0.125) -> torch.Tensor:
    """The MONAI >= 1.2 gaussian (version hazard): ``exp(-x^2/2sigma^2)`` on the
    half-integer grid ``x = -(n-1)/2 .. (n-1)/2`` as an outer product, floored at ``max(min, 1e-3)``."""
    patch_size = tuple(int(p) for p in patch_size)
    scales = ensure_tuple_rep(sigma_scale, len(patch_size))
    sigmas = [p * s for p, s in zip(patch_size, scales)]
    imap: Optional[torch.Tensor] = None
    for i, n in enumerate(patch_size):
        x = torch.arange(start=-(n - 1) / 2.0, end=(n - 1) / 2.0 + 1, dtype=torch.float)
        x = torch.exp(x**2 / (-2 * sigmas[i] ** 2))
        imap = x if imap is None else imap.unsqueeze(-1) * x[(None,) * i]
    floor = max(torch.min(imap).item(), 1e-3)
    return torch.clamp(imap.to(torch.float), min=floor)
