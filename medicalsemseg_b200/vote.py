"""Ensemble majority vote on the GPU: the arithmetic of majority_vote.py:23-37 (file/NIfTI handling is the caller's)."""
from __future__ import annotations

import ctypes as C
from typing import Any, Optional, Sequence, Union

import numpy as np
import torch

from . import _lib


def _to_u8_cuda(m: Any, device: torch.device) -> torch.Tensor:
    """Label maps arrive as uint8 tensors, or as the float64 arrays nibabel's get_fdata() yields
    (majority_vote.py:20).  Values that are not integers in [0, 254] can match no class: map them to 255."""
    t = torch.as_tensor(m)
    if t.dtype != torch.uint8:
        t = t.to(device)
        ok = (t == t.round()) & (t >= 0) & (t <= 254)
        t = torch.where(ok, t, torch.full_like(t, 255)).to(torch.uint8)
    return t.to(device).contiguous()


def majority_vote(label_maps: Union[torch.Tensor, Sequence[Any]], n_classes: Optional[int] = None,
                  device: Any = None) -> torch.Tensor:
    """Per-voxel vote over ``M`` label maps ``[X, Y, Z]`` -> uint8 ``[X, Y, Z]`` on the GPU.

    Semantics of majority_vote.py:23-37 exactly: background holds the constant single vote (it is never
    counted), class ``c >= 1`` gets one vote per map equal to ``c``, labels ``>= n_classes`` are ignored
    and the first maximum wins - i.e. the lowest foreground class with the maximal count wins iff that
    count is >= 2.  ``n_classes`` defaults to ``max label + 1``.
    """
    if not torch.cuda.is_available():
        raise _lib.MssError("medicalsemseg_b200 needs a CUDA device (B200); there is no CPU fallback")
    maps = list(label_maps) if not isinstance(label_maps, torch.Tensor) else list(label_maps.unbind(0))
    if not maps:
        raise ValueError("majority_vote needs at least one label map")
    dev = torch.device(device) if device is not None else (
        maps[0].device if isinstance(maps[0], torch.Tensor) and maps[0].is_cuda else torch.device("cuda", torch.cuda.current_device()))
    maps = [_to_u8_cuda(m, dev) for m in maps]
    shape = maps[0].shape
    if any(m.shape != shape for m in maps):
        raise ValueError("all label maps must have the same shape")
    if n_classes is None:
        n_classes = int(max(int(m.max().item()) for m in maps)) + 1
    if len(maps) > _lib.MAX_VOTE_MAPS or n_classes > _lib.MAX_VOTE_CLASSES:
        raise _lib.MssError(f"majority_vote supports up to {_lib.MAX_VOTE_MAPS} maps and {_lib.MAX_VOTE_CLASSES} classes")
    out = torch.empty(shape, dtype=torch.uint8, device=dev)
    ptrs = (C.c_void_p * len(maps))(*[m.data_ptr() for m in maps])
    with torch.cuda.device(dev):
        rc = _lib.load().mss_majority_vote(ptrs, len(maps), int(n_classes), maps[0].numel(), out.data_ptr(),
                                           torch.cuda.current_stream().cuda_stream)
    _lib.check(rc, "mss_majority_vote")
    return out


def get_new_label(fdata: Sequence[np.ndarray], n_folds: int, n_classes: int) -> np.ndarray:
    """Call-compatible with majority_vote.py:35-37: tuple of per-fold arrays in, int64 NumPy label map out."""
    return majority_vote(list(fdata)[:n_folds], n_classes).cpu().numpy().astype(np.int64)
