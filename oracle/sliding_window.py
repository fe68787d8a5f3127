"""CPU restatement of the reference's sliding-window inference (TEST INFRASTRUCTURE).

Follows ``/root/reference/engine/utils.py:19-159`` (itself a fork of MONAI 0.8
``sliding_window_inference``) and the label post-processing of
``/root/reference/engine/test.py:140-141``.  Checked bit-for-bit against the
reference file itself (run verbatim under ``oracle/monai_shim``) by
``tests/test_oracle_golden.py`` via the fixtures ``tests/golden/make_golden.py``
writes.  See ``oracle/__init__.py`` for the pinning status.
"""
from __future__ import annotations

from typing import Any, Callable, List, Optional, Sequence, Tuple, Union

import numpy as np
import torch
import torch.nn.functional as F

from . import monai08 as M


def window_grid(
    image_size: Sequence[int], roi_size: Sequence[int], overlap: float
) -> Tuple[Tuple[int, ...], List[List[int]]]:
    """Scan interval and per-axis window starts (engine/utils.py:105-108)."""
    interval = M.get_scan_interval(image_size, roi_size, len(image_size), overlap)
    patch = M.get_valid_patch_size(image_size, roi_size)
    return interval, [M.axis_starts(image_size[d], patch[d], interval[d]) for d in range(len(image_size))]


def window_centers(stops: Sequence[int], roi: Sequence[int], image_size: Sequence[int]) -> List[float]:
    """Relative window centre handed to the predictor (engine/utils.py:126-128):
    ``(slice.stop - roi // 2) / image_size`` per axis, as Python floats."""
    return [(stops[d] - roi[d] // 2) / image_size[d] for d in range(3)]


def sliding_window_inference(
    inputs: torch.Tensor,
    affine: Optional[torch.Tensor],
    roi_size: Union[Sequence[int], int],
    sw_batch_size: int,
    predictor: Callable[..., torch.Tensor],
    overlap: float = 0.25,
    mode: Any = M.BlendMode.CONSTANT,
    sigma_scale: Union[Sequence[float], float] = 0.125,
    padding_mode: Any = M.PytorchPadMode.CONSTANT,
    cval: float = 0.0,
    *args: Any,
    tuple_input: bool = True,
    importance_map: Optional[torch.Tensor] = None,
    on_window_batch: Optional[Callable[[int, torch.Tensor, torch.Tensor], None]] = None,
    **kwargs: Any,
) -> torch.Tensor:
    """Restatement of engine/utils.py:19-159 on CPU tensors.

    ``tuple_input=True`` feeds the predictor the reference's 3-tuple
    ``(patches, centers, affine)`` (engine/utils.py:134); ``False`` feeds the
    plain patch tensor like stock MONAI does for ``run_evaluation.py:68-74``
    (quirk Q6).  ``importance_map`` overrides the computed map (test hook);
    ``on_window_batch(first_window, patches, logits)`` observes every predictor call.
    """
    nsp = inputs.dim() - 2
    if overlap < 0 or overlap >= 1:  # engine/utils.py:82-83
        raise AssertionError("overlap must be >= 0 and < 1.")
    orig_size = list(inputs.shape[2:])
    nb = inputs.shape[0]

    roi = M.fall_back_tuple(roi_size, orig_size)  # :95
    image_size = tuple(max(orig_size[i], roi[i]) for i in range(nsp))  # :97
    pad: List[int] = []  # :98-103, F.pad order = last dim first
    for k in range(inputs.dim() - 1, 1, -1):
        diff = max(roi[k - 2] - inputs.shape[k], 0)
        pad.extend([diff // 2, diff - diff // 2])
    inputs = F.pad(inputs, pad=pad, mode=M.look_up_option(padding_mode, M.PytorchPadMode).value, value=cval)

    interval = M.get_scan_interval(image_size, roi, nsp, overlap)  # :105
    windows = M.dense_patch_slices(image_size, roi, interval)  # :108
    n_win = len(windows)
    total = n_win * nb

    if importance_map is None:  # :113-115
        importance_map = M.compute_importance_map(
            M.get_valid_patch_size(image_size, roi), mode=mode, sigma_scale=sigma_scale, device="cpu"
        )

    out = cnt = None
    for first in range(0, total, sw_batch_size):  # :120
        idxs = range(first, min(first + sw_batch_size, total))
        where = []
        for idx in idxs:  # :122-125 (int(idx / num_win): float division, quirk Q4)
            b = int(idx / n_win)
            where.append((slice(b, b + 1), slice(None)) + tuple(windows[idx % n_win]))
        centers = torch.stack(  # :126-130
            [torch.tensor(window_centers([w[2].stop, w[3].stop, w[4].stop], roi, image_size)) for w in where]
        ).float()
        if sw_batch_size == 1:  # :131-132, quirk Q3
            centers = centers.unsqueeze(0)
        patches = torch.cat([inputs[w] for w in where])  # :133
        model_in = (patches, centers, affine) if tuple_input else patches  # :134
        logits = predictor(model_in, *args, **kwargs)  # :135
        if on_window_batch is not None:
            on_window_batch(first, patches, logits)
        if out is None:  # :137-143
            shape = [nb, logits.shape[1]] + list(image_size)
            out = torch.zeros(shape, dtype=torch.float32)
            cnt = torch.zeros(shape, dtype=torch.float32)
        for j, w in enumerate(where):  # :146-148 - product rounded, then add rounded, ascending windows
            out[w] += importance_map * logits[j]
            cnt[w] += importance_map
    out = out / cnt  # :151

    crop: List[slice] = []  # :153-159
    for sp in range(nsp):
        lo = pad[sp * 2]
        crop.insert(0, slice(lo, orig_size[nsp - sp - 1] + lo))
    while len(crop) < out.dim():
        crop.insert(0, slice(None))
    return out[tuple(crop)]


def labels_from_logits(logits: torch.Tensor) -> np.ndarray:
    """engine/test.py:140-141 (and :81-82): softmax over classes, to NumPy, first-max argmax, uint8, batch element 0."""
    probs = torch.softmax(logits, 1).cpu().numpy()
    return np.argmax(probs, axis=1).astype(np.uint8)[0]


def top2_relative_gap(logits: torch.Tensor) -> np.ndarray:
    """Per-voxel ``(top1 - top2) / max(|top1|, |top2|)`` of ``logits[0]`` - the quantity the parity
    criterion of BASELINE.json ("labels bit-exact except voxels whose top-2 gap is below tolerance") is stated on."""
    top = torch.topk(logits[0].float(), 2, dim=0).values
    denom = torch.maximum(top[0].abs(), top[1].abs()).clamp_min(torch.finfo(torch.float32).tiny)
    return ((top[0] - top[1]) / denom).numpy()
