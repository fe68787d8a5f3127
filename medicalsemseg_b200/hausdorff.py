"""95th-percentile Hausdorff distance on the GPU: what engine/test.py:31,55-57 gets out of
``HausdorffDistanceMetric(include_background=True, percentile=95, reduction="mean", get_not_nans=True)`` (MONAI 0.8:
``compute_hausdorff_distance`` -> ``get_mask_edges`` + ``get_surface_distance`` + ``np.percentile``), computed from the
uint8 label maps instead of one-hot tensors.

Per class: bounding box of pred | gt, surface voxels of both masks (``mss_mask_edges``), the exact squared Euclidean
distance transform of each surface (three ``mss_edt_pass`` launches), the distances at the other surface's voxels,
float64 sqrt and NumPy's linear-interpolation percentile.  Box search, transposes, boolean gathers and the sort are torch
plumbing; the two stencil / scan kernels are libmss_b200.so."""
from __future__ import annotations

from typing import Any, Optional, Tuple

import numpy as np
import torch

from . import _lib


def _bbox(mask: torch.Tensor) -> Optional[Tuple[Tuple[int, int, int], Tuple[int, int, int]]]:
    """``generate_spatial_bounding_box`` of a boolean volume (None when empty)."""
    lo, hi = [], []
    for a in range(3):
        other = tuple(x for x in range(3) if x != a)
        idx = torch.nonzero(mask.any(dim=other)).flatten()
        if idx.numel() == 0:
            return None
        lo.append(int(idx[0]))
        hi.append(int(idx[-1]) + 1)
    return tuple(lo), tuple(hi)


def _edges(labels: torch.Tensor, cls: int, lo, hi) -> torch.Tensor:
    lib = _lib.load()
    n = tuple(h - l for l, h in zip(lo, hi))
    edges = torch.empty(n, dtype=torch.uint8, device=labels.device)
    rc = lib.mss_mask_edges(labels.data_ptr(), _lib.I3(*labels.shape), int(cls), _lib.I3(*lo), _lib.I3(*hi), edges.data_ptr(),
                            None, torch.cuda.current_stream().cuda_stream)
    _lib.check(rc, "mss_mask_edges")
    return edges


def squared_edt(feature: torch.Tensor) -> torch.Tensor:
    """Exact squared distance to the nearest non-zero voxel of the uint8 volume ``feature`` (int32; 2**29 when there is
    none), returned in the TRANSPOSED layout ``[D, W, H]``: the pass along the contiguous axis runs on a transposed copy
    so its line loop is coalesced too.  An int32 ``feature`` is taken as the transform's input itself (0 on features,
    2**29 elsewhere)."""
    lib = _lib.load()
    stream = torch.cuda.current_stream().cuda_stream
    d, h, w = feature.shape
    t1 = torch.empty((d, h, w), dtype=torch.int32, device=feature.device)
    t2, s, t = torch.empty_like(t1), torch.empty_like(t1), torch.empty_like(t1)
    dims = _lib.I3(d, h, w)
    if feature.dtype == torch.uint8:
        rc = lib.mss_edt_pass_mask(feature.contiguous().data_ptr(), t1.data_ptr(), s.data_ptr(), t.data_ptr(), dims, 0, stream)
    else:
        rc = lib.mss_edt_pass(feature.contiguous().data_ptr(), t1.data_ptr(), s.data_ptr(), t.data_ptr(), dims, 0, stream)
    _lib.check(rc, "mss_edt_pass")
    _lib.check(lib.mss_edt_pass(t1.data_ptr(), t2.data_ptr(), s.data_ptr(), t.data_ptr(), dims, 1, stream), "mss_edt_pass")
    at = t2.transpose(1, 2).contiguous()  # [D, W, H]
    bt = torch.empty_like(at)
    _lib.check(lib.mss_edt_pass(at.data_ptr(), bt.data_ptr(), s.data_ptr(), t.data_ptr(), _lib.I3(d, w, h), 1, stream),
               "mss_edt_pass")
    return bt


def _np_percentile(sorted_vals: torch.Tensor, q: float) -> float:
    """``numpy.percentile(x, q)`` (linear) of an ascending float64 CUDA vector, NumPy's own arithmetic (incl. inf - inf)."""
    n = sorted_vals.numel()
    quant = q / 100.0
    pos = (n - 1) * quant  # numpy's virtual index for method='linear'
    if pos >= n - 1:
        lo = hi = n - 1
    elif pos < 0:
        lo = hi = 0
    else:
        lo = int(np.floor(pos))
        hi = lo + 1
    t = pos - np.floor(pos)
    a, b = float(sorted_vals[lo]), float(sorted_vals[hi])
    with np.errstate(invalid="ignore"):
        d = np.float64(b) - np.float64(a)
        return float(np.float64(b) - d * (1.0 - t)) if t >= 0.5 else float(np.float64(a) + d * t)


def _percent_distance(edges_from: torch.Tensor, edges_to: torch.Tensor, percentile: Optional[float]) -> float:
    """``compute_percent_hausdorff_distance(edges_from, edges_to)``: distances from the voxels of one surface to the other."""
    n_from, n_to = int(edges_from.sum()), int(edges_to.sum())
    inf = float("inf")
    if n_to == 0:      # get_surface_distance: `dis = inf * ones; return dis[seg_pred]`
        dist = torch.full((n_from,), inf, dtype=torch.float64, device=edges_from.device)
    elif n_from == 0:  # ... `if not np.any(seg_pred): return dis[seg_gt]` - infinities for the OTHER surface's voxels
        dist = torch.full((n_to,), inf, dtype=torch.float64, device=edges_from.device)
    else:
        dt_t = squared_edt(edges_to)             # [D, W, H]
        dist = dt_t[edges_from.transpose(1, 2).bool()].to(torch.float64).sqrt()
    if dist.numel() == 0:
        return float("nan")                      # surface_distance.shape == (0,)
    if not percentile:
        return float(dist.max())
    return _np_percentile(torch.sort(dist).values, float(percentile))


def hausdorff_distance(pred: torch.Tensor, label: torch.Tensor, n_classes: int, percentile: Optional[float] = 95,
                       include_background: bool = True, directed: bool = False) -> np.ndarray:
    """Per-class (percentile) Hausdorff distance in voxels, float64 ``[K]`` (``[K-1]`` without background): NaN when a
    class is absent from both maps, inf / NaN by NumPy's rules when it is absent from one."""
    if not (pred.is_cuda and label.is_cuda):
        raise _lib.MssError("hausdorff_distance needs CUDA tensors; there is no CPU fallback")
    p = pred.reshape(pred.shape[-3:]).to(torch.uint8).contiguous()
    y = label.reshape(label.shape[-3:])
    y = (y if y.dtype == torch.uint8 else y.round().clamp(0, 255).to(torch.uint8)).contiguous()
    if p.shape != y.shape:
        raise ValueError(f"pred {tuple(p.shape)} and label {tuple(y.shape)} differ")
    out = []
    with torch.cuda.device(p.device):
        for c in range(0 if include_background else 1, n_classes):
            box = _bbox((p == c) | (y == c))
            if box is None:
                out.append(float("nan"))
                continue
            lo, hi = box
            ep = _edges(p, c, lo, hi)
            ey = _edges(y, c, lo, hi)
            d1 = _percent_distance(ep, ey, percentile)
            if directed:
                out.append(d1)
                continue
            d2 = _percent_distance(ey, ep, percentile)
            out.append(max(d1, d2))               # Python's max, as the reference calls it (NaN handling included)
    return np.asarray(out, dtype=np.float64)


def mean_hausdorff(hd: Any) -> Tuple[float, int]:
    """``HausdorffDistanceMetric(reduction="mean", get_not_nans=True).aggregate()`` for one volume ``[K]`` or a batch
    ``[B, K]``: mean over the batch per class ignoring NaNs, then over the classes that had any value; returns
    ``(value, not_nans)``."""
    f = np.atleast_2d(np.asarray(hd, dtype=np.float64)).copy()
    nans = np.isnan(f)
    not_nans = (~nans).astype(np.float64)
    f[nans] = 0
    nn_c = not_nans.sum(0)
    with np.errstate(invalid="ignore", divide="ignore"):
        per_class = np.where(nn_c > 0, f.sum(0) / nn_c, 0.0)
    n = int((nn_c > 0).sum())
    return (float(per_class.sum() / n) if n > 0 else 0.0), n
