"""CPU oracle (TEST INFRASTRUCTURE) for the 95th-percentile Hausdorff distance of engine/test.py:31,55-57.

Restates MONAI 0.8.1 ``monai/metrics/hausdorff_distance.py::compute_hausdorff_distance`` and the helpers it calls from
``monai/metrics/utils.py`` (``get_mask_edges``, ``get_surface_distance``, ``do_metric_reduction``) with the scipy / NumPy
calls MONAI itself makes (binary_erosion, distance_transform_edt, np.percentile).  MONAI is a third-party dependency
absent from /root/reference and un-pinned in its requirements.txt, and the reference holds no test or fixture for this
metric: PARITY UNPINNED - anchored on the call site only.
"""
from __future__ import annotations

from typing import Optional, Tuple

import numpy as np


def get_mask_edges(seg_pred: np.ndarray, seg_gt: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
    """monai.metrics.utils.get_mask_edges(label_idx=1, crop=True) on boolean masks."""
    from scipy.ndimage import binary_erosion

    if not np.any(seg_pred | seg_gt):
        return np.zeros_like(seg_pred), np.zeros_like(seg_gt)
    union = seg_pred | seg_gt
    sl = []
    for a in range(union.ndim):  # generate_spatial_bounding_box
        other = tuple(x for x in range(union.ndim) if x != a)
        idx = np.nonzero(union.any(axis=other))[0]
        sl.append(slice(int(idx[0]), int(idx[-1]) + 1))
    seg_pred, seg_gt = np.squeeze(seg_pred[tuple(sl)][None]), np.squeeze(seg_gt[tuple(sl)][None])
    return binary_erosion(seg_pred) ^ seg_pred, binary_erosion(seg_gt) ^ seg_gt


def get_surface_distance(seg_pred: np.ndarray, seg_gt: np.ndarray) -> np.ndarray:
    """monai.metrics.utils.get_surface_distance(distance_metric='euclidean')."""
    from scipy.ndimage import distance_transform_edt

    if not np.any(seg_gt):
        dis = np.inf * np.ones_like(seg_gt)
    else:
        if not np.any(seg_pred):
            dis = np.inf * np.ones_like(seg_gt)
            return np.asarray(dis[seg_gt])
        dis = distance_transform_edt(~seg_gt)
    return np.asarray(dis[seg_pred])


def percent_hausdorff(edges_pred: np.ndarray, edges_gt: np.ndarray, percentile: Optional[float]) -> float:
    d = get_surface_distance(edges_pred, edges_gt)
    if d.shape == (0,):
        return float("nan")
    if not percentile:
        return float(d.max())
    with np.errstate(invalid="ignore"):
        return float(np.percentile(d, percentile))


def hausdorff_distance(pred: np.ndarray, label: np.ndarray, n_classes: int, percentile: Optional[float] = 95,
                       include_background: bool = True, directed: bool = False) -> np.ndarray:
    """compute_hausdorff_distance on the one-hot channels of two label maps, per class."""
    out = []
    for c in range(0 if include_background else 1, n_classes):
        ep, eg = get_mask_edges(pred == c, label == c)
        d1 = percent_hausdorff(ep, eg, percentile)
        out.append(d1 if directed else max(d1, percent_hausdorff(eg, ep, percentile)))
    return np.asarray(out, dtype=np.float64)


def mean_hausdorff(hd: np.ndarray) -> Tuple[float, int]:
    """do_metric_reduction(f, 'mean') as HausdorffDistanceMetric.aggregate() applies it."""
    f = np.atleast_2d(np.asarray(hd, dtype=np.float64)).copy()
    nans = np.isnan(f)
    not_nans = (~nans).astype(np.float64)
    f[nans] = 0
    nn = not_nans.sum(0)
    with np.errstate(invalid="ignore", divide="ignore"):
        f = np.where(nn > 0, f.sum(0) / nn, 0.0)
    n = int((nn > 0).sum())
    return (float(f.sum() / n) if n > 0 else 0.0), n
