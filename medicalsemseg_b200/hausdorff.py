"""95th-percentile Hausdorff distance on the GPU: what engine/test.py:31,55-57 gets out of
``HausdorffDistanceMetric(include_background=True, percentile=95, reduction="mean", get_not_nans=True)`` (MONAI 0.8:
``compute_hausdorff_distance`` -> ``get_mask_edges`` + ``get_surface_distance`` + ``np.percentile``), computed from the
uint8 label maps instead of one-hot tensors.

Per class: bounding box of pred | gt, surface voxels of both masks (``mss_mask_edges``), the exact squared Euclidean
distance transform of each surface (three ``mss_edt_pass`` launches), the distances at the other surface's voxels,
float64 sqrt and NumPy's linear-interpolation percentile.  Round 2: everything data-sized runs in libmss_b200.so kernels -
``mss_class_boxes`` (all boxes in one pass), ``mss_mask_edges``, ``mss_edt_row_mask`` + ``mss_edt_pass`` x 2,
``mss_select2`` (the order statistics np.percentile interpolates, by radix histograms masked with the surface: no boolean
gather, no sort) - with one host sync for the boxes and one device-to-host copy of the results.  The 2 K directed
distances are independent, and one envelope pass of the distance transform fills only part of the GPU (one thread per line:
0.6 waves of CTAs at a third of the issue rate), so they are spread over a few CUDA streams and run side by side."""
from __future__ import annotations

import os
from typing import Any, Optional, Tuple

import numpy as np
import torch

from . import _lib


def _edges(labels: torch.Tensor, cls: int, lo, hi) -> torch.Tensor:
    lib = _lib.load()
    n = tuple(h - l for l, h in zip(lo, hi))
    edges = torch.empty(n, dtype=torch.uint8, device=labels.device)
    rc = lib.mss_mask_edges(labels.data_ptr(), _lib.I3(*labels.shape), int(cls), _lib.I3(*lo), _lib.I3(*hi), edges.data_ptr(),
                            None, torch.cuda.current_stream().cuda_stream)
    _lib.check(rc, "mss_mask_edges")
    return edges


def squared_edt(feature: torch.Tensor) -> torch.Tensor:
    """Exact squared distance to the nearest non-zero voxel of the uint8 volume ``feature`` (int32 ``[D, H, W]``; 2**29 when
    there is none).  Pass along W: one ballot scan per row straight from the mask (``mss_edt_row_mask``, no envelope);
    passes along H and D: the lower envelope of parabolas per line with the stack top in registers (``mss_edt_pass``),
    consecutive threads on lines that are adjacent in memory - no transposes.  An int32 ``feature`` is taken as the
    transform's input itself (0 on features, 2**29 elsewhere) and takes three envelope passes."""
    lib = _lib.load()
    stream = torch.cuda.current_stream().cuda_stream
    d, h, w = feature.shape
    t1 = torch.empty((d, h, w), dtype=torch.int32, device=feature.device)
    t2, s, t = torch.empty_like(t1), torch.empty_like(t1), torch.empty_like(t1)
    dims = _lib.I3(d, h, w)
    if feature.dtype == torch.uint8:
        _lib.check(lib.mss_edt_row_mask(feature.contiguous().data_ptr(), t1.data_ptr(), dims, stream), "mss_edt_row_mask")
    else:
        _lib.check(lib.mss_edt_pass(feature.contiguous().data_ptr(), t1.data_ptr(), s.data_ptr(), t.data_ptr(), dims, 2, stream),
                   "mss_edt_pass")
    _lib.check(lib.mss_edt_pass(t1.data_ptr(), t2.data_ptr(), s.data_ptr(), t.data_ptr(), dims, 1, stream), "mss_edt_pass")
    _lib.check(lib.mss_edt_pass(t2.data_ptr(), t1.data_ptr(), s.data_ptr(), t.data_ptr(), dims, 0, stream), "mss_edt_pass")
    return t1


def _np_lerp(a: float, b: float, n: int, q: float) -> float:
    """``numpy.percentile(x, q)`` (method 'linear') from the two order statistics it interpolates: ``a`` = x_sorted[lo],
    ``b`` = x_sorted[hi] with lo = floor((n - 1) q / 100), hi = min(lo + 1, n - 1).  NumPy's own arithmetic (``_lerp``,
    incl. inf - inf = nan)."""
    pos = (n - 1) * (q / 100.0)
    t = pos - np.floor(pos) if 0 <= pos < n - 1 else 0.0
    with np.errstate(invalid="ignore"):
        d = np.float64(b) - np.float64(a)
        return float(np.float64(b) - d * (1.0 - t)) if t >= 0.5 else float(np.float64(a) + d * t)


_INF_KEY = 1 << 29  # squared_edt's "no feature" value


def _directed_async(edges_from: torch.Tensor, edges_to: torch.Tensor, quant: float, out_row: torch.Tensor, scratch: torch.Tensor,
                    n_to_slot: torch.Tensor) -> None:
    """Everything of ``compute_percent_hausdorff_distance(edges_from, edges_to)`` that needs the GPU, enqueued without a host
    sync: the exact squared distance transform of ``edges_to`` and the two order statistics of its values at the voxels of
    ``edges_from`` (``mss_select2``: radix histograms masked by the surface; ranks derived on the device from the surface's
    voxel count).  ``out_row`` (uint64[4]) receives {n_from, d2_lo, d2_hi, d2_max}, ``n_to_slot`` the voxel count of
    ``edges_to``."""
    lib = _lib.load()
    dt = squared_edt(edges_to)
    n_to_slot.copy_(edges_to.sum(dtype=torch.int64))
    rc = lib.mss_select2(dt.data_ptr(), 0, edges_from.data_ptr(), dt.numel(), 1, float(quant), 0, 0, scratch.data_ptr(),
                         out_row.data_ptr(), torch.cuda.current_stream().cuda_stream)
    _lib.check(rc, "mss_select2")


def _directed_value(row: np.ndarray, n_to: int, percentile: Optional[float]) -> float:
    """The host side of one direction, from the numbers the device produced (get_surface_distance's special cases included)."""
    n_from, k_lo, k_hi, k_max = (int(v) for v in row)
    inf = float("inf")
    if n_to == 0:      # `dis = inf * ones; return dis[seg_pred]`: n_from infinities
        n, a, b, mx = n_from, inf, inf, inf
    elif n_from == 0:  # `if not np.any(seg_pred): return dis[seg_gt]`: infinities for the OTHER surface's voxels
        n, a, b, mx = n_to, inf, inf, inf
    else:
        n = n_from
        a, b, mx = (inf if k >= _INF_KEY else float(np.sqrt(np.float64(k))) for k in (k_lo, k_hi, k_max))
    if n == 0:
        return float("nan")  # surface_distance.shape == (0,)
    if not percentile:
        return mx
    return _np_lerp(a, b, n, float(percentile))


_N_STREAMS = int(os.environ.get("MSS_HD_STREAMS", "4"))  # directed distances in flight together (each holds four int32 volumes of its class's box)
_STREAMS: dict = {}


def _side_streams(device: torch.device):
    key = (device.type, device.index if device.index is not None else torch.cuda.current_device())
    if key not in _STREAMS:
        _STREAMS[key] = [torch.cuda.Stream(device=device) for _ in range(_N_STREAMS)]
    return _STREAMS[key]


def class_boxes(pred: torch.Tensor, label: torch.Tensor, n_classes: int) -> np.ndarray:
    """``generate_spatial_bounding_box`` of ``(pred == c) | (label == c)`` for every class in ONE launch and one D2H:
    int32 ``[K, 6]`` = lo (3) then hi (3, exclusive); lo > hi where the class occurs in neither map."""
    lib = _lib.load()
    boxes = torch.tensor([[1 << 30] * 3 + [0] * 3] * n_classes, dtype=torch.int32, device=pred.device)
    rc = lib.mss_class_boxes(pred.data_ptr(), label.data_ptr(), _lib.I3(*pred.shape), int(n_classes), boxes.data_ptr(),
                             torch.cuda.current_stream().cuda_stream)
    _lib.check(rc, "mss_class_boxes")
    return boxes.cpu().numpy()


def hausdorff_distance(pred: torch.Tensor, label: torch.Tensor, n_classes: int, percentile: Optional[float] = 95,
                       include_background: bool = True, directed: bool = False) -> np.ndarray:
    """Per-class (percentile) Hausdorff distance in voxels, float64 ``[K]`` (``[K-1]`` without background): NaN when a
    class is absent from both maps, inf / NaN by NumPy's rules when it is absent from one.

    Device-driven: one launch finds every class's bounding box (the only host sync before the end), then per class the
    surfaces, the two distance transforms and the order statistics are enqueued back to back; the percentile arithmetic
    runs on the host from ONE device-to-host copy of 2 x 5 integers per class."""
    if not (pred.is_cuda and label.is_cuda):
        raise _lib.MssError("hausdorff_distance needs CUDA tensors; there is no CPU fallback")
    p = pred.reshape(pred.shape[-3:]).to(torch.uint8).contiguous()
    y = label.reshape(label.shape[-3:])
    y = (y if y.dtype == torch.uint8 else y.round().clamp(0, 255).to(torch.uint8)).contiguous()
    if p.shape != y.shape:
        raise ValueError(f"pred {tuple(p.shape)} and label {tuple(y.shape)} differ")
    lib = _lib.load()
    classes = list(range(0 if include_background else 1, n_classes))
    quant = (float(percentile) / 100.0) if percentile else 1.0
    with torch.cuda.device(p.device):
        boxes = class_boxes(p, y, n_classes)
        res = torch.zeros((len(classes), 2, 4), dtype=torch.int64, device=p.device)   # uint64 bit patterns (values < 2^63)
        n_to = torch.zeros((len(classes), 2), dtype=torch.int64, device=p.device)
        main = torch.cuda.current_stream()
        streams = _side_streams(p.device)
        n_scr = int(lib.mss_select_scratch_bytes()) // 8 + 1
        scratch = torch.empty((len(streams), n_scr), dtype=torch.int64, device=p.device)  # one select scratch per stream
        present, keep, job = [], [], 0
        for s in streams:
            s.wait_stream(main)  # res / n_to / scratch exist before a side stream touches them
        for i, c in enumerate(classes):
            lo, hi = tuple(int(v) for v in boxes[c, :3]), tuple(int(v) for v in boxes[c, 3:])
            present.append(all(h > l for l, h in zip(lo, hi)))
            if not present[-1]:
                continue
            ep = _edges(p, c, lo, hi)
            ey = _edges(y, c, lo, hi)
            keep.append((ep, ey))  # allocated on the main stream, read on side streams: alive until those have been joined
            ready = torch.cuda.Event()
            ready.record(main)
            for d, (e_from, e_to) in enumerate(((ep, ey), (ey, ep))[: 1 if directed else 2]):
                s = streams[job % len(streams)]
                s.wait_event(ready)
                with torch.cuda.stream(s):
                    _directed_async(e_from, e_to, quant, res[i, d], scratch[job % len(streams)], n_to[i, d])
                job += 1
        for s in streams:
            main.wait_stream(s)
        res_h, n_to_h = res.cpu().numpy(), n_to.cpu().numpy()
        del keep
    out = []
    for i, _c in enumerate(classes):
        if not present[i]:
            out.append(float("nan"))
            continue
        d1 = _directed_value(res_h[i, 0], int(n_to_h[i, 0]), percentile)
        if directed:
            out.append(d1)
            continue
        d2 = _directed_value(res_h[i, 1], int(n_to_h[i, 1]), percentile)
        out.append(max(d1, d2))                   # Python's max, as the reference calls it (NaN handling included)
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
