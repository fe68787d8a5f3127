"""CPU restatement of the reference's Dice evaluation (TEST INFRASTRUCTURE).

Follows the MONAI 0.8 pieces the reference wires together at
``/root/reference/engine/test.py:28-31,50-69``: ``AsDiscrete(argmax, to_onehot)``,
``DiceMetric(include_background=True, reduction="none", get_not_nans=True)``
(-> ``compute_meandice``) and the per-class ``nanmean`` bookkeeping.  MONAI is
absent from ``/root/reference`` (un-vendored, un-pinned) - parity for this file is
unpinned by the reference; the exact integer counts are the quantity compared
bit-for-bit, Dice itself is derived from them.
"""
from __future__ import annotations

from typing import Tuple

import numpy as np
import torch


def dice_counts(pred: np.ndarray, label: np.ndarray, n_classes: int) -> np.ndarray:
    """Exact per-class counts behind ``compute_meandice``: row 0 ``TP = #(pred==c & y==c)``,
    row 1 ``P = #(pred==c)``, row 2 ``Y = #(y==c)``; int64 ``[3, K]``.
    Values outside ``[0, K)`` belong to no class (``one_hot`` would reject them)."""
    pred = np.asarray(pred).reshape(-1)
    label = np.asarray(label).reshape(-1)
    out = np.zeros((3, n_classes), dtype=np.int64)
    for c in range(n_classes):
        p = pred == c
        y = label == c
        out[0, c] = np.count_nonzero(p & y)
        out[1, c] = np.count_nonzero(p)
        out[2, c] = np.count_nonzero(y)
    return out


def dice_from_counts(counts: np.ndarray) -> np.ndarray:
    """``2 TP / (Y + P)`` where ``Y > 0`` else NaN (``compute_meandice``), in float64."""
    tp, p, y = (counts[i].astype(np.float64) for i in range(3))
    with np.errstate(divide="ignore", invalid="ignore"):
        d = 2.0 * tp / (y + p)
    return np.where(y > 0, d, np.nan)


def monai_meandice(logits: torch.Tensor, label: torch.Tensor, n_classes: int) -> torch.Tensor:
    """What the reference computes per volume (engine/test.py:50-56): argmax one-hot of the
    prediction, one-hot of the label, ``compute_meandice`` on float32 one-hots -> ``[1, K]`` with NaNs kept.
    ``logits`` is ``[K, D, H, W]``, ``label`` is ``[1, D, H, W]`` (integer valued)."""
    am = torch.argmax(logits, dim=0, keepdim=True)
    y_pred = torch.zeros((n_classes,) + tuple(am.shape[1:]), dtype=torch.float32).scatter_(0, am, 1.0)[None]
    y = torch.zeros_like(y_pred[0]).scatter_(0, label.long(), 1.0)[None]
    axes = list(range(2, y.dim()))
    inter = torch.sum(y * y_pred, dim=axes)
    y_o = torch.sum(y, dim=axes)
    p_o = torch.sum(y_pred, dim=axes)
    return torch.where(y_o > 0, (2.0 * inter) / (y_o + p_o), torch.tensor(float("nan")))


def class_means(dice_scores: np.ndarray) -> Tuple[np.ndarray, float]:
    """engine/test.py:59-69: per class ``nanmean`` over the batch if any entry is not NaN else NaN;
    ``mDice`` = ``nanmean`` over classes.  ``dice_scores`` is ``[B, K]``."""
    dice_scores = np.asarray(dice_scores, dtype=np.float64)
    means = np.full(dice_scores.shape[1], np.nan)
    for c in range(dice_scores.shape[1]):
        col = dice_scores[:, c]
        if np.any(~np.isnan(col)):
            means[c] = np.nanmean(col)
    m = float(np.nanmean(means)) if np.any(~np.isnan(means)) else float("nan")
    return means, m


def eval_meters(per_volume_dice: np.ndarray) -> Tuple[np.ndarray, float]:
    """What ``eval_model`` reports for a SET of volumes (engine/test.py:37-94 with ``utils/misc.py:93-100,16-60``): the loader
    yields one volume per iteration; each iteration computes ``class_means`` of that single volume, feeds every class value
    that is not ``np.nan`` to the meter ``class{c}Dice`` and ``mDice = nanmean`` over that volume's classes to the meter
    ``mDice``; the result is every meter's ``global_avg`` (sum / count).  So ``eval/mDice`` is the mean over volumes of
    the per-volume class mean, NOT the mean over classes of the per-class means.  ``per_volume_dice`` is ``[N, K]``."""
    d = np.asarray(per_volume_dice, dtype=np.float64)
    k = d.shape[1]
    tot, cnt = np.zeros(k), np.zeros(k)
    m_tot, m_cnt = 0.0, 0
    for vol in d:
        means, m = class_means(vol[None])
        for c in range(k):
            if not np.isnan(means[c]):  # `v is np.nan` is skipped by MetricLogger.update (utils/misc.py:96)
                tot[c] += means[c]
                cnt[c] += 1
        m_tot += m          # a tensor .item() NaN is not the np.nan object: it would be added (and poison the mean)
        m_cnt += 1
    with np.errstate(invalid="ignore", divide="ignore"):
        cls = np.where(cnt > 0, tot / np.maximum(cnt, 1), np.nan)
    return cls, (m_tot / m_cnt if m_cnt else float("nan"))
