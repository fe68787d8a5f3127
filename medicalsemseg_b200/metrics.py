"""Per-class Dice from exact confusion counts: what engine/test.py:28-31,50-69 gets out of MONAI's
AsDiscrete + DiceMetric(include_background=True, reduction="none", get_not_nans=True)."""
from __future__ import annotations

from typing import Any, List, Optional, Tuple

import numpy as np
import torch

from . import _lib


def dice_counts(pred: torch.Tensor, label: torch.Tensor, n_classes: int,
                out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """``int64[3, K]`` on the GPU: row 0 ``#(pred==c & label==c)``, row 1 ``#(pred==c)``, row 2 ``#(label==c)``.

    ``pred`` is a uint8 label map (any shape), ``label`` uint8 or integer-valued float32 of the same number of
    voxels (the reference's loaders give float32 ``[1, 1, D, H, W]``).  Counts are ADDED into ``out`` when given.
    """
    if not (pred.is_cuda and label.is_cuda):
        raise _lib.MssError("dice_counts needs CUDA tensors; there is no CPU fallback")
    if n_classes > _lib.MAX_DICE_CLASSES:
        raise _lib.MssError(f"dice_counts supports up to {_lib.MAX_DICE_CLASSES} classes")
    pred = pred.contiguous()
    if pred.dtype != torch.uint8:
        pred = pred.to(torch.uint8)
    if label.dtype == torch.uint8:
        ldt = 0
    else:
        label = label.to(torch.float32)
        ldt = 1
    label = label.contiguous()
    if pred.numel() != label.numel():
        raise ValueError(f"pred has {pred.numel()} voxels, label {label.numel()}")
    if out is None:
        out = torch.zeros((3, n_classes), dtype=torch.int64, device=pred.device)
    with torch.cuda.device(pred.device):
        rc = _lib.load().mss_dice_counts(pred.data_ptr(), label.data_ptr(), ldt, pred.numel(), int(n_classes),
                                         out.data_ptr(), torch.cuda.current_stream().cuda_stream)
    _lib.check(rc, "mss_dice_counts")
    return out


def dice_counts_batched(pred: torch.Tensor, label: torch.Tensor, n_classes: int) -> torch.Tensor:
    """Per-volume counts ``int64[B, 3, K]`` of a batch of label maps ``[B, ...]`` in ONE launch (cfg5: a rank's share of
    the evaluation set); ``label`` uint8 or integer-valued float32 of the same shape."""
    if not (pred.is_cuda and label.is_cuda):
        raise _lib.MssError("dice_counts needs CUDA tensors; there is no CPU fallback")
    if n_classes > _lib.MAX_DICE_CLASSES:
        raise _lib.MssError(f"dice_counts supports up to {_lib.MAX_DICE_CLASSES} classes")
    b = pred.shape[0]
    pred = pred.to(torch.uint8).reshape(b, -1).contiguous()
    if label.dtype == torch.uint8:
        ldt = 0
    else:
        label, ldt = label.to(torch.float32), 1
    label = label.reshape(b, -1).contiguous()
    if pred.shape != label.shape:
        raise ValueError(f"pred {tuple(pred.shape)} and label {tuple(label.shape)} differ")
    out = torch.zeros((b, 3, n_classes), dtype=torch.int64, device=pred.device)
    with torch.cuda.device(pred.device):
        rc = _lib.load().mss_dice_counts_batched(pred.data_ptr(), label.data_ptr(), ldt, pred.shape[1], b, int(n_classes),
                                                 out.data_ptr(), torch.cuda.current_stream().cuda_stream)
    _lib.check(rc, "mss_dice_counts_batched")
    return out


def dice_from_counts(counts: Any) -> np.ndarray:
    """``2 TP / (Y + P)`` where ``Y > 0`` else NaN (MONAI compute_meandice), float64, per class (last axis)."""
    c = counts.detach().cpu().numpy() if isinstance(counts, torch.Tensor) else np.asarray(counts)
    tp, p, y = (c[..., i, :].astype(np.float64) for i in range(3))
    with np.errstate(divide="ignore", invalid="ignore"):
        d = 2.0 * tp / (y + p)
    return np.where(y > 0, d, np.nan)


class DiceMeter:
    """Per-volume Dice bookkeeping of ``eval_model`` (engine/test.py:37-94): the loader yields one volume per iteration,
    each iteration feeds its per-class Dice (NaN classes skipped, utils/misc.py:96) and ``mDice = nanmean`` over that
    volume's classes into ``MetricLogger`` meters, and the reported values are the meters' global averages.  Hence

    * class mean = mean over the volumes in which the class occurs (NaN if it never does);
    * ``mDice``  = mean over VOLUMES of the per-volume nanmean over classes - not the nanmean of the class means
      (A = (1.0, 0.5), B = (0.8, nan) gives 0.775, not 0.70)."""

    def __init__(self, n_classes: int) -> None:
        self.k = n_classes
        self.per_volume: List[np.ndarray] = []

    def update(self, pred: torch.Tensor, label: torch.Tensor) -> torch.Tensor:
        counts = dice_counts(pred, label, self.k)
        self.per_volume.append(counts.cpu().numpy())
        return counts

    def update_batch(self, pred: torch.Tensor, label: torch.Tensor) -> torch.Tensor:
        """A batch of volumes ``[B, ...]`` in one launch (``mss_dice_counts_batched``); one Dice vector per volume."""
        counts = dice_counts_batched(pred, label, self.k)
        self.per_volume.extend(list(counts.cpu().numpy()))
        return counts

    def add_counts(self, counts: Any) -> None:
        c = counts.detach().cpu().numpy() if isinstance(counts, torch.Tensor) else np.asarray(counts)
        self.per_volume.extend(list(c.reshape(-1, 3, self.k)))

    def per_volume_dice(self) -> np.ndarray:
        return np.stack([dice_from_counts(c) for c in self.per_volume]) if self.per_volume else np.full((0, self.k), np.nan)

    def class_means(self) -> Tuple[np.ndarray, float]:
        """``(eval/class{c}Dice for every c, eval/mDice)`` as ``eval_model`` would log them for the volumes seen so far."""
        d = self.per_volume_dice()
        means = np.full(self.k, np.nan)
        if d.shape[0] == 0:
            return means, float("nan")
        seen = ~np.isnan(d)
        for c in range(self.k):
            if seen[:, c].any():
                means[c] = d[seen[:, c], c].mean()
        # per-volume mDice (engine/test.py:70); a volume with no class at all contributes NaN, as the reference's meter would
        per_vol = np.array([v[s].mean() if s.any() else np.nan for v, s in zip(d, seen)])
        return means, float(per_vol.mean())

    def mean_of_class_means(self) -> float:
        """nanmean over classes of the class means (NOT what the reference logs; kept for set-level summaries)."""
        means, _ = self.class_means()
        return float(np.nanmean(means)) if np.any(~np.isnan(means)) else float("nan")


def gather_volume_counts(counts: Any, group: Any = None) -> np.ndarray:
    """Per-volume Dice counts of ALL ranks, ``int64 [N_total, 3, K]`` in rank order (volumes sharded over the GPUs,
    BASELINE.json configs[4]).  The reference averages Dice per volume (``engine/test.py:59-69``: nan-mean over the
    batch per class), so the counts travel per volume - 3K int64 each - instead of being summed.  Ranks may hold
    different numbers of volumes.  Works on whatever device ``counts`` lives on (NCCL for CUDA, gloo for CPU)."""
    import torch.distributed as dist

    c = counts if isinstance(counts, torch.Tensor) else torch.as_tensor(np.asarray(counts))
    c = c.reshape(-1, 3, c.shape[-1]).to(torch.int64).contiguous()
    world = dist.get_world_size(group)
    n = torch.tensor([c.shape[0]], dtype=torch.int64, device=c.device)
    ns = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(ns, n, group=group)
    n_max = max(int(x.item()) for x in ns)
    pad = torch.zeros((n_max, 3, c.shape[-1]), dtype=torch.int64, device=c.device)
    pad[: c.shape[0]] = c
    bufs = [torch.zeros_like(pad) for _ in range(world)]
    dist.all_gather(bufs, pad, group=group)
    return np.concatenate([b[: int(k.item())].cpu().numpy() for b, k in zip(bufs, ns)], axis=0)
