"""The evaluation loss of the reference on the GPU: ``DiceCELoss(to_onehot_y=True, softmax=True, squared_pred=True)``
(run_evaluation.py:53) applied to the stitched logits at engine/test.py:48 - there on ``outputs.cpu(), labels.cpu()``, here in
one pass over the logits where they already are (``mss_dice_ce_sums``); the 3K + 1 sums become the loss in float64."""
from __future__ import annotations

from typing import Dict, Tuple

import numpy as np
import torch

from . import _lib


def dice_ce_sums(logits: torch.Tensor, labels: torch.Tensor, squared_pred: bool = True) -> np.ndarray:
    """float64 ``[B, 3K + 1]``: per volume ``I[K]``, ``P[K]``, ``G[K]`` and the summed cross-entropy."""
    if not (logits.is_cuda and labels.is_cuda):
        raise _lib.MssError("dice_ce_loss needs CUDA tensors; there is no CPU fallback")
    if logits.dim() != 5:
        raise ValueError("logits must be [B, K, D, H, W]")
    b, k, d, h, w = logits.shape
    if k > 16:
        raise _lib.MssError("dice_ce_loss supports up to 16 classes")
    lab = labels.reshape(b, d, h, w)
    if lab.dtype == torch.uint8:
        ldt = 0
    else:
        lab, ldt = lab.to(torch.float32), 1
    lab = lab.contiguous()
    x = logits.to(torch.float32)
    # rows (b, d, h) must be uniformly strided: true for the stitcher's W-pitched accumulator views, else copy
    if not (x.stride(4) == 1 and x.stride(2) == x.stride(3) * h):
        x = x.contiguous()
    sums = torch.zeros((b, 3 * k + 1), dtype=torch.float64, device=logits.device)
    lib = _lib.load()
    with torch.cuda.device(logits.device):
        for i in range(b):
            rc = lib.mss_dice_ce_sums(x[i].data_ptr(), x.stride(1), x.stride(3), d * h, w, k, lab[i].data_ptr(), ldt,
                                      1 if squared_pred else 0, sums[i].data_ptr(), torch.cuda.current_stream().cuda_stream)
            _lib.check(rc, "mss_dice_ce_sums")
    return sums.cpu().numpy()


def dice_ce_loss(logits: torch.Tensor, labels: torch.Tensor, squared_pred: bool = True, include_background: bool = True,
                 smooth_nr: float = 1e-5, smooth_dr: float = 1e-5, lambda_dice: float = 1.0, lambda_ce: float = 1.0
                 ) -> Tuple[float, Dict[str, float]]:
    """``DiceCELoss(to_onehot_y=True, softmax=True, squared_pred=..., smooth_nr, smooth_dr)(logits, labels)``:
    returns ``(loss, {"dice": ..., "ce": ...})``; MONAI's reductions (Dice: mean over batch and classes, CE: mean over voxels)."""
    s = dice_ce_sums(logits, labels, squared_pred)
    b, k = logits.shape[0], logits.shape[1]
    inter, pred, ground, ce = s[:, :k], s[:, k:2 * k], s[:, 2 * k:3 * k], s[:, 3 * k]
    if not include_background:
        inter, pred, ground = inter[:, 1:], pred[:, 1:], ground[:, 1:]
    f = 1.0 - (2.0 * inter + smooth_nr) / (ground + pred + smooth_dr)
    dice = float(f.mean())
    n_vox = logits.shape[2] * logits.shape[3] * logits.shape[4]
    ce_mean = float(ce.sum() / (b * n_vox))
    return lambda_dice * dice + lambda_ce * ce_mean, {"dice": dice, "ce": ce_mean}
