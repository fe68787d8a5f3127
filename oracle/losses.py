"""CPU oracle (TEST INFRASTRUCTURE) for the evaluation loss at engine/test.py:48: restates MONAI 0.8.1
``DiceCELoss(to_onehot_y=True, softmax=True, squared_pred=True)`` (run_evaluation.py:53) with the torch ops MONAI issues
(``monai/losses/dice.py``: DiceLoss.forward + CrossEntropyLoss).  MONAI is absent from /root/reference and un-pinned:
PARITY UNPINNED, anchored on the call site."""
from __future__ import annotations

from typing import Dict, Tuple

import torch
import torch.nn.functional as F


def dice_ce_loss(logits: torch.Tensor, labels: torch.Tensor, squared_pred: bool = True, include_background: bool = True,
                 smooth_nr: float = 1e-5, smooth_dr: float = 1e-5, lambda_dice: float = 1.0, lambda_ce: float = 1.0
                 ) -> Tuple[float, Dict[str, float]]:
    n_pred_ch = logits.shape[1]
    inp = torch.softmax(logits, 1)
    target = F.one_hot(labels.reshape(labels.shape[0], *labels.shape[-3:]).long(), n_pred_ch).permute(0, 4, 1, 2, 3).to(inp.dtype)
    if not include_background:
        inp, target = inp[:, 1:], target[:, 1:]
    reduce_axis = [2, 3, 4]
    intersection = torch.sum(target * inp, dim=reduce_axis)
    if squared_pred:
        target = torch.pow(target, 2)
        inp = torch.pow(inp, 2)
    ground_o = torch.sum(target, dim=reduce_axis)
    pred_o = torch.sum(inp, dim=reduce_axis)
    f = 1.0 - (2.0 * intersection + smooth_nr) / (ground_o + pred_o + smooth_dr)
    dice = torch.mean(f)
    ce = F.cross_entropy(logits, labels.reshape(labels.shape[0], *labels.shape[-3:]).long())
    return float(lambda_dice * dice + lambda_ce * ce), {"dice": float(dice), "ce": float(ce)}
