"""CPU oracle (TEST INFRASTRUCTURE) for the nnU-Net-style tiled predictor of the reference's nnFormer branch
(models/segmentors/nnformer_official/neural_network.py).  Restates, citing the lines it follows:

* ``compute_steps``        :274-298  `_compute_steps_for_sliding_window`
* ``get_gaussian``         :258-271  `_get_gaussian` (scipy.ndimage.gaussian_filter, as the reference calls it)
* ``mirror_and_pred``      :511-568  `_internal_maybe_mirror_and_pred_3D`
* ``predict_3D_tiled``     :300-437  `_internal_predict_3D_3Dconv_tiled`: the float32 (`all_in_gpu=False`) branch and the
                                     half-precision `all_in_gpu=True` branch (:346-372, :399-400, :420-431)

Pinned against outputs of those very methods (tests/golden/nnunet_*.npz; make_golden.py imports the reference module
with stand-ins for its absent third-party imports and runs it on the CPU).  ``pad_nd_image`` (batchgenerators, absent
from /root/reference, unpinned) is restated from its published behaviour: centred constant padding up to the patch size.
"""
from __future__ import annotations

from typing import Callable, List, Optional, Sequence, Tuple

import numpy as np
import torch


def compute_steps(patch_size: Sequence[int], image_size: Sequence[int], step_size: float) -> List[List[int]]:
    target = [i * step_size for i in patch_size]
    num_steps = [int(np.ceil((i - k) / j)) + 1 for i, j, k in zip(image_size, target, patch_size)]
    steps = []
    for dim in range(len(patch_size)):
        max_step_value = image_size[dim] - patch_size[dim]
        actual = max_step_value / (num_steps[dim] - 1) if num_steps[dim] > 1 else 99999999999
        steps.append([int(np.round(actual * i)) for i in range(num_steps[dim])])
    return steps


def get_gaussian(patch_size: Sequence[int], sigma_scale: float = 1.0 / 8) -> np.ndarray:
    from scipy.ndimage import gaussian_filter

    tmp = np.zeros(patch_size)
    tmp[tuple(i // 2 for i in patch_size)] = 1
    g = gaussian_filter(tmp, [i * sigma_scale for i in patch_size], 0, mode="constant", cval=0)
    g = g / np.max(g) * 1
    g = g.astype(np.float32)
    g[g == 0] = np.min(g[g != 0])
    return g


def mirror_and_pred(x: torch.Tensor, network: Callable, nonlin: Callable, num_classes: int, mirror_axes: Sequence[int],
                    do_mirroring: bool = True, mult: Optional[torch.Tensor] = None) -> torch.Tensor:
    result = torch.zeros([1, num_classes] + list(x.shape[2:]), dtype=torch.float)
    if do_mirroring:
        mirror_idx, num_results = 8, 2 ** len(mirror_axes)
    else:
        mirror_idx, num_results = 1, 1
    flips = {0: (), 1: (4,), 2: (3,), 3: (4, 3), 4: (2,), 5: (4, 2), 6: (3, 2), 7: (4, 3, 2)}
    need = {0: (), 1: (2,), 2: (1,), 3: (2, 1), 4: (0,), 5: (0, 2), 6: (0, 1), 7: (0, 1, 2)}
    for m in range(mirror_idx):
        if not all(a in mirror_axes for a in need[m]):
            continue
        if m == 0:
            pred = nonlin(network(x))
            result += 1 / num_results * pred
        else:
            pred = nonlin(network(torch.flip(x, flips[m])))
            result += 1 / num_results * torch.flip(pred, flips[m])
    if mult is not None:
        result[:, :] *= mult
    return result


def pad_to_patch(x: np.ndarray, patch_size: Sequence[int]) -> Tuple[np.ndarray, Tuple[slice, ...]]:
    """batchgenerators pad_nd_image(x, patch_size, 'constant', {'constant_values': 0}, True, None)."""
    old = np.array(x.shape[-3:])
    new = np.maximum(np.array(patch_size), old)
    diff = new - old
    below, above = diff // 2, diff // 2 + diff % 2
    pads = [(0, 0)] + [(int(b), int(a)) for b, a in zip(below, above)]
    res = np.pad(x, pads, "constant", constant_values=0) if diff.any() else x
    slicer = tuple(slice(int(b), int(b + o)) for b, o in zip(below, old))
    return res, slicer


def predict_3D_tiled(x: np.ndarray, network: Callable, nonlin: Callable, num_classes: int, patch_size: Sequence[int],
                     step_size: float = 0.5, do_mirroring: bool = True, mirror_axes: Sequence[int] = (0, 1, 2),
                     use_gaussian: bool = True, all_in_gpu: bool = False) -> Tuple[np.ndarray, np.ndarray]:
    data, slicer = pad_to_patch(x, patch_size)
    steps = compute_steps(patch_size, data.shape[1:], step_size)
    num_tiles = len(steps[0]) * len(steps[1]) * len(steps[2])
    if use_gaussian and num_tiles > 1:
        gmap = get_gaussian(patch_size, 1.0 / 8)
        mult = torch.from_numpy(gmap)
        add_nb = gmap
    else:
        mult = None
        add_nb = np.ones(data.shape[1:], dtype=np.float32)
    if all_in_gpu:
        # :346-372 - the importance map, the aggregated results and the aggregated counts are HALF tensors; every `+=` on
        # them computes in float32 and rounds the sum to half (torch type promotion for in-place ops on a half tensor)
        if mult is not None:
            mult = mult.half()
            mult[mult == 0] = mult[mult != 0].min()
            add_t = mult
        else:
            add_t = torch.ones(data.shape[1:])
        agg_t = torch.zeros([num_classes] + list(data.shape[1:]), dtype=torch.half)
        nb_t = torch.zeros([num_classes] + list(data.shape[1:]), dtype=torch.half)
        data_t = torch.from_numpy(np.ascontiguousarray(data))
        for sx in steps[0]:
            for sy in steps[1]:
                for sz in steps[2]:
                    sl = (slice(None), slice(sx, sx + patch_size[0]), slice(sy, sy + patch_size[1]), slice(sz, sz + patch_size[2]))
                    pred = mirror_and_pred(data_t[sl][None], network, nonlin, num_classes, mirror_axes, do_mirroring, mult)[0]
                    agg_t[sl] += pred.half()   # :399-400, :405
                    nb_t[sl] += add_t          # :406
        agg_t = agg_t[(slice(None),) + slicer]
        nb_t = nb_t[(slice(None),) + slicer]
        probs_t = agg_t / nb_t                 # half / half -> half (:420)
        return probs_t.argmax(0).numpy(), probs_t.numpy()
    agg = np.zeros([num_classes] + list(data.shape[1:]), dtype=np.float32)
    nb = np.zeros([num_classes] + list(data.shape[1:]), dtype=np.float32)
    for sx in steps[0]:
        for sy in steps[1]:
            for sz in steps[2]:
                sl = (slice(None), slice(sx, sx + patch_size[0]), slice(sy, sy + patch_size[1]), slice(sz, sz + patch_size[2]))
                patch = torch.from_numpy(np.ascontiguousarray(data[sl][None])).float()
                pred = mirror_and_pred(patch, network, nonlin, num_classes, mirror_axes, do_mirroring, mult)[0].numpy()
                agg[sl] += pred
                nb[sl] += add_nb
    agg = agg[(slice(None),) + slicer]
    nb = nb[(slice(None),) + slicer]
    probs = agg / nb
    return probs.argmax(0), probs
