"""nnU-Net-style tiled prediction on the GPU: the stitching policy of the reference's nnFormer branch
(models/segmentors/nnformer_official/neural_network.py:257-437 `_internal_predict_3D_3Dconv_tiled`, :511-568
`_internal_maybe_mirror_and_pred_3D`) behind the same kernels as the MONAI-style path.

What differs from engine/utils.py is host-side policy only:
* window starts: ``num_steps = ceil((image - patch) / (patch * step_size)) + 1`` evenly spread, ``np.round``-ed
  (:274-298) instead of a fixed interval with a clamped last window;
* importance map: ``scipy.ndimage.gaussian_filter`` of a centred delta, sigma = patch / 8, divided by its maximum,
  zeros lifted to the smallest non-zero value (:258-271).  The filter of a delta is the outer product of the
  three normalised 1-D kernels, evaluated here in float64 in scipy's association order - no scipy needed;
* the predictor returns softmax probabilities averaged over up to 8 mirrored forward passes (:511-568): the mirrored
  input copies and the un-mirroring + averaging of the 8 outputs are ``mss_flip_copy`` / ``mss_mirror_merge``.
Accumulation (``aggregated_results += pred * gaussian``, ``aggregated_nb_of_predictions += gaussian``), the division
and the argmax are ``mss_accumulate`` / ``mss_finalize_labels`` unchanged.
"""
from __future__ import annotations

import ctypes as C
from typing import Any, Callable, List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import _lib
from .grid import make_grid
from .inferer import _DTYPES, InferStats, Stitcher, StitchPlan, _crop, _tma_ready, labels_from_logits  # noqa: F401


def compute_steps_for_sliding_window(patch_size: Sequence[int], image_size: Sequence[int], step_size: float) -> List[List[int]]:
    """neural_network.py:274-298, statement for statement (Python floats, ``np.round`` = round half to even)."""
    assert all(i >= j for i, j in zip(image_size, patch_size)), "image size must be as large or larger than patch_size"
    assert 0 < step_size <= 1, "step_size must be larger than 0 and smaller or equal to 1"
    target = [i * step_size for i in patch_size]
    num_steps = [int(np.ceil((i - k) / j)) + 1 for i, j, k in zip(image_size, target, patch_size)]
    steps = []
    for dim in range(len(patch_size)):
        max_step_value = image_size[dim] - patch_size[dim]
        actual = max_step_value / (num_steps[dim] - 1) if num_steps[dim] > 1 else 99999999999
        steps.append([int(np.round(actual * i)) for i in range(num_steps[dim])])
    return steps


def _gaussian_profile(n: int, sigma: float) -> np.ndarray:
    """scipy.ndimage.gaussian_filter1d (truncate 4, mode constant) applied to a delta at n // 2, float64."""
    radius = int(4.0 * float(sigma) + 0.5)
    x = np.arange(-radius, radius + 1)
    phi = np.exp(-0.5 / (sigma * sigma) * x ** 2)
    phi = phi / phi.sum()
    out = np.zeros(n, dtype=np.float64)
    c = n // 2
    for i in range(n):
        j = i - c + radius
        if 0 <= j < phi.shape[0]:
            out[i] = phi[j]
    return out


def gaussian_importance_map(patch_size: Sequence[int], sigma_scale: float = 1.0 / 8) -> np.ndarray:
    """neural_network.py:258-271 (`_get_gaussian`) without scipy: bit-identical float32 map."""
    prof = [_gaussian_profile(int(n), n * sigma_scale) for n in patch_size]
    g = (prof[0][:, None, None] * prof[1][None, :, None]) * prof[2][None, None, :]  # the order the three 1-D passes multiply
    g = g / np.max(g) * 1
    g = g.astype(np.float32)
    g[g == 0] = np.min(g[g != 0])
    return g


def _mirror_masks(mirror_axes: Sequence[int], do_mirroring: bool) -> List[int]:
    """The passes neural_network.py:535-565 runs, as bit masks (bit 0 = W / axis 2, bit 1 = H / axis 1, bit 2 = D / axis 0)."""
    if not do_mirroring:
        return [0]
    need = {0: (), 1: (2,), 2: (1,), 3: (2, 1), 4: (0,), 5: (0, 2), 6: (0, 1), 7: (0, 1, 2)}
    return [m for m in range(8) if all(a in mirror_axes for a in need[m])]


class MirrorTTA:
    """``_internal_maybe_mirror_and_pred_3D`` as a predictor: ``result = sum_m (1 / n) * unflip_m(nonlin(net(flip_m(x))))``
    with ``n = 2 ** len(mirror_axes)`` (the reference's divisor, :531), accumulated in the reference's order."""

    def __init__(self, network: Callable[[torch.Tensor], torch.Tensor], mirror_axes: Sequence[int] = (0, 1, 2),
                 do_mirroring: bool = True, nonlin: Optional[Callable[[torch.Tensor], torch.Tensor]] = None) -> None:
        self.network = network
        self.masks = _mirror_masks(tuple(mirror_axes), do_mirroring)
        self.scale = 1.0 / (2 ** len(tuple(mirror_axes))) if do_mirroring else 1.0
        self.nonlin = nonlin if nonlin is not None else (lambda t: torch.softmax(t, 1))
        self.gpu_launches = 0

    def __call__(self, x: torch.Tensor, *args: Any, **kwargs: Any) -> torch.Tensor:
        lib = _lib.load()
        x = x.contiguous()
        stream = torch.cuda.current_stream().cuda_stream
        dims = _lib.I3(*x.shape[2:])
        preds = []
        for m in self.masks:
            if m == 0:
                xin = x
            else:
                xin = torch.empty_like(x)
                _lib.check(lib.mss_flip_copy(x.data_ptr(), xin.data_ptr(), x.shape[0] * x.shape[1], dims, m, stream),
                           "mss_flip_copy")
                self.gpu_launches += 1
            preds.append(self.nonlin(self.network(xin, *args, **kwargs)).to(torch.float32).contiguous())
        out = torch.empty_like(preds[0])
        ptrs = (C.c_void_p * len(preds))(*[t.data_ptr() for t in preds])
        masks = (C.c_int32 * len(preds))(*self.masks)
        _lib.check(lib.mss_mirror_merge(ptrs, masks, len(preds), float(self.scale), out.data_ptr(),
                                        out.shape[0] * out.shape[1], _lib.I3(*out.shape[2:]), stream), "mss_mirror_merge")
        self.gpu_launches += 1
        return out


def predict_3D_tiled(x: torch.Tensor, network: Callable[[torch.Tensor], torch.Tensor], patch_size: Sequence[int],
                     step_size: float = 0.5, do_mirroring: bool = True, mirror_axes: Sequence[int] = (0, 1, 2),
                     use_gaussian: bool = True, nonlin: Optional[Callable[[torch.Tensor], torch.Tensor]] = None,
                     sw_batch_size: int = 1, stats: Optional[InferStats] = None,
                     group_bytes: Optional[int] = None, all_in_gpu: bool = False) -> Tuple[torch.Tensor, torch.Tensor]:
    """``_internal_predict_3D_3Dconv_tiled`` (neural_network.py:300-437): ``x`` is ``[C, X, Y, Z]``; returns
    ``(segmentation uint8 [X, Y, Z], class_probabilities [K, X, Y, Z])`` as CUDA tensors - fp32 probabilities for the
    default branch, float16 for ``all_in_gpu=True`` (:346-372, :399-406, :420-423: half importance map, half aggregated
    results and counts, half division; ``mss_accumulate_half`` reproduces every rounding).

    ``sw_batch_size`` windows share one backbone call (the reference runs one; results are identical per window)."""
    if not torch.cuda.is_available():
        raise _lib.MssError("medicalsemseg_b200 needs a CUDA device (B200); there is no CPU fallback")
    if x.dim() != 4:
        raise AssertionError("x must be (c, x, y, z)")
    vol = x.to(device="cuda" if not x.is_cuda else x.device, dtype=torch.float32).unsqueeze(0).contiguous()
    dev = vol.device
    patch = tuple(int(v) for v in patch_size)
    with torch.cuda.device(dev):
        spatial = tuple(vol.shape[2:])
        image = tuple(max(s, r) for s, r in zip(spatial, patch))  # pad_nd_image: at least the patch, centred, zeros
        steps = compute_steps_for_sliding_window(patch, image, step_size)
        grid = make_grid(spatial, patch, 0.0, starts=steps)
        num_tiles = grid.n_windows
        plan = StitchPlan(grid, dev, 1)
        if use_gaussian and num_tiles > 1:
            imp = torch.from_numpy(gaussian_importance_map(patch, 1.0 / 8)).to(dev)
        else:
            imp = torch.ones(patch, dtype=torch.float32, device=dev)
        predictor = MirrorTTA(network, mirror_axes, do_mirroring, nonlin)
        if all_in_gpu:
            return _predict_half(vol, grid, plan, imp, predictor, use_gaussian and num_tiles > 1, sw_batch_size, stats)
        st = Stitcher(plan, imp, fuse=_lib.FUSE_LOGITS, sw_batch=sw_batch_size, group_bytes=group_bytes, stats=stats)
        if stats is not None:
            stats.n_windows = st.total
            stats._near_ties = st.near
        vol_r = _tma_ready(vol, grid, 0.0)
        for _first, n, patches, _centers in st.batches(vol_r, 0.0, grid.pad_lo):
            st.push(predictor(patches), n)
            if stats is not None:
                stats.n_predictor_calls += 1
        st.flush()
        labels = labels_from_logits(st.acc, st, normalise=False)
        if stats is not None:
            stats.gpu_launches += predictor.gpu_launches
    return _crop(labels, grid)[0], _crop(st.acc, grid)[0]


def _predict_half(vol: torch.Tensor, grid: Any, plan: StitchPlan, imp: torch.Tensor, predictor: MirrorTTA, gaussian: bool,
                  sw_batch_size: int, stats: Optional[InferStats]) -> Tuple[torch.Tensor, torch.Tensor]:
    """The ``all_in_gpu`` branch: every tile's prediction is kept (fp32, as the mirror merge leaves it) and ONE launch of
    ``mss_accumulate_half`` walks them per voxel in tile order with the reference's binary16 roundings."""
    lib = _lib.load()
    dev = vol.device
    if gaussian:
        imp_h = imp.half()  # neural_network.py:352-356
        imp_h[imp_h == 0] = imp_h[imp_h != 0].min()
        imp = imp_h.float().contiguous()
    n_batches = -(-plan.n_local // sw_batch_size)
    if n_batches > _lib.MAX_BATCH_PTRS:
        sw_batch_size = -(-plan.n_local // _lib.MAX_BATCH_PTRS)
    st = Stitcher(plan, imp, fuse=_lib.FUSE_LOGITS, sw_batch=sw_batch_size, stats=stats)
    if stats is not None:
        stats.n_windows = st.total
    vol_r = _tma_ready(vol, grid, 0.0)
    preds: List[torch.Tensor] = []
    for _first, _n, patches, _centers in st.batches(vol_r, 0.0, grid.pad_lo):
        preds.append(predictor(patches).to(torch.float32).contiguous())
        if stats is not None:
            stats.n_predictor_calls += 1
    k = int(preds[0].shape[1])
    lay = plan.layout(k)
    ext = plan.extent
    probs = torch.empty((1, k, ext[0], ext[1], plan.pitch_w), dtype=torch.float32, device=dev)
    labels = torch.empty((1,) + tuple(ext), dtype=torch.uint8, device=dev)
    ptrs = (C.c_void_p * len(preds))(*[t.data_ptr() for t in preds])
    rc = lib.mss_accumulate_half(C.byref(lay), ptrs, len(preds), sw_batch_size, imp.data_ptr(), probs.data_ptr(),
                                 labels.data_ptr(), ext[2], torch.cuda.current_stream().cuda_stream)
    _lib.check(rc, "mss_accumulate_half")
    if stats is not None:
        stats.gpu_launches += 1 + predictor.gpu_launches
    return _crop(labels, grid)[0], _crop(probs, grid)[0].to(torch.float16)
