"""Nearest-neighbour resampling of label maps back to the original grid on the GPU: the arithmetic of
utils/misc.py:420-425 (``resample_3d`` = ``scipy.ndimage.zoom(order=0, prefilter=False)``), the step right after the
argmax on the reference's test path (engine/test.py:143-147)."""
from __future__ import annotations

from typing import Any, Dict, Sequence, Tuple

import numpy as np
import torch

from . import _lib

_TABLES: Dict[Tuple[int, int, str], torch.Tensor] = {}


def zoom_index_table(n_in: int, n_out: int) -> np.ndarray:
    """Source index per output index along one axis (scipy zoom order 0; -1 = scipy writes the constant 0)."""
    out = np.empty(int(n_out), dtype=np.int32)
    _lib.check(_lib.load().mss_zoom_index_table(int(n_in), int(n_out), out.ctypes.data), "mss_zoom_index_table")
    return out


def _device_table(n_in: int, n_out: int, device: torch.device) -> torch.Tensor:
    key = (n_in, n_out, str(device))
    t = _TABLES.get(key)
    if t is None:
        if len(_TABLES) > 256:
            _TABLES.clear()
        t = torch.from_numpy(zoom_index_table(n_in, n_out)).to(device)
        _TABLES[key] = t
    return t


def zoomed_shape(shape: Sequence[int], target_size: Sequence[int]) -> Tuple[int, ...]:
    """Output shape exactly as the reference gets it: ``round(n * (float(t) / float(n)))`` per axis
    (utils/misc.py:423 builds the ratio, scipy rounds the product)."""
    return tuple(int(round(int(n) * (float(t) / float(n)))) for n, t in zip(shape, target_size))


def resample_3d(img: Any, target_size: Sequence[int]) -> torch.Tensor:
    """Call-compatible with utils/misc.py:420-425: ``img`` is a uint8 label map ``[X, Y, Z]`` (or a batch
    ``[N, X, Y, Z]``), the result the nearest-neighbour zoom to ``target_size`` as a CUDA uint8 tensor."""
    if not torch.cuda.is_available():
        raise _lib.MssError("medicalsemseg_b200 needs a CUDA device (B200); there is no CPU fallback")
    t = torch.as_tensor(img)
    if t.dtype != torch.uint8:
        raise ValueError("resample_3d expects a uint8 label map (engine/test.py:141 casts before resampling)")
    if not t.is_cuda:
        t = t.cuda()
    batched = t.dim() == 4
    if t.dim() not in (3, 4):
        raise ValueError("resample_3d expects [X, Y, Z] or [N, X, Y, Z]")
    if len(target_size) != 3:
        raise ValueError("target_size must have 3 entries")
    t = t.contiguous()
    nb = t.shape[0] if batched else 1
    in_dims = tuple(int(v) for v in t.shape[-3:])
    out_dims = zoomed_shape(in_dims, target_size)
    if min(out_dims) < 1:
        raise ValueError(f"target_size {tuple(target_size)} gives an empty output")
    out = torch.empty(((nb,) if batched else ()) + out_dims, dtype=torch.uint8, device=t.device)
    with torch.cuda.device(t.device):
        tabs = [_device_table(in_dims[a], out_dims[a], t.device) for a in range(3)]
        rc = _lib.load().mss_resample_nearest(t.data_ptr(), _lib.I3(*in_dims), out.data_ptr(), _lib.I3(*out_dims), nb,
                                              tabs[0].data_ptr(), tabs[1].data_ptr(), tabs[2].data_ptr(),
                                              torch.cuda.current_stream().cuda_stream)
    _lib.check(rc, "mss_resample_nearest")
    return out
