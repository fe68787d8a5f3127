"""The reference's own op sequence for the hot path, run on the GPU: what this repo's kernels replace on the reference's
real deployment target ("the bar is the ATen op sequence").  A benchmark baseline, NOT the product
and NOT the parity oracle (oracle/ is the CPU checker; this file never imports it).

Restated from reading the reference, statement by statement:
  engine/utils.py:120-125   window batches in C order, ragged last batch
  engine/utils.py:133       torch.cat of B strided slices                       -> patch batch
  engine/utils.py:135       predictor(...)                                      (plain-tensor convention)
  engine/utils.py:137-143   zeros output_image and a K-channel-replicated count_map on first use
  engine/utils.py:146-148   per window: output_image[idx] += importance_map * seg_prob[i]; count_map[idx] += importance_map
  engine/utils.py:151       output_image / count_map
  engine/test.py:140-141    torch.softmax(outputs, 1).cpu().numpy() -> np.argmax(axis=1).astype(uint8)[0]
The window grid and the importance map come from this repo's host code (identical to MONAI 0.8's by the parity tests);
neither is inside the reference's per-volume cost anyway.
"""
from __future__ import annotations

from typing import Any, Callable

import numpy as np
import torch

from medicalsemseg_b200.grid import make_grid
from medicalsemseg_b200.importance import importance_map


def aten_sliding_window_labels(host_volume: torch.Tensor, predictor: Callable[..., torch.Tensor], roi: Any, sw_batch_size: int,
                               overlap: float, device: torch.device) -> np.ndarray:
    inputs = host_volume.to(device)                                   # engine/test.py:116
    nb = inputs.shape[0]
    grid = make_grid(tuple(inputs.shape[2:]), roi, overlap)
    roi3 = grid.roi
    imp = importance_map(roi3, "gaussian", 0.125, device)
    n = grid.n_windows
    slices = []
    for i in range(n):
        s = grid.window_start(i)
        slices.append(tuple(slice(s[a], s[a] + roi3[a]) for a in range(3)))
    total = n * nb
    output_image = count_map = None
    with torch.no_grad():
        for g in range(0, total, sw_batch_size):                      # :120
            rng = range(g, min(g + sw_batch_size, total))
            unravel = [(slice(idx // n, idx // n + 1), slice(None)) + slices[idx % n] for idx in rng]   # :122-125
            window_data = torch.cat([inputs[w] for w in unravel])     # :133
            seg_prob = predictor(window_data)                         # :135
            if output_image is None:                                  # :137-143
                shape = [nb, seg_prob.shape[1]] + list(grid.image_size)
                output_image = torch.zeros(shape, dtype=torch.float32, device=device)
                count_map = torch.zeros(shape, dtype=torch.float32, device=device)
            for j, w in zip(rng, unravel):                            # :146-148
                output_image[w] += imp * seg_prob[j - g]
                count_map[w] += imp
        output_image = output_image / count_map                       # :151
        probs = torch.softmax(output_image, 1).cpu().numpy()          # engine/test.py:140
    return np.argmax(probs, axis=1).astype(np.uint8)[0]               # engine/test.py:141
