"""Deterministic stand-in predictors for stitching-parity tests (TEST INFRASTRUCTURE).

The parity criterion of BASELINE.json is stated "given identical per-window
logits".  A real backbone (cuDNN/cuBLAS) does not give bit-identical logits on
CPU and GPU, so the stitching tests use a predictor built ONLY from single
IEEE-754 float32 multiplies and adds issued as separate torch ops (no FMA
contraction, no transcendental) - bit-identical on CPU and CUDA.  It consumes
the reference's 3-tuple ``(patches, centers, affine)`` (engine/utils.py:134) or a
plain patch tensor (stock MONAI inferer, run_evaluation.py:68-74).
"""
from __future__ import annotations

from typing import Any, List, Optional

import torch


class ArithmeticPredictor:
    """``logits[:, k] = (x * a_k + b_k) * (x * c_k + d_k) + e_k * t + f_k`` with
    ``x`` the running channel sum of the patch and ``t`` a per-window scalar made
    from the window centres and the affine diagonal."""

    def __init__(self, n_classes: int, use_centers: bool = True) -> None:
        self.k = n_classes
        self.use_centers = use_centers
        self.calls: List[Any] = []
        # small dyadic-friendly coefficients; any float32 values work
        self.a = [0.75 + 0.125 * k for k in range(n_classes)]
        self.b = [0.5 - 0.0625 * k for k in range(n_classes)]
        self.c = [(-1.0) ** k * (0.3 + 0.05 * k) for k in range(n_classes)]
        self.d = [0.2 * (k % 3) - 0.1 for k in range(n_classes)]
        self.e = [0.37 * ((k * 5) % 7 - 3) for k in range(n_classes)]
        self.f = [0.01 * k for k in range(n_classes)]

    def __call__(self, model_in: Any, *args: Any, **kwargs: Any) -> torch.Tensor:
        centers: Optional[torch.Tensor] = None
        affine: Optional[torch.Tensor] = None
        if isinstance(model_in, (tuple, list)):
            patches, centers, affine = model_in
        else:
            patches = model_in
        self.calls.append((tuple(patches.shape), None if centers is None else tuple(centers.shape)))
        x = patches[:, 0]
        for ch in range(1, patches.shape[1]):
            x = x + patches[:, ch]
        t = None
        if self.use_centers and centers is not None:
            c = centers.reshape(-1, 3).to(torch.float32)
            t = c[:, 0] * 0.5
            t = t + c[:, 1] * 0.25
            t = t + c[:, 2] * 0.125
            if affine is not None:
                t = t + affine.reshape(-1, 3)[0, 0].to(torch.float32) * 0.015625
            t = t.reshape(-1, 1, 1, 1)
        planes = []
        for k in range(self.k):
            u = x * self.a[k]
            u = u + self.b[k]
            v = x * self.c[k]
            v = v + self.d[k]
            w = u * v
            if t is not None:
                w = w + t * self.e[k]
            w = w + self.f[k]
            planes.append(w)
        return torch.stack(planes, dim=1).contiguous()


class PositionalPredictor:
    """Plain-tensor predictor that is NOT flip-equivariant (mirror test-time augmentation needs one): every class
    plane is the channel sum of the patch pushed through one multiply-add and then multiplied by a fixed spatial ramp,
    ``out[:, k] = (x * a_k + b_k) * ramp_k`` - again only single IEEE float32 multiplies / adds issued as separate
    torch ops, so CPU and CUDA agree bit for bit."""

    def __init__(self, n_classes: int, patch_size) -> None:
        self.k = n_classes
        d, h, w = (int(v) for v in patch_size)
        zz, yy, xx = torch.meshgrid(torch.arange(d), torch.arange(h), torch.arange(w), indexing="ij")
        self.ramps = [(1.0 + ((zz * 3 + yy * 5 + xx * 7 + k) % 11).to(torch.float32) * 0.125) for k in range(n_classes)]
        self.a = [0.5 + 0.25 * k for k in range(n_classes)]
        self.b = [0.125 * ((k * 3) % 5) - 0.25 for k in range(n_classes)]

    def __call__(self, x: torch.Tensor, *args, **kwargs) -> torch.Tensor:
        s = x[:, 0]
        for c in range(1, x.shape[1]):
            s = s + x[:, c]
        planes = []
        for k in range(self.k):
            u = s * self.a[k]
            u = u + self.b[k]
            planes.append(u * self.ramps[k].to(x.device))
        return torch.stack(planes, dim=1).contiguous()
