"""Predictor callbacks for the benchmarks (NOT part of the product path).

BASELINE.json keeps the backbone as the reference's own PyTorch module; /root/reference does not travel to the GPU
box and must not be copied, so these are from-scratch PyTorch modules with the same architecture, layer sizes and
call convention as the reference's wired models, built only from stock torch.nn layers:

* ``SwinUNETRStyle``  - encoder after models/backbones/swin_nnformer.py:478-659 (conv patch embed, 4 stages of
  windowed / shifted-window attention with relative position bias, strided-conv patch merging after EVERY stage)
  under the UNETR-style residual conv decoder of models/segmentors/swin_unetr.py:20-147; hidden 48, depths
  2-2-2-2, heads 3-6-12-24, patch 2, window 6.  Takes the reference's 3-tuple
  ``(patches, centers, affine)`` (swin_nnformer.py:612) or a plain tensor.
* ``PlainUNet``       - the "UNet" of cfg1 from Conv+BN+ReLU / deconv blocks in the spirit of
  models/segmentors/unetr.py:9-52 (quirk Q11), plain-tensor input.

Weights are random-init under torch.manual_seed(13) (the reference's default seed, utils/arguments.py:301).
"""
from __future__ import annotations

from typing import Any, List, Sequence, Tuple

import torch
import torch.nn as nn
import torch.nn.functional as F


def _unpack(model_in: Any) -> torch.Tensor:
    return model_in[0] if isinstance(model_in, (tuple, list)) else model_in


# ---- UNETR-style residual conv blocks (MONAI UnetResBlock / UnetrBasicBlock / UnetrUpBlock semantics) ----------

class ResBlock(nn.Module):
    def __init__(self, cin: int, cout: int) -> None:
        super().__init__()
        self.conv1 = nn.Conv3d(cin, cout, 3, padding=1, bias=False)
        self.norm1 = nn.InstanceNorm3d(cout)
        self.conv2 = nn.Conv3d(cout, cout, 3, padding=1, bias=False)
        self.norm2 = nn.InstanceNorm3d(cout)
        self.skip = None
        if cin != cout:
            self.skip = nn.Sequential(nn.Conv3d(cin, cout, 1, bias=False), nn.InstanceNorm3d(cout))

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        y = F.leaky_relu(self.norm1(self.conv1(x)), 0.01)
        y = self.norm2(self.conv2(y))
        r = x if self.skip is None else self.skip(x)
        return F.leaky_relu(y + r, 0.01)


class UpBlock(nn.Module):
    def __init__(self, cin: int, cout: int, up: int = 2) -> None:
        super().__init__()
        self.up = nn.ConvTranspose3d(cin, cout, up, stride=up, bias=False)
        self.block = ResBlock(2 * cout, cout)

    def forward(self, x: torch.Tensor, skip: torch.Tensor) -> torch.Tensor:
        return self.block(torch.cat([self.up(x), skip], dim=1))


# ---- windowed attention encoder ---------------------------------------------------------------------------------

def _windows(x: torch.Tensor, w: int) -> torch.Tensor:
    b, s, h, ww, c = x.shape
    x = x.view(b, s // w, w, h // w, w, ww // w, w, c)
    return x.permute(0, 1, 3, 5, 2, 4, 6, 7).reshape(-1, w * w * w, c)


def _unwindows(t: torch.Tensor, w: int, b: int, s: int, h: int, ww: int) -> torch.Tensor:
    c = t.shape[-1]
    t = t.view(b, s // w, h // w, ww // w, w, w, w, c)
    return t.permute(0, 1, 4, 2, 5, 3, 6, 7).reshape(b, s, h, ww, c)


class WindowAttention(nn.Module):
    def __init__(self, dim: int, window: int, heads: int) -> None:
        super().__init__()
        self.heads, self.window = heads, window
        self.qkv = nn.Linear(dim, 3 * dim)
        self.proj = nn.Linear(dim, dim)
        n = 2 * window - 1
        self.bias_table = nn.Parameter(torch.zeros(n * n * n, heads))
        nn.init.trunc_normal_(self.bias_table, std=0.02)
        coords = torch.stack(torch.meshgrid(*[torch.arange(window)] * 3, indexing="ij")).flatten(1)
        rel = (coords[:, :, None] - coords[:, None, :]).permute(1, 2, 0) + (window - 1)
        self.register_buffer("bias_index", (rel[..., 0] * n + rel[..., 1]) * n + rel[..., 2], persistent=False)

    def forward(self, x: torch.Tensor, mask: torch.Tensor | None) -> torch.Tensor:
        bw, n, c = x.shape
        qkv = self.qkv(x).view(bw, n, 3, self.heads, c // self.heads).permute(2, 0, 3, 1, 4)
        bias = self.bias_table[self.bias_index.view(-1)].view(n, n, self.heads).permute(2, 0, 1).unsqueeze(0)
        if mask is not None:  # [nW, n, n] -> broadcast over batch and heads
            nw = mask.shape[0]
            bias = (bias.unsqueeze(0) + mask.view(1, nw, 1, n, n)).expand(bw // nw, nw, self.heads, n, n)
            bias = bias.reshape(bw, self.heads, n, n)
        out = F.scaled_dot_product_attention(qkv[0], qkv[1], qkv[2], attn_mask=bias.to(x.dtype))
        return self.proj(out.transpose(1, 2).reshape(bw, n, c))


class SwinBlock(nn.Module):
    def __init__(self, dim: int, res: int, heads: int, window: int, shift: int) -> None:
        super().__init__()
        if res <= window:
            window, shift = res, 0
        self.res, self.window, self.shift = res, window, shift
        self.norm1 = nn.LayerNorm(dim)
        self.attn = WindowAttention(dim, window, heads)
        self.norm2 = nn.LayerNorm(dim)
        self.mlp = nn.Sequential(nn.Linear(dim, 4 * dim), nn.GELU(), nn.Linear(4 * dim, dim))
        mask = None
        if shift > 0:
            img = torch.zeros(1, res, res, res, 1)
            cnt = 0
            cuts = (slice(0, -window), slice(-window, -shift), slice(-shift, None))
            for a in cuts:
                for b in cuts:
                    for c in cuts:
                        img[:, a, b, c, :] = cnt
                        cnt += 1
            mw = _windows(img, window).squeeze(-1)
            diff = mw.unsqueeze(1) - mw.unsqueeze(2)
            mask = torch.zeros_like(diff).masked_fill(diff != 0, -100.0)
        self.register_buffer("mask", mask, persistent=False)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        b, l, c = x.shape
        r, w = self.res, self.window
        y = self.norm1(x).view(b, r, r, r, c)
        if self.shift:
            y = torch.roll(y, shifts=(-self.shift,) * 3, dims=(1, 2, 3))
        y = self.attn(_windows(y, w), self.mask)
        y = _unwindows(y, w, b, r, r, r)
        if self.shift:
            y = torch.roll(y, shifts=(self.shift,) * 3, dims=(1, 2, 3))
        x = x + y.reshape(b, l, c)
        return x + self.mlp(self.norm2(x))


class Stage(nn.Module):
    def __init__(self, dim: int, res: int, depth: int, heads: int, window: int) -> None:
        super().__init__()
        self.res = res
        self.blocks = nn.ModuleList(SwinBlock(dim, res, heads, window, 0 if i % 2 == 0 else window // 2) for i in range(depth))
        self.merge_norm = nn.LayerNorm(dim)
        self.merge = nn.Conv3d(dim, 2 * dim, 3, stride=2, padding=1)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        for blk in self.blocks:
            x = blk(x)
        b, _, c = x.shape
        y = self.merge_norm(F.gelu(x)).view(b, self.res, self.res, self.res, c).permute(0, 4, 1, 2, 3)
        y = self.merge(y.contiguous())
        return y.flatten(2).transpose(1, 2)  # tokens of the next (halved) resolution, doubled width


class SwinEncoder(nn.Module):
    def __init__(self, in_ch: int, roi: int = 96, dim: int = 48, depths: Sequence[int] = (2, 2, 2, 2),
                 heads: Sequence[int] = (3, 6, 12, 24), patch: int = 2, window: int = 6) -> None:
        super().__init__()
        self.dim, self.res0 = dim, roi // patch
        self.embed = nn.Conv3d(in_ch, dim, patch, stride=patch)
        self.embed_norm = nn.LayerNorm(dim)
        self.stages = nn.ModuleList(
            Stage(dim * 2**i, self.res0 // 2**i, depths[i], heads[i], window) for i in range(len(depths)))
        self.out_norms = nn.ModuleList(nn.LayerNorm(dim * 2 ** (i + 1)) for i in range(len(depths)))

    def forward(self, vol: torch.Tensor) -> List[torch.Tensor]:
        x = self.embed(vol)
        b, c, r = x.shape[0], x.shape[1], x.shape[2]
        tok = self.embed_norm(x.flatten(2).transpose(1, 2))
        feats = [tok.transpose(1, 2).reshape(b, c, r, r, r)]
        for stage, norm in zip(self.stages, self.out_norms):
            tok = stage(tok)
            r //= 2
            feats.append(norm(tok).transpose(1, 2).reshape(b, tok.shape[-1], r, r, r).contiguous())
        return feats


class SwinUNETRStyle(nn.Module):
    def __init__(self, in_ch: int = 1, n_classes: int = 14, roi: int = 96, dim: int = 48) -> None:
        super().__init__()
        self.encoder = SwinEncoder(in_ch, roi, dim)
        n = len(self.encoder.stages)
        self.enc_blocks = nn.ModuleList([ResBlock(in_ch, dim), ResBlock(dim, dim)] +
                                        [ResBlock(dim * 2 ** (i + 1), dim * 2 ** (i + 1)) for i in range(n)])
        self.dec_blocks = nn.ModuleList([UpBlock(dim, dim, 2)] + [UpBlock(dim * 2 ** (i + 1), dim * 2**i, 2) for i in range(n)])
        self.head = nn.Conv3d(dim, n_classes, 1)

    def forward(self, model_in: Any) -> torch.Tensor:
        vol = _unpack(model_in)
        z = self.encoder(vol)
        x = self.dec_blocks[-1](self.enc_blocks[-1](z[-1]), self.enc_blocks[-2](z[-2]))
        for i in range(1, len(self.encoder.stages)):
            x = self.dec_blocks[-(i + 1)](x, self.enc_blocks[-(i + 2)](z[-(i + 2)]))
        x = self.dec_blocks[0](x, self.enc_blocks[0](vol))
        return self.head(x)


# ---- plain UNet for cfg1 ------------------------------------------------------------------------------------------

def _cbr(cin: int, cout: int) -> nn.Sequential:
    return nn.Sequential(nn.Conv3d(cin, cout, 3, padding=1), nn.BatchNorm3d(cout), nn.ReLU(True))


class PlainUNet(nn.Module):
    def __init__(self, in_ch: int = 1, n_classes: int = 14, base: int = 16) -> None:
        super().__init__()
        self.e1 = nn.Sequential(_cbr(in_ch, base), _cbr(base, base))
        self.e2 = nn.Sequential(_cbr(base, 2 * base), _cbr(2 * base, 2 * base))
        self.e3 = nn.Sequential(_cbr(2 * base, 4 * base), _cbr(4 * base, 4 * base))
        self.u2 = nn.ConvTranspose3d(4 * base, 2 * base, 2, stride=2)
        self.d2 = nn.Sequential(_cbr(4 * base, 2 * base), _cbr(2 * base, 2 * base))
        self.u1 = nn.ConvTranspose3d(2 * base, base, 2, stride=2)
        self.d1 = nn.Sequential(_cbr(2 * base, base), _cbr(base, base))
        self.head = nn.Conv3d(base, n_classes, 1)

    def forward(self, model_in: Any) -> torch.Tensor:
        x = _unpack(model_in)
        a = self.e1(x)
        b = self.e2(F.max_pool3d(a, 2))
        c = self.e3(F.max_pool3d(b, 2))
        y = self.d2(torch.cat([self.u2(c), b], 1))
        y = self.d1(torch.cat([self.u1(y), a], 1))
        return self.head(y)


def build_backbone(name: str, in_ch: int, n_classes: int, roi: int = 96, seed: int = 13) -> nn.Module:
    torch.manual_seed(seed)
    if name == "swin_unetr":
        m: nn.Module = SwinUNETRStyle(in_ch, n_classes, roi)
    elif name == "unet":
        m = PlainUNet(in_ch, n_classes)
    else:
        raise ValueError(f"unknown backbone '{name}'")
    return m.eval()
