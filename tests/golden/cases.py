"""Seeded case table shared by tests/golden/make_golden.py (which runs the reference) and the tests.

Inputs come from ``numpy.random.RandomState`` (the frozen legacy generator), so they are
regenerated identically on any box; only reference OUTPUTS are stored as fixtures.
"""
from __future__ import annotations

import numpy as np

# shape = [Nb, Cin, D, H, W]; roi may contain non-positive entries (fall back to the image dim)
SW_CASES = {
    # sw_batch 1 -> centres get the extra leading dim (quirk Q3); clamped last windows on every axis
    "basic_b1": dict(shape=[1, 1, 40, 36, 44], roi=[16, 16, 16], overlap=0.25, mode="gaussian", sw_batch=1, k=3, seed=1),
    # anisotropic roi, overlap .5, ragged final batch, W not a multiple of 4
    "aniso_ragged": dict(shape=[1, 2, 41, 37, 50], roi=[24, 16, 32], overlap=0.5, mode="gaussian", sw_batch=4, k=5, seed=2),
    # two volumes: predictor batches straddle the volume boundary (quirk Q10)
    "two_volumes": dict(shape=[2, 1, 30, 28, 33], roi=[16, 16, 16], overlap=0.5, mode="gaussian", sw_batch=3, k=4, seed=3),
    # one dim smaller than the roi -> symmetric constant pad with the reference's air value (engine/test.py:103-104)
    "padded_cval": dict(shape=[1, 1, 10, 40, 21], roi=[16, 16, 16], overlap=0.25, mode="gaussian", sw_batch=2, k=3,
                        seed=4, cval=(0.0 - 0.1943) / 0.2786),
    "constant_mode": dict(shape=[1, 1, 33, 35, 38], roi=[16, 16, 16], overlap=0.5, mode="constant", sw_batch=4, k=3, seed=5),
    # BraTS-like: 4 input channels, 3 classes, odd W
    "brats_like": dict(shape=[1, 4, 36, 36, 31], roi=[16, 16, 16], overlap=0.5, mode="gaussian", sw_batch=4, k=3, seed=6),
    # roi spans the image on one axis (interval = roi), scalar roi with fall-back (-1) on another
    "full_axis": dict(shape=[1, 1, 16, 40, 40], roi=[16, -1, 24], overlap=0.25, mode="gaussian", sw_batch=2, k=3, seed=7),
    "overlap_zero": dict(shape=[1, 1, 32, 40, 48], roi=[16, 16, 16], overlap=0.0, mode="gaussian", sw_batch=5, k=2, seed=8),
    # high overlap: 3+ windows cover a voxel per axis
    "overlap_075": dict(shape=[1, 1, 30, 30, 30], roi=[16, 16, 16], overlap=0.75, mode="gaussian", sw_batch=8, k=3, seed=9),
    # BASELINE.json configs[0] stitching geometry: 128^3, roi 96^3, overlap .25, 14 classes, N=8
    "cfg1_geometry": dict(shape=[1, 1, 128, 128, 128], roi=[96, 96, 96], overlap=0.25, mode="gaussian", sw_batch=4, k=14, seed=10),
}

# BASELINE.json configs at their REAL geometry: the reference's engine/utils.py is run on them once in
# the build container (minutes, gigabytes); fixtures hold sha256 of the full logits / labels plus a strided sample.
#   cfg2: 512x512x200, K=14, 400 windows (starts ... 384, 416 | 0, 48, 96, 104)
#   cfg4: 4-channel 240x240x155, K=3, 48 windows (starts 0, 48, 96, 144 | 0, 48, 59: unaligned clamped start, odd W)
#   cfg3: 512x512x1024, 2100 windows (z starts ... 912, 928); K cut to 2 so the reference's K-replicated count map fits
FULLSIZE_SW_CASES = {
    "cfg2_geometry": dict(shape=[1, 1, 512, 512, 200], roi=[96, 96, 96], overlap=0.5, mode="gaussian", sw_batch=4, k=14, seed=11),
    "cfg4_geometry": dict(shape=[1, 4, 240, 240, 155], roi=[96, 96, 96], overlap=0.5, mode="gaussian", sw_batch=4, k=3, seed=12),
    "cfg3_geometry_k2": dict(shape=[1, 1, 512, 512, 1024], roi=[96, 96, 96], overlap=0.5, mode="gaussian", sw_batch=4, k=2, seed=13),
}
FULLSIZE_SAMPLE_STRIDE = 99991   # prime; logits sample
FULLSIZE_LABEL_STRIDE = 9973

VOTE_CASES = {
    "k3_m5": dict(shape=[24, 20, 31], k=3, m=5, seed=21),        # BraTS ensemble shape class (cfg4: K=3, M=5)
    "k14_m5": dict(shape=[17, 23, 29], k=14, m=5, seed=22),
    "k14_m3_stray": dict(shape=[16, 16, 18], k=14, m=3, seed=23, stray=True),  # labels >= K must be ignored
    "k2_m1": dict(shape=[8, 9, 10], k=2, m=1, seed=24),          # a single fold can never out-vote background
    "k16_m15": dict(shape=[12, 12, 13], k=16, m=15, seed=25),
}


def make_volume(case: dict) -> np.ndarray:
    rs = np.random.RandomState(case["seed"])
    return rs.standard_normal(case["shape"]).astype(np.float32)


def make_vote_maps(case: dict) -> list:
    """Base map + per-model random relabel (SURVEY.md section 8d synthetic ensemble), uint8."""
    rs = np.random.RandomState(case["seed"])
    k, m = case["k"], case["m"]
    base = rs.randint(0, k, size=case["shape"]).astype(np.uint8)
    maps = []
    for _ in range(m):
        flip = rs.random_sample(case["shape"]) < 0.35
        hi = k + 3 if case.get("stray") else k
        noise = rs.randint(0, hi, size=case["shape"]).astype(np.uint8)
        maps.append(np.where(flip, noise, base).astype(np.uint8))
    return maps


# label-map resampling after the argmax (utils/misc.py:420-425): seeded uint8 maps zoomed to a target grid.
# (8, 12) -> (26, 86) are size pairs where scipy's coordinate overshoots and the last plane becomes 0.
RESAMPLE_CASES = {
    "up_quirk": dict(shape=[8, 12, 20], target=[26, 86, 20], k=14, seed=31),
    "down": dict(shape=[40, 36, 31], target=[17, 36, 12], k=14, seed=32),
    "mixed": dict(shape=[33, 20, 48], target=[50, 7, 61], k=3, seed=33),
    "identity": dict(shape=[16, 16, 16], target=[16, 16, 16], k=5, seed=34),
    "to_one": dict(shape=[9, 10, 11], target=[1, 10, 23], k=4, seed=35),
    "spacing_like": dict(shape=[96, 96, 64], target=[128, 128, 43], k=14, seed=36),   # Spacingd round trip shape class
}


def make_label_map(case: dict) -> np.ndarray:
    rs = np.random.RandomState(case["seed"])
    return rs.randint(0, case["k"], size=case["shape"]).astype(np.uint8)


# test-time intensity transform (the reference's own ScaleCubedIntensityRange, data/transforms.py:17-71, configured as at
# data/dataset_builder.py:333-341): seeded CT-like volumes in Hounsfield units
INTENSITY_CASES = {
    "ct_default": dict(shape=[1, 24, 20, 28], a_min=-1000, a_max=1000, b_min=0.0, b_max=1.0, clip=True, seed=41),
    "ct_soft": dict(shape=[1, 16, 16, 33], a_min=-175, a_max=250, b_min=0.0, b_max=1.0, clip=True, seed=42),
    "noclip": dict(shape=[2, 9, 10, 11], a_min=-500, a_max=1500, b_min=-1.0, b_max=1.0, clip=False, seed=43),
    "no_b": dict(shape=[1, 8, 8, 8], a_min=-1000, a_max=1000, b_min=None, b_max=None, clip=False, seed=44),
}


def make_ct_volume(case: dict) -> np.ndarray:
    rs = np.random.RandomState(case["seed"])
    return (rs.standard_normal(case["shape"]) * 700.0 - 200.0).astype(np.float32)


# nnU-Net-style tiled predictor (models/segmentors/nnformer_official/neural_network.py:300-437, :511-568): seeded volumes
# [C, X, Y, Z], PositionalPredictor (not flip-equivariant) with an identity non-linearity so CPU and GPU logits agree
NNUNET_CASES = {
    "mirror_all": dict(shape=[1, 40, 36, 44], patch=[16, 16, 16], step=0.5, mirror=True, axes=(0, 1, 2), gaussian=True, k=3, seed=51),
    "mirror_two_axes": dict(shape=[2, 30, 28, 33], patch=[16, 12, 20], step=0.5, mirror=True, axes=(0, 2), gaussian=True, k=4, seed=52),
    # (use_gaussian=False with more than one tile raises in the reference: its ones-map has the image's shape, :381)
    "no_mirror": dict(shape=[1, 33, 35, 38], patch=[16, 16, 16], step=0.75, mirror=False, axes=(0, 1, 2), gaussian=True, k=3, seed=53),
    "padded_single_tile": dict(shape=[1, 10, 14, 16], patch=[16, 16, 16], step=0.5, mirror=True, axes=(0, 1, 2), gaussian=True, k=2, seed=54),
    "fine_steps": dict(shape=[1, 30, 30, 30], patch=[16, 16, 16], step=0.3, mirror=True, axes=(1,), gaussian=True, k=3, seed=55),
    # all_in_gpu=True (:349-356, :399-400, :420-431): half importance map, half aggregated results / counts, half division
    "half_mirror_all": dict(shape=[1, 40, 36, 44], patch=[16, 16, 16], step=0.5, mirror=True, axes=(0, 1, 2), gaussian=True, k=3,
                            seed=56, all_in_gpu=True),
    "half_no_mirror_odd": dict(shape=[2, 33, 29, 35], patch=[16, 12, 20], step=0.4, mirror=False, axes=(0, 1, 2), gaussian=True,
                               k=4, seed=57, all_in_gpu=True),
    "half_single_tile": dict(shape=[1, 10, 14, 16], patch=[16, 16, 16], step=0.5, mirror=True, axes=(0, 2), gaussian=True, k=2,
                             seed=58, all_in_gpu=True),
}


def make_nnunet_volume(case: dict) -> np.ndarray:
    rs = np.random.RandomState(case["seed"])
    return rs.standard_normal(case["shape"]).astype(np.float32)
