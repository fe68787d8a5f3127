"""B200-native sliding-window volumetric inference for MedicalSemSeg (hot path only).

Python host side of libmss_b200.so.  See DESIGN.md for the path, INTEGRATION.md for the reference-side hook.
"""
from . import _lib  # noqa: F401
from .inferer import (  # noqa: F401
    InferStats,
    SlidingWindowInferer,
    logits_to_labels,
    sliding_window_infer,
    sliding_window_inference,
)
from .hausdorff import hausdorff_distance, mean_hausdorff  # noqa: F401
from .metrics import DiceMeter, dice_counts, dice_counts_batched, dice_from_counts, gather_volume_counts  # noqa: F401
from .resample import resample_3d  # noqa: F401
from .vote import get_new_label, majority_vote  # noqa: F401

__all__ = [
    "sliding_window_infer", "sliding_window_inference", "SlidingWindowInferer", "logits_to_labels", "InferStats",
    "majority_vote", "get_new_label", "dice_counts", "dice_from_counts", "DiceMeter", "dice_counts_batched", "gather_volume_counts", "resample_3d", "hausdorff_distance", "mean_hausdorff",
]
