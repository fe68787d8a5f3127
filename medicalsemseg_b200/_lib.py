"""ctypes binding of include/mss_b200.h.  There is no fallback: a missing library is a hard error."""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

from .build import LIB_PATH

c_i32 = C.c_int32
c_i64 = C.c_int64
c_f32 = C.c_float
vp = C.c_void_p
I3 = c_i32 * 3

MSS_F32, MSS_F16, MSS_BF16 = 0, 1, 2
BLEND_CONSTANT, BLEND_PROFILES = 0, 1
GAUSS_MONAI08_ERF, GAUSS_MONAI12_EXP = 0, 1
FUSE_NONE, FUSE_LOGITS, FUSE_LABELS = 0, 1, 2
INT_CBRT, INT_SCALE, INT_RESCALE, INT_CLIP_LO, INT_CLIP_HI, INT_NORM, INT_NONZERO, INT_F64 = 1, 2, 4, 8, 16, 32, 64, 128
MAX_BATCH_PTRS = 640
MAX_VOTE_MAPS = 15
MAX_VOTE_CLASSES = 16
MAX_DICE_CLASSES = 16


class Layout(C.Structure):
    """mss_layout_t"""

    _fields_ = [
        ("image", I3), ("roi", I3), ("n_starts", I3), ("win_lo", I3), ("win_hi", I3), ("origin", I3),
        ("extent", I3), ("pitch_w", c_i32), ("n_volumes", c_i32), ("n_classes", c_i32),
        ("table_host", vp), ("table_dev", vp),
    ]


class MssError(RuntimeError):
    pass


_SIGNATURES = {
    "mss_abi_version": (C.c_int, []),
    "mss_last_error": (C.c_char_p, []),
    "mss_accumulate_last_path": (C.c_int, []),
    "mss_axis_starts": (C.c_int, [c_i32, c_i32, c_i32, C.POINTER(c_i32), c_i32]),
    "mss_geom_table_len": (c_i64, [I3, I3]),
    "mss_geom_table_build": (C.c_int, [I3, I3, I3, vp, vp, vp, vp, c_i64]),
    "mss_gaussian_profile": (C.c_int, [vp, c_i32, c_f32, c_i32, vp]),
    "mss_importance_map": (C.c_int, [vp, I3, c_i32, vp, vp, vp, c_f32, vp, vp]),
    "mss_extract_patches": (C.c_int, [vp, I3, I3, c_i32, c_f32, C.POINTER(Layout), c_i64, c_i32, vp, vp, c_i32, vp]),
    "mss_accumulate": (C.c_int, [C.POINTER(Layout), C.POINTER(vp), c_i32, c_i32, c_i32, c_i64, c_i64, vp, vp, c_i32,
                                 vp, c_i32, c_f32, vp, vp]),
    "mss_accumulate_range": (C.c_int, [C.POINTER(Layout), C.POINTER(vp), c_i32, c_i32, c_i32, c_i64, c_i64, c_i64, c_i64, vp,
                                       vp, vp]),
    "mss_finalize_labels": (C.c_int, [C.POINTER(Layout), vp, vp, c_i32, I3, I3, vp, c_i32, vp, vp, c_f32, vp, vp]),
    "mss_finalize_gather": (C.c_int, [C.POINTER(Layout), c_i32, C.POINTER(vp), vp, vp, vp, vp, vp, c_i32, vp, c_f32, vp, vp]),
    "mss_majority_vote": (C.c_int, [C.POINTER(vp), c_i32, c_i32, c_i64, vp, vp]),
    "mss_dice_counts": (C.c_int, [vp, vp, c_i32, c_i64, c_i32, vp, vp]),
    "mss_dice_counts_batched": (C.c_int, [vp, vp, c_i32, c_i64, c_i64, c_i32, vp, vp]),
    "mss_halo_add": (C.c_int, [vp, c_i64, vp, c_i64, c_i64, c_i64, vp]),
    "mss_halo_add_nd": (C.c_int, [vp, c_i64 * 4, vp, c_i64 * 4, c_i64 * 4, c_i64, vp]),
    "mss_zoom_index_table": (C.c_int, [c_i32, c_i32, vp]),
    "mss_resample_nearest": (C.c_int, [vp, I3, vp, I3, c_i64, vp, vp, vp, vp]),
    "mss_flip_copy": (C.c_int, [vp, vp, c_i64, I3, c_i32, vp]),
    "mss_mirror_merge": (C.c_int, [C.POINTER(vp), C.POINTER(c_i32), c_i32, c_f32, vp, c_i64, I3, vp]),
    "mss_accumulate_half": (C.c_int, [C.POINTER(Layout), C.POINTER(vp), c_i32, c_i32, vp, vp, vp, c_i32, vp]),
    "mss_class_boxes": (C.c_int, [vp, vp, I3, c_i32, vp, vp]),
    "mss_select_scratch_bytes": (c_i64, []),
    "mss_select2": (C.c_int, [vp, c_i32, vp, c_i64, c_i32, C.c_double, c_i64, c_i64, vp, vp, vp]),
    "mss_mask_edges": (C.c_int, [vp, I3, c_i32, I3, I3, vp, vp, vp]),
    "mss_edt_pass": (C.c_int, [vp, vp, vp, vp, I3, c_i32, vp]),
    "mss_edt_row_mask": (C.c_int, [vp, vp, I3, vp]),
    "mss_edt_pass_mask": (C.c_int, [vp, vp, vp, vp, I3, c_i32, vp]),
    "mss_dice_ce_sums": (C.c_int, [vp, c_i64, c_i64, c_i64, c_i32, c_i32, vp, c_i32, c_i32, vp, vp]),
    "mss_intensity_transform": (C.c_int, [vp, vp, c_i64, c_i32, C.c_double, C.c_double, C.c_double, C.c_double,
                                          C.c_double, C.c_double, vp]),
}

EXPORTED = tuple(sorted(_SIGNATURES))

_lib: Optional[C.CDLL] = None


def load() -> C.CDLL:
    """dlopen lib/libmss_b200.so and type its entry points.  Raises if the library has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise MssError(
            f"{LIB_PATH} is missing: build it with `python -m medicalsemseg_b200.build` "
            "(or __graft_entry__.build()).  medicalsemseg_b200 has no CPU or PyTorch fallback."
        )
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in _SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the header and the library drifted apart
        fn.restype = res
        fn.argtypes = args
    if lib.mss_abi_version() != 1:
        raise MssError(f"libmss_b200.so has ABI version {lib.mss_abi_version()}, expected 1")
    _lib = lib
    return lib


def check(rc: int, what: str) -> None:
    if rc == 0:
        return
    msg = load().mss_last_error().decode(errors="replace")
    raise MssError(f"{what} failed with code {rc}: {msg}")
