"""The C-ABI library builds, loads without a GPU and exports exactly what include/mss_b200.h declares."""
import os
import re
import subprocess

import pytest

import medicalsemseg_b200 as mss
from medicalsemseg_b200 import _lib, build

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "mss_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(mss_[a-z0-9_]+)\s*\(", text)))


def test_library_is_built_and_loads():
    assert os.path.exists(build.LIB_PATH), "run __graft_entry__.build() first"
    lib = _lib.load()
    assert lib.mss_abi_version() == 1


def test_header_and_library_agree():
    decl = declared_symbols()
    assert decl == sorted(_lib.EXPORTED), "ctypes table and header drifted apart"
    out = subprocess.run(["nm", "-D", "--defined-only", build.LIB_PATH], capture_output=True, text=True, check=True).stdout
    exported = set(re.findall(r" T (mss_[a-z0-9_]+)", out))
    assert set(decl) <= exported, f"missing exports: {sorted(set(decl) - exported)}"


def test_argument_errors_are_reported_not_thrown():
    lib = _lib.load()
    assert lib.mss_axis_starts(10, 20, 5, None, 0) == -1  # roi > image
    assert b"roi" in lib.mss_last_error()
    assert lib.mss_gaussian_profile(None, 4, 1.0, 0, None) == -1
    assert lib.mss_majority_vote(None, 1, 2, 10, None, None) == -1
    assert lib.mss_dice_counts(None, None, 0, 10, 3, None, None) == -1
    assert lib.mss_dice_counts_batched(None, None, 0, 10, 2, 3, None, None) == -1
    assert lib.mss_halo_add(None, 1, None, 1, 1, 1, None) == -1
    I4 = _lib.c_i64 * 4
    assert lib.mss_halo_add_nd(None, I4(0, 0, 0, 0), None, I4(0, 0, 0, 0), I4(1, 1, 1, 1), 4, None) == -1
    assert lib.mss_zoom_index_table(0, 4, None) == -1
    assert lib.mss_dice_ce_sums(None, 1, 1, 1, 1, 2, None, 0, 1, None, None) == -1
    assert lib.mss_mask_edges(None, _lib.I3(1, 1, 1), 0, _lib.I3(0, 0, 0), _lib.I3(1, 1, 1), None, None, None) == -1
    assert lib.mss_edt_pass(None, None, None, None, _lib.I3(1, 1, 1), 0, None) == -1
    assert lib.mss_edt_pass_mask(None, None, None, None, _lib.I3(1, 1, 1), 0, None) == -1
    assert lib.mss_flip_copy(None, None, 1, _lib.I3(1, 1, 1), 0, None) == -1
    assert lib.mss_mirror_merge(None, None, 1, 1.0, None, 1, _lib.I3(1, 1, 1), None) == -1
    assert lib.mss_intensity_transform(None, None, 8, 2, 0.0, 1.0, 0.0, 1.0, 0.0, 1.0, None) == -1
    assert lib.mss_resample_nearest(None, _lib.I3(1, 1, 1), None, _lib.I3(1, 1, 1), 1, None, None, None, None) == -1


def test_no_silent_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    vol = torch.zeros(1, 1, 16, 16, 16)
    with pytest.raises(_lib.MssError):
        mss.sliding_window_infer(vol, lambda x: x, roi=8)
    with pytest.raises(_lib.MssError):
        mss.majority_vote([torch.zeros(4, 4, 4, dtype=torch.uint8)], 2)
    with pytest.raises(_lib.MssError):
        mss.dice_counts(torch.zeros(8, dtype=torch.uint8), torch.zeros(8, dtype=torch.uint8), 2)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "medicalsemseg_b200")
    for dirpath, _dirs, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f"{f} imports the oracle"


def test_docs_state_the_real_entry_point_count():
    n = len(declared_symbols())
    for doc in ("DESIGN.md", "README.md"):
        text = open(os.path.join(ROOT, doc)).read()
        claimed = {int(m) for m in re.findall(r"(\d+) entry points", text)}
        assert claimed == {n}, f"{doc} says {sorted(claimed)} entry points, include/mss_b200.h declares {n}"
