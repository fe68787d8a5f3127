#!/usr/bin/env python
"""Generate the golden fixtures under tests/golden/ by running the REFERENCE's own code.

Run in the build container only (needs /root/reference; the GPU box does not have it):

    python tests/golden/make_golden.py

* ``engine/utils.py::sliding_window_inference`` is imported from ``/root/reference`` and
  executed VERBATIM on CPU; its MONAI imports (engine/utils.py:5-13) are served by
  ``oracle/monai_shim`` (MONAI itself is not installable here - no network).
* ``get_class_votes`` / ``get_new_label`` are AST-extracted from
  ``/root/reference/majority_vote.py:23-37`` (the file runs argparse at import time and needs
  nibabel, so it cannot be imported) and executed unchanged.

Nothing is copied from the reference; only its OUTPUTS on seeded inputs are stored.
Inputs are regenerated in the tests from ``numpy.random.RandomState`` seeds (frozen legacy
generator), small outputs are stored in full, large ones as sha256 + a strided sample.
"""
from __future__ import annotations

import ast
import hashlib
import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle", "monai_shim"))
sys.path.insert(0, REF)

from oracle.predictors import ArithmeticPredictor  # noqa: E402
from tests.golden.cases import (INTENSITY_CASES, RESAMPLE_CASES, SW_CASES, VOTE_CASES, make_ct_volume,  # noqa: E402
                                make_label_map, make_volume, make_vote_maps)


def sha(a: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def load_reference_vote():
    src = open(os.path.join(REF, "majority_vote.py")).read()
    tree = ast.parse(src)
    keep = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name in ("get_class_votes", "get_new_label")]
    ns = {"np": np}
    exec(compile(ast.Module(body=keep, type_ignores=[]), "majority_vote.py", "exec"), ns)  # noqa: S102
    return ns["get_class_votes"], ns["get_new_label"]


def load_reference_resample():
    """utils/misc.py imports half the project at module level; its resample_3d (:420-425) is pure scipy and is
    AST-extracted and executed unchanged."""
    from scipy import ndimage

    src = open(os.path.join(REF, "utils", "misc.py")).read()
    tree = ast.parse(src)
    keep = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name == "resample_3d"]
    ns = {"ndimage": ndimage, "np": np}
    exec(compile(ast.Module(body=keep, type_ignores=[]), "utils/misc.py", "exec"), ns)  # noqa: S102
    return ns["resample_3d"]


def load_reference_cubed_scaler():
    """data/transforms.py imports MONAI at module level (absent here); the class ScaleCubedIntensityRange (:17-71) only
    needs a base class, a clip and a dtype cast from it.  It is AST-extracted and executed unchanged with minimal
    stand-ins for those three names (np.clip and ndarray.astype are what MONAI's helpers do for NumPy inputs)."""
    from typing import Optional
    from warnings import warn

    src = open(os.path.join(REF, "data", "transforms.py")).read()
    tree = ast.parse(src)
    keep = [n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == "ScaleCubedIntensityRange"]

    class _Backends:
        TORCH, NUMPY = "torch", "numpy"

    ns = {"np": np, "Optional": Optional, "warn": warn, "Transform": object, "TransformBackends": _Backends,
          "DtypeLike": object, "NdarrayOrTensor": object,
          "clip": lambda a, lo, hi: np.clip(a, lo, hi),
          "convert_data_type": lambda img, dtype=None: (np.asarray(img).astype(dtype), None, None)}
    exec(compile(ast.Module(body=keep, type_ignores=[]), "data/transforms.py", "exec"), ns)  # noqa: S102
    return ns["ScaleCubedIntensityRange"]


def main() -> None:
    torch.set_num_threads(os.cpu_count() or 1)
    from engine.utils import sliding_window_inference as ref_swi  # the reference itself

    manifest = {"sliding_window": {}, "vote": {}, "importance_map": {}, "resample": {}, "intensity": {}}
    for name, c in SW_CASES.items():
        vol = torch.from_numpy(make_volume(c))
        pred = ArithmeticPredictor(c["k"])
        affine = torch.tensor([[1.5, 1.5, 2.0]] * c["shape"][0], dtype=torch.float32)
        with torch.no_grad():
            out = ref_swi(
                inputs=vol, affine=affine, roi_size=c["roi"], sw_batch_size=c["sw_batch"], predictor=pred,
                overlap=c["overlap"], mode=c["mode"], cval=c.get("cval", 0.0), device="cpu", sw_device="cpu",
            )
        out_np = out.contiguous().numpy()
        probs = torch.softmax(out, 1).cpu().numpy()  # engine/test.py:140-141 run here as-is
        labels = np.argmax(probs, axis=1).astype(np.uint8)[0]
        entry = {
            "shape": list(out_np.shape), "sha256": sha(out_np), "labels_sha256": sha(labels),
            "calls": [[list(p), None if q is None else list(q)] for p, q in pred.calls],
        }
        if out_np.nbytes <= 600_000:
            np.savez_compressed(os.path.join(HERE, f"sw_{name}.npz"), logits=out_np, labels=labels)
            entry["stored"] = "full"
        else:
            flat = out_np.reshape(-1)
            np.savez_compressed(os.path.join(HERE, f"sw_{name}.npz"), sample=flat[:: 997].copy(),
                                labels_sample=labels.reshape(-1)[:: 499].copy())
            entry["stored"] = "sample"
        manifest["sliding_window"][name] = entry
        print(name, entry["shape"], entry["sha256"][:12], len(pred.calls), "predictor calls")

    get_class_votes, get_new_label = load_reference_vote()
    for name, c in VOTE_CASES.items():
        maps = make_vote_maps(c)
        fdata = tuple(m.astype(np.float64) for m in maps)  # nib get_fdata() yields float64 (majority_vote.py:20)
        votes = get_class_votes(fdata, len(maps), c["k"])
        new = get_new_label(fdata, len(maps), c["k"]).astype(np.uint8)  # cast of majority_vote.py:83
        np.savez_compressed(os.path.join(HERE, f"vote_{name}.npz"), voted=new, votes_sum=votes.sum(axis=(1, 2, 3)))
        manifest["vote"][name] = {"sha256": sha(new), "shape": list(new.shape)}
        print("vote", name, new.shape, sha(new)[:12])

    ref_resample = load_reference_resample()
    for name, c in RESAMPLE_CASES.items():
        img = make_label_map(c)
        out = ref_resample(img, c["target"])
        np.savez_compressed(os.path.join(HERE, f"resample_{name}.npz"), out=out)
        manifest["resample"][name] = {"sha256": sha(out), "shape": list(out.shape), "zero_planes_last": [
            bool((np.take(out, -1, axis=a) == 0).all()) for a in range(3)]}
        print("resample", name, out.shape, sha(out)[:12])

    scaler_cls = load_reference_cubed_scaler()
    for name, c in INTENSITY_CASES.items():
        vol = make_ct_volume(c)
        out = scaler_cls(c["a_min"], c["a_max"], c["b_min"], c["b_max"], c["clip"])(vol)
        np.savez_compressed(os.path.join(HERE, f"intensity_{name}.npz"), out=out)
        manifest["intensity"][name] = {"sha256": sha(out), "dtype": str(out.dtype), "numpy": np.__version__,
                                       "min": float(out.min()), "max": float(out.max())}
        print("intensity", name, out.shape, out.dtype, sha(out)[:12])

    # importance maps as the shimmed MONAI-0.8 restatement produces them (unpinned, recorded for drift detection)
    from oracle.monai08 import compute_importance_map
    for roi in [(96, 96, 96), (16, 16, 16), (24, 16, 32), (8, 12, 20)]:
        m = compute_importance_map(roi, mode="gaussian", sigma_scale=0.125).numpy()
        key = "x".join(map(str, roi))
        manifest["importance_map"][key] = {
            "sha256": sha(m), "min": float(m.min()), "max": float(m.max()),
            "axis0_profile_head": [float(v) for v in m[:3, roi[1] // 2, roi[2] // 2]],
            "axis0_profile_tail": [float(v) for v in m[-2:, roi[1] // 2, roi[2] // 2]],
        }
    with open(os.path.join(HERE, "manifest.json"), "w") as f:
        json.dump(manifest, f, indent=1, sort_keys=True)
    print("wrote", os.path.join(HERE, "manifest.json"))


if __name__ == "__main__":
    main()
