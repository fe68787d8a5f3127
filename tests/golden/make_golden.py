#!/usr/bin/env python
"""Generate the golden fixtures under tests/golden/ by running the REFERENCE's own code.

Run in the build container only (needs /root/reference; the GPU box does not have it):

    python tests/golden/make_golden.py                      # everything (the full-size cases need ~25 GB RAM, minutes)
    python tests/golden/make_golden.py --sections fullsize  # one section, merged into the existing manifest.json

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
from tests.golden.cases import (FULLSIZE_LABEL_STRIDE, FULLSIZE_SAMPLE_STRIDE, FULLSIZE_SW_CASES, INTENSITY_CASES,  # noqa: E402
                                NNUNET_CASES, RESAMPLE_CASES, SW_CASES, VOTE_CASES, make_ct_volume, make_label_map,
                                make_nnunet_volume, make_volume, make_vote_maps)


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


def load_reference_segmentation_network():
    """Imports models/segmentors/nnformer_official/neural_network.py itself (unchanged, from /root/reference) with
    stand-ins for what it imports and this image lacks: batchgenerators' pad_nd_image (restated: oracle.nnunet.pad_to_patch),
    scipy.ndimage.filters (alias of scipy.ndimage), utils.misc.no_op (a null context manager; utils/misc.py drags in the
    whole project).  `Tensor.cuda` is made the identity while the reference runs, so its GPU-only code path executes on
    the CPU with the same float32 arithmetic."""
    import importlib.util
    import types

    import scipy.ndimage

    from oracle.nnunet import pad_to_patch

    def pad_nd_image(image, new_shape=None, mode="constant", kwargs=None, return_slicer=False, shape_must_be_divisible_by=None):
        res, slicer = pad_to_patch(image, new_shape)
        return res, [slice(0, image.shape[0])] + list(slicer)

    stubs = {
        "batchgenerators": types.ModuleType("batchgenerators"),
        "batchgenerators.augmentations": types.ModuleType("batchgenerators.augmentations"),
        "batchgenerators.augmentations.utils": types.ModuleType("batchgenerators.augmentations.utils"),
        "utils.misc": types.ModuleType("utils.misc"),
    }
    stubs["batchgenerators.augmentations.utils"].pad_nd_image = pad_nd_image

    class no_op:
        def __enter__(self):
            return self

        def __exit__(self, *a):
            return False

    stubs["utils.misc"].no_op = no_op
    if "scipy.ndimage.filters" not in sys.modules:
        try:
            import scipy.ndimage.filters  # noqa: F401
        except Exception:  # noqa: BLE001
            stubs["scipy.ndimage.filters"] = scipy.ndimage
    saved = {k: sys.modules.get(k) for k in stubs}
    sys.modules.update(stubs)
    try:
        path = os.path.join(REF, "models", "segmentors", "nnformer_official", "neural_network.py")
        spec = importlib.util.spec_from_file_location("ref_neural_network", path)
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    return mod


SECTIONS = ("sliding_window", "fullsize", "vote", "resample", "intensity", "nnunet", "importance_map")


def run_sliding_window(manifest: dict) -> None:
    from engine.utils import sliding_window_inference as ref_swi  # the reference itself

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


def run_fullsize(manifest: dict) -> None:
    """BASELINE.json configs[1], [2] (K=2) and [3] at their real geometry through the reference's own engine/utils.py:19-159,
    then the reference's own post-processing (engine/test.py:140-141).  Stored: sha256 of the full fp32 logits and uint8
    labels, strided samples, the count of voxels whose softmax-argmax differs from the plain argmax of the logits (softmax
    rounding can merge a near-tie), and the number of predictor calls."""
    import time

    from engine.utils import sliding_window_inference as ref_swi

    for name, c in FULLSIZE_SW_CASES.items():
        t0 = time.time()
        vol = torch.from_numpy(make_volume(c))
        pred = ArithmeticPredictor(c["k"])
        affine = torch.tensor([[1.5, 1.5, 2.0]] * c["shape"][0], dtype=torch.float32)
        with torch.no_grad():
            out = ref_swi(
                inputs=vol, affine=affine, roi_size=c["roi"], sw_batch_size=c["sw_batch"], predictor=pred,
                overlap=c["overlap"], mode=c["mode"], cval=c.get("cval", 0.0), device="cpu", sw_device="cpu",
            )
        out_np = out.contiguous().numpy()
        probs = torch.softmax(out, 1).cpu().numpy()  # engine/test.py:140
        labels = np.argmax(probs, axis=1).astype(np.uint8)[0]  # engine/test.py:141
        del probs
        plain = np.argmax(out_np, axis=1).astype(np.uint8)[0]
        top2 = np.sort(np.partition(out_np[0], -2, axis=0)[-2:], axis=0)
        gap = (top2[1] - top2[0]) / np.maximum(np.abs(top2[1]), 1e-30)
        entry = {
            "shape": list(out_np.shape), "sha256": sha(out_np), "labels_sha256": sha(labels),
            "labels_plain_argmax_sha256": sha(plain), "softmax_vs_plain_argmax_mismatch": int((plain != labels).sum()),
            "near_ties_1e-5": int((gap < 1e-5).sum()), "n_predictor_calls": len(pred.calls),
            "sample_stride": FULLSIZE_SAMPLE_STRIDE, "label_stride": FULLSIZE_LABEL_STRIDE, "stored": "sample",
            "seconds_on_cpu": round(time.time() - t0, 1), "cpu_threads": torch.get_num_threads(),
        }
        # the voxels where softmax->argmax (the reference) and the plain argmax disagree are stored so the GPU test can allow
        # exactly those (at most a handful)
        diff_idx = np.flatnonzero(plain.reshape(-1) != labels.reshape(-1)).astype(np.int64)
        np.savez_compressed(os.path.join(HERE, f"sw_{name}.npz"), sample=out_np.reshape(-1)[::FULLSIZE_SAMPLE_STRIDE].copy(),
                            labels_sample=labels.reshape(-1)[::FULLSIZE_LABEL_STRIDE].copy(), softmax_diff_index=diff_idx,
                            softmax_diff_label=labels.reshape(-1)[diff_idx].copy())
        manifest["fullsize"][name] = entry
        print(name, entry)
        del out, out_np, labels, plain, top2, gap, vol


def run_vote(manifest: dict) -> None:
    get_class_votes, get_new_label = load_reference_vote()
    for name, c in VOTE_CASES.items():
        maps = make_vote_maps(c)
        fdata = tuple(m.astype(np.float64) for m in maps)  # nib get_fdata() yields float64 (majority_vote.py:20)
        votes = get_class_votes(fdata, len(maps), c["k"])
        new = get_new_label(fdata, len(maps), c["k"]).astype(np.uint8)  # cast of majority_vote.py:83
        np.savez_compressed(os.path.join(HERE, f"vote_{name}.npz"), voted=new, votes_sum=votes.sum(axis=(1, 2, 3)))
        manifest["vote"][name] = {"sha256": sha(new), "shape": list(new.shape)}
        print("vote", name, new.shape, sha(new)[:12])


def run_resample(manifest: dict) -> None:
    ref_resample = load_reference_resample()
    for name, c in RESAMPLE_CASES.items():
        img = make_label_map(c)
        out = ref_resample(img, c["target"])
        np.savez_compressed(os.path.join(HERE, f"resample_{name}.npz"), out=out)
        manifest["resample"][name] = {"sha256": sha(out), "shape": list(out.shape), "zero_planes_last": [
            bool((np.take(out, -1, axis=a) == 0).all()) for a in range(3)]}
        print("resample", name, out.shape, sha(out)[:12])


def run_intensity(manifest: dict) -> None:
    scaler_cls = load_reference_cubed_scaler()
    for name, c in INTENSITY_CASES.items():
        vol = make_ct_volume(c)
        out = scaler_cls(c["a_min"], c["a_max"], c["b_min"], c["b_max"], c["clip"])(vol)
        np.savez_compressed(os.path.join(HERE, f"intensity_{name}.npz"), out=out)
        manifest["intensity"][name] = {"sha256": sha(out), "dtype": str(out.dtype), "numpy": np.__version__,
                                       "min": float(out.min()), "max": float(out.max())}
        print("intensity", name, out.shape, out.dtype, sha(out)[:12])


def run_nnunet(manifest: dict) -> None:
    from oracle.predictors import PositionalPredictor
    nn_mod = load_reference_segmentation_network()

    class RefNet(nn_mod.SegmentationNetwork):
        def __init__(self, k, patch, device=0):
            super().__init__()
            self.num_classes = k
            self.conv_op = torch.nn.Conv3d
            self.inference_apply_nonlin = lambda t: t
            self.pred = PositionalPredictor(k, patch)
            self._device = device

        def get_device(self):
            return self._device

        def forward(self, t):
            return self.pred(t)

    real_cuda = torch.Tensor.cuda
    torch.Tensor.cuda = lambda self, *a, **k: self
    try:
        for name, c in NNUNET_CASES.items():
            vol = make_nnunet_volume(c)
            half = bool(c.get("all_in_gpu", False))
            # all_in_gpu allocates with device=self.get_device() (:352-365): a torch.device("cpu") keeps the reference's
            # `get_device() != "cpu"` assertion true (a torch.device never equals a str) and its half-precision branch on the CPU
            net = RefNet(c["k"], c["patch"], torch.device("cpu") if half else 0)
            with torch.no_grad():
                seg, probs = net._internal_predict_3D_3Dconv_tiled(
                    vol, c["step"], c["mirror"], tuple(c["axes"]), tuple(c["patch"]), None, c["gaussian"], "constant",
                    {"constant_values": 0}, half, False)
            if half:
                assert probs.dtype == np.float16, probs.dtype
            np.savez_compressed(os.path.join(HERE, f"nnunet_{name}.npz"), seg=seg.astype(np.uint8),
                                probs=probs.astype(np.float32))
            manifest["nnunet"][name] = {"sha256_probs": sha(probs.astype(np.float32)), "sha256_seg": sha(seg.astype(np.uint8)),
                                        "shape": list(probs.shape), "probs_dtype_in_reference": str(probs.dtype),
                                        "steps": nn_mod.SegmentationNetwork._compute_steps_for_sliding_window(
                                            tuple(c["patch"]), tuple(max(a, b) for a, b in zip(c["shape"][1:], c["patch"])), c["step"])}
            print("nnunet", name, probs.shape, probs.dtype, sha(probs.astype(np.float32))[:12])
        for ps in [(96, 96, 96), (16, 16, 16), (16, 12, 20)]:
            g = nn_mod.SegmentationNetwork._get_gaussian(ps, sigma_scale=1.0 / 8)
            manifest["nnunet"]["gaussian_" + "x".join(map(str, ps))] = {"sha256": sha(g), "min": float(g.min())}
    finally:
        torch.Tensor.cuda = real_cuda


def run_importance_map(manifest: dict) -> None:
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


def main() -> None:
    import argparse

    ap = argparse.ArgumentParser()
    ap.add_argument("--sections", nargs="*", default=list(SECTIONS), choices=SECTIONS)
    args = ap.parse_args()
    torch.set_num_threads(os.cpu_count() or 1)
    path = os.path.join(HERE, "manifest.json")
    manifest = json.load(open(path)) if os.path.exists(path) else {}
    for sec in args.sections:
        manifest[sec] = {}
        globals()["run_" + sec](manifest)
    with open(path, "w") as f:
        json.dump(manifest, f, indent=1, sort_keys=True)
    print("wrote", path)


if __name__ == "__main__":
    main()
