"""Shared pieces of the GPU parity tests: run the oracle (CPU) and the CUDA path on the same seeded inputs."""
import numpy as np
import torch

import medicalsemseg_b200 as mss
from oracle import monai08 as M
from oracle import sliding_window as osw
from oracle.predictors import ArithmeticPredictor
from tests.golden.cases import make_volume

TOL = 1e-5  # BASELINE.json: accumulated logits within 1e-5 relative; labels exact outside top-2 gaps below it


def oracle_run(case, tuple_input=True, importance_map=None):
    vol = torch.from_numpy(make_volume(case))
    pred = ArithmeticPredictor(case["k"])
    affine = torch.tensor([[1.5, 1.5, 2.0]] * case["shape"][0], dtype=torch.float32)
    out = osw.sliding_window_inference(
        vol, affine, case["roi"], case["sw_batch"], pred, overlap=case["overlap"], mode=case["mode"],
        cval=case.get("cval", 0.0), tuple_input=tuple_input, importance_map=importance_map)
    return out, pred


def cuda_inputs(case):
    vol = torch.from_numpy(make_volume(case)).cuda()
    affine = torch.tensor([[1.5, 1.5, 2.0]] * case["shape"][0], dtype=torch.float32, device="cuda")
    return vol, affine


def assert_labels_match(labels_cuda, logits_ref, tol=TOL):
    """Bit-exact labels except voxels whose top-2 relative gap (in the oracle's logits) is below tol; returns #mismatch."""
    want = np.stack([osw.labels_from_logits(logits_ref[b:b + 1]) for b in range(logits_ref.shape[0])])
    got = labels_cuda.cpu().numpy()
    assert got.shape == want.shape and got.dtype == np.uint8
    bad = got != want
    if bad.any():
        gaps = np.stack([osw.top2_relative_gap(logits_ref[b:b + 1]) for b in range(logits_ref.shape[0])])
        assert np.all(gaps[bad] < tol), f"{int(bad.sum())} label mismatches, {int((gaps[bad] >= tol).sum())} outside near-ties"
    return int(bad.sum())


def rel_err(got, ref):
    """max |got - ref| / max(|ref|, scale) with scale = rms of ref: the 1e-5 'relative' bar for stitched logits."""
    ref64 = ref.double()
    scale = ref64.pow(2).mean().sqrt().clamp_min(1e-30)
    return ((got.double() - ref64).abs() / torch.maximum(ref64.abs(), scale)).max().item()
