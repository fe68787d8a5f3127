"""The oracle restatements vs. the fixtures produced by RUNNING the reference's own code
(tests/golden/make_golden.py): bit-for-bit on logits, labels and vote maps."""
import hashlib
import json
import os

import numpy as np
import pytest
import torch

from oracle import sliding_window as osw
from oracle import vote as ovote
from oracle.predictors import ArithmeticPredictor
from tests.golden.cases import SW_CASES, VOTE_CASES, make_volume, make_vote_maps

GOLD = os.path.join(os.path.dirname(__file__), "golden")
MANIFEST = json.load(open(os.path.join(GOLD, "manifest.json")))


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def run_oracle(case):
    vol = torch.from_numpy(make_volume(case))
    pred = ArithmeticPredictor(case["k"])
    affine = torch.tensor([[1.5, 1.5, 2.0]] * case["shape"][0], dtype=torch.float32)
    out = osw.sliding_window_inference(
        vol, affine, case["roi"], case["sw_batch"], pred, overlap=case["overlap"], mode=case["mode"],
        cval=case.get("cval", 0.0),
    )
    return out, pred


@pytest.mark.parametrize("name", sorted(SW_CASES))
def test_sliding_window_matches_reference_run(name):
    case, gold = SW_CASES[name], MANIFEST["sliding_window"][name]
    out, pred = run_oracle(case)
    out_np = out.contiguous().numpy()
    assert list(out_np.shape) == gold["shape"]
    assert sha(out_np) == gold["sha256"]
    labels = osw.labels_from_logits(out)
    assert sha(labels) == gold["labels_sha256"]
    # the predictor saw the same patch / centre shapes, batch by batch (quirks Q3, Q10)
    assert [[list(p), None if q is None else list(q)] for p, q in pred.calls] == gold["calls"]
    fx = np.load(os.path.join(GOLD, f"sw_{name}.npz"))
    if gold["stored"] == "full":
        assert np.array_equal(fx["logits"], out_np)
        assert np.array_equal(fx["labels"], labels)
    else:
        assert np.array_equal(fx["sample"], out_np.reshape(-1)[::997])
        assert np.array_equal(fx["labels_sample"], labels.reshape(-1)[::499])


@pytest.mark.parametrize("name", sorted(VOTE_CASES))
def test_vote_matches_reference_run(name):
    case = VOTE_CASES[name]
    maps = make_vote_maps(case)
    voted = ovote.majority_vote(maps, case["k"])
    fx = np.load(os.path.join(GOLD, f"vote_{name}.npz"))
    assert np.array_equal(fx["voted"], voted)
    assert sha(voted) == MANIFEST["vote"][name]["sha256"]
    # closed-form rule of SURVEY.md section 8 a-8 is the same function
    assert np.array_equal(ovote.majority_vote_rule(maps, case["k"]), voted)
    votes = ovote.class_votes(maps, case["k"])
    assert np.array_equal(fx["votes_sum"], votes.sum(axis=(1, 2, 3)))
    assert np.all(votes[0] == 1)


from oracle import resample as oresample  # noqa: E402
from tests.golden.cases import RESAMPLE_CASES, make_label_map  # noqa: E402


@pytest.mark.parametrize("name", sorted(RESAMPLE_CASES))
def test_resample_matches_reference_run(name):
    """oracle.resample (scipy call and the index rule the kernel implements) vs outputs of the reference's own
    resample_3d (utils/misc.py:420-425, AST-extracted and run by make_golden.py)."""
    case = RESAMPLE_CASES[name]
    img = make_label_map(case)
    fx = np.load(os.path.join(GOLD, f"resample_{name}.npz"))["out"]
    assert sha(fx) == MANIFEST["resample"][name]["sha256"]
    assert np.array_equal(oresample.resample_3d(img, case["target"]), fx)
    assert np.array_equal(oresample.resample_3d_rule(img, case["target"]), fx)


def test_resample_quirk_is_pinned():
    # scipy's coordinate overshoot zeroes the last plane for (8 -> 26) and (12 -> 86): the reference does this too
    assert MANIFEST["resample"]["up_quirk"]["zero_planes_last"][:2] == [True, True]
    assert oresample.zoom_index_rule(8, 26)[-1] == -1 and oresample.zoom_index_rule(12, 86)[-1] == -1
    assert oresample.zoom_index_rule(20, 20)[-1] == 19


def test_zoom_index_table_host_matches_rule():
    from medicalsemseg_b200.resample import zoom_index_table, zoomed_shape
    for n_in in list(range(1, 40)) + [96, 155, 200, 512]:
        for n_out in list(range(1, 60)) + [155, 240, 333, 1024]:
            assert np.array_equal(zoom_index_table(n_in, n_out), oresample.zoom_index_rule(n_in, n_out)), (n_in, n_out)
    assert zoomed_shape((8, 12, 20), (26, 86, 20)) == (26, 86, 20)


from oracle import transforms as otransforms  # noqa: E402
from tests.golden.cases import INTENSITY_CASES, make_ct_volume  # noqa: E402


@pytest.mark.parametrize("name", sorted(INTENSITY_CASES))
def test_cubed_intensity_scaler_matches_reference_run(name):
    """oracle.transforms vs outputs of the reference's own ScaleCubedIntensityRange class (data/transforms.py:17-71,
    AST-extracted and run by make_golden.py under the NumPy recorded in the manifest)."""
    c = INTENSITY_CASES[name]
    fx = np.load(os.path.join(GOLD, f"intensity_{name}.npz"))["out"]
    meta = MANIFEST["intensity"][name]
    assert sha(fx) == meta["sha256"] and fx.dtype == np.float32
    vol = make_ct_volume(c)
    # the fixture was produced under NumPy >= 2: float64 intermediates (see the oracle's precision note)
    assert int(meta["numpy"].split(".")[0]) >= 2
    out64 = otransforms.scale_cubed_intensity_range(vol, c["a_min"], c["a_max"], c["b_min"], c["b_max"], c["clip"],
                                                    dtype=np.float64)
    assert np.array_equal(out64, fx)
    out32 = otransforms.scale_cubed_intensity_range(vol, c["a_min"], c["a_max"], c["b_min"], c["b_max"], c["clip"])
    assert np.max(np.abs(out32 - fx)) <= 2.5e-7 * max(1.0, float(np.abs(fx).max()))  # NumPy < 2 path: <= 1 ulp away


from oracle import nnunet as onnunet  # noqa: E402
from oracle.predictors import PositionalPredictor  # noqa: E402
from tests.golden.cases import NNUNET_CASES, make_nnunet_volume  # noqa: E402


@pytest.mark.parametrize("name", sorted(NNUNET_CASES))
def test_nnunet_tiler_matches_reference_run(name):
    """oracle.nnunet vs outputs of the reference's own SegmentationNetwork._internal_predict_3D_3Dconv_tiled
    (models/segmentors/nnformer_official/neural_network.py, imported and run by make_golden.py)."""
    c = NNUNET_CASES[name]
    fx = np.load(os.path.join(GOLD, f"nnunet_{name}.npz"))
    meta = MANIFEST["nnunet"][name]
    assert sha(fx["probs"]) == meta["sha256_probs"] and sha(fx["seg"]) == meta["sha256_seg"]
    vol = make_nnunet_volume(c)
    with torch.no_grad():
        seg, probs = onnunet.predict_3D_tiled(vol, PositionalPredictor(c["k"], c["patch"]), lambda t: t, c["k"], c["patch"],
                                              c["step"], c["mirror"], tuple(c["axes"]), c["gaussian"],
                                              all_in_gpu=c.get("all_in_gpu", False))
    assert str(probs.dtype) == meta["probs_dtype_in_reference"]
    assert np.array_equal(probs.astype(np.float32), fx["probs"]) and np.array_equal(seg.astype(np.uint8), fx["seg"])
    image = tuple(max(a, b) for a, b in zip(c["shape"][1:], c["patch"]))
    assert onnunet.compute_steps(c["patch"], image, c["step"]) == meta["steps"]


def test_nnunet_host_policy_matches_reference():
    from medicalsemseg_b200 import nnunet as N
    for name, c in NNUNET_CASES.items():
        image = tuple(max(a, b) for a, b in zip(c["shape"][1:], c["patch"]))
        assert N.compute_steps_for_sliding_window(c["patch"], image, c["step"]) == MANIFEST["nnunet"][name]["steps"]
    assert N.compute_steps_for_sliding_window((64, 64, 64), (110, 64, 200), 0.5) == [[0, 23, 46], [0], [0, 27, 54, 82, 109, 136]]
    for ps in [(96, 96, 96), (16, 16, 16), (16, 12, 20)]:
        g = N.gaussian_importance_map(ps)  # closed form, no scipy
        key = "gaussian_" + "x".join(map(str, ps))
        assert sha(g) == MANIFEST["nnunet"][key]["sha256"]
        assert np.array_equal(g, onnunet.get_gaussian(ps))
    assert N._mirror_masks((0, 1, 2), True) == list(range(8)) and N._mirror_masks((0, 2), True) == [0, 1, 4, 5]
    assert N._mirror_masks((1,), True) == [0, 2] and N._mirror_masks((0, 1, 2), False) == [0]
