"""Known-answer tests for the MONAI-0.8 helper restatement (oracle/monai08.py).

The reference holds no tests; these answers are hand-derived from the reference call sites
(engine/utils.py:95-115) and tabulated in SURVEY.md section 8(a-1), 8(a-3) and BASELINE.md section 3."""
import math

import numpy as np
import pytest
import torch

from oracle import dice as odice
from oracle import monai08 as M
from oracle import sliding_window as osw


@pytest.mark.parametrize(
    "image,roi,overlap,interval,starts",
    [
        ((128, 128, 128), (96, 96, 96), 0.25, (72, 72, 72), [[0, 32]] * 3),
        ((512, 512, 200), (96, 96, 96), 0.5, (48, 48, 48),
         [[0, 48, 96, 144, 192, 240, 288, 336, 384, 416]] * 2 + [[0, 48, 96, 104]]),
        ((240, 240, 155), (96, 96, 96), 0.5, (48, 48, 48), [[0, 48, 96, 144]] * 2 + [[0, 48, 59]]),
        ((96, 200, 96), (96, 96, 96), 0.5, (96, 48, 96), [[0], [0, 48, 96, 104], [0]]),
    ],
)
def test_window_grid_known_answers(image, roi, overlap, interval, starts):
    got_interval, got_starts = osw.window_grid(image, roi, overlap)
    assert got_interval == interval
    assert got_starts == starts


def test_whole_body_grid():
    _, starts = osw.window_grid((512, 512, 1024), (96, 96, 96), 0.5)
    assert [len(s) for s in starts] == [10, 10, 21]
    assert starts[2][-2:] == [912, 928]
    assert len(M.dense_patch_slices((512, 512, 1024), (96, 96, 96), (48, 48, 48))) == 2100


def test_dense_patch_slices_c_order():
    sl = M.dense_patch_slices((20, 20, 20), (16, 16, 16), (12, 12, 12))
    assert len(sl) == 8
    assert [tuple(s.start for s in w) for w in sl[:3]] == [(0, 0, 0), (0, 0, 4), (0, 4, 0)]


def test_fall_back_tuple_and_interval_errors():
    assert M.fall_back_tuple(96, (128, 100, 90)) == (96, 96, 96)
    assert M.fall_back_tuple((96, -1, None), (128, 100, 90)) == (96, 100, 90)
    with pytest.raises(ValueError):
        M.get_scan_interval((10, 10), (4, 4, 4), 2, 0.5)
    assert M.get_scan_interval((10, 10, 10), (4, 4, 4), 3, 0.99) == (1, 1, 1)


def test_importance_map_96_profile():
    m = M.compute_importance_map((96, 96, 96), mode="gaussian", sigma_scale=0.125)
    assert m.dtype == torch.float32 and tuple(m.shape) == (96, 96, 96)
    assert m[48, 48, 48].item() == 1.0 and m.max().item() == 1.0  # centred at i//2 = 48 (quirk Q2)
    t = 0.70710678 / 12.0

    def g(x):
        return 0.5 * (math.erf(t * (x + 0.5)) - math.erf(t * (x - 0.5)))

    prof = m[:, 48, 48].numpy()
    want = np.array([g(i - 48) / g(0) for i in range(96)])
    # float32 erf differences cancel badly in the tails (|erf|~1): ~1e-3 relative there, exact at the centre
    assert np.allclose(prof, want, rtol=3e-3) and np.allclose(prof[24:72], want[24:72], rtol=1e-5)
    assert prof[0] == pytest.approx(3.37e-4, rel=2e-2) and prof[95] == pytest.approx(4.69e-4, rel=2e-2)
    assert prof[0] < prof[95]  # asymmetric
    assert m.min().item() == pytest.approx(prof[0] ** 3, rel=1e-3) and m.min().item() > 0
    # the map is the rounded outer product ((p_i * p_j) * p_k) / max in float32
    g32 = M.gaussian_1d(torch.tensor(12.0))[:96]  # taps x=-48..47
    outer = (g32[:, None, None] * g32[None, :, None]) * g32[None, None, :]
    assert torch.equal(outer / outer.max(), m)


def test_importance_map_constant_and_v12():
    assert torch.equal(M.compute_importance_map((4, 5, 6), mode="constant"), torch.ones(4, 5, 6))
    v12 = M.compute_importance_map_v12((96, 96, 96))
    assert v12.min().item() == pytest.approx(1e-3) and v12.max().item() < 1.0


def test_dice_counts_and_nan_rules():
    pred = np.array([0, 1, 1, 2, 2, 2], dtype=np.uint8)
    lab = np.array([0, 1, 2, 2, 2, 0], dtype=np.uint8)
    c = odice.dice_counts(pred, lab, 4)
    assert c.tolist() == [[1, 1, 2, 0], [1, 2, 3, 0], [2, 1, 3, 0]]
    d = odice.dice_from_counts(c)
    assert d[:3].tolist() == [2 / 3, 2 / 3, 4 / 6] and np.isnan(d[3])
    means, m = odice.class_means(np.stack([d, d]))
    assert np.isnan(means[3]) and m == pytest.approx(np.nanmean(d))
    logits = torch.zeros(4, 1, 1, 6)
    for i, p in enumerate(pred):
        logits[p, 0, 0, i] = 1.0
    md = odice.monai_meandice(logits, torch.from_numpy(lab.astype(np.float32)).reshape(1, 1, 1, 6), 4)
    assert np.allclose(md.numpy()[0, :3], d[:3]) and np.isnan(md.numpy()[0, 3])


def test_eval_meters_mdice_is_the_mean_over_volumes_of_the_per_volume_class_mean():
    """engine/test.py:59-76 + utils/misc.py:93-100: mDice is logged per iteration (one volume) and averaged by the meter.
    A = (1.0, 0.5), B = (0.8, nan): the reference reports (0.75 + 0.8) / 2 = 0.775, not nanmean((0.9, 0.5)) = 0.70."""
    from medicalsemseg_b200.metrics import DiceMeter

    d = np.array([[1.0, 0.5], [0.8, np.nan]])
    cls, m = odice.eval_meters(d)
    assert cls.tolist() == [0.9, 0.5] and m == pytest.approx(0.775)
    # the product's meter, fed exact counts that produce those Dice values: TP, P, Y per class
    a = np.array([[5, 1], [5, 2], [5, 2]], dtype=np.int64)       # dice (1.0, 0.5)
    b = np.array([[4, 0], [5, 3], [5, 0]], dtype=np.int64)       # dice (0.8, nan: class 1 absent from the label)
    meter = DiceMeter(2)
    meter.add_counts(np.stack([a, b]))
    means, mm = meter.class_means()
    assert np.allclose(means, cls) and mm == pytest.approx(m)
    assert meter.mean_of_class_means() == pytest.approx(0.70)
    rs = np.random.RandomState(3)
    counts = rs.randint(0, 50, size=(7, 3, 5)).astype(np.int64)
    counts[:, 2, 3] = 0           # class 3 never occurs
    counts[2:5, 2, 1] = 0         # class 1 missing from three volumes
    meter = DiceMeter(5)
    meter.add_counts(counts)
    means, mm = meter.class_means()
    want, want_m = odice.eval_meters(np.stack([odice.dice_from_counts(c) for c in counts]))
    assert np.allclose(means, want, equal_nan=True) and mm == pytest.approx(want_m) and np.isnan(means[3])


# ---------------------------------------------------------------------------------------------------------------------
# Known answers of MONAI 0.8.1's OWN unit tests for the metrics the reference wires in (engine/test.py:28-31), RECALLED
# from memory of tests/test_compute_meandice.py and tests/test_hausdorff_distance.py - unverifiable offline (MONAI is not
# installable here).  Each is also a geometric / arithmetic truth, so a wrong recollection could not make them pass.
# (The erf taps of GaussianFilter have no such pin: the expected array remembered from tests/test_gaussian_filter.py
# reproduces to 4e-8 with the pre-0.4 kernel exp(-x^2/2s^2)/sum, i.e. it predates the erf formula of 0.8 - not used.)
# ---------------------------------------------------------------------------------------------------------------------

def test_monai_compute_meandice_case_1_and_nan_case():
    # TEST_CASE_1: y_pred [[[[1, 0], [0, 1]]]], y [[[[1, 0], [1, 1]]]], include_background=True -> [[0.8]]
    pred = np.array([[1, 0], [0, 1]], dtype=np.uint8)
    y = np.array([[1, 0], [1, 1]], dtype=np.uint8)
    c = odice.dice_counts(pred, y, 2)
    assert odice.dice_from_counts(c)[1] == pytest.approx(0.8)          # 2 * 2 / (3 + 2)
    logits = torch.stack([torch.from_numpy(1.0 - pred), torch.from_numpy(pred * 1.0)]).reshape(2, 1, 2, 2).float()
    md = odice.monai_meandice(logits, torch.from_numpy(y).reshape(1, 1, 2, 2).float(), 2)
    assert md[0, 1].item() == pytest.approx(0.8)
    # TEST_NAN_CASE: an all-zero ground truth channel -> NaN
    empty = odice.dice_counts(pred, np.zeros_like(y), 2)
    assert np.isnan(odice.dice_from_counts(empty)[1])


def _spherical_seg_3d(radius=20.0, centre=(49, 49, 49), value=1, im_shape=(99, 99, 99)):
    """tests/test_hausdorff_distance.py::create_spherical_seg_3d"""
    image = np.zeros(im_shape, dtype=np.int32)
    spy, spx, spz = np.ogrid[-centre[0]:im_shape[0] - centre[0], -centre[1]:im_shape[1] - centre[1],
                             -centre[2]:im_shape[2] - centre[2]]
    image[(spx * spx + spy * spy + spz * spz) <= radius * radius] = value
    return image


def test_monai_hausdorff_sphere_cases():
    from oracle import hausdorff as ohd
    same = ohd.hausdorff_distance(_spherical_seg_3d(), _spherical_seg_3d(), 2, percentile=None)
    assert same[1] == 0.0
    a = _spherical_seg_3d(radius=20, centre=(20, 20, 20))
    b = _spherical_seg_3d(radius=20, centre=(19, 19, 19))
    assert ohd.hausdorff_distance(a, b, 2, percentile=None)[1] == pytest.approx(1.7320508075688772)   # sqrt(3)
    a = _spherical_seg_3d(radius=33, value=2, centre=(19, 33, 22))
    b = _spherical_seg_3d(radius=33, value=2, centre=(20, 33, 22))
    assert ohd.hausdorff_distance(a, b, 3, percentile=None)[2] == 1.0
    assert ohd.hausdorff_distance(a, b, 3, percentile=95)[2] == 1.0
