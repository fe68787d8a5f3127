"""GPU parity for the individual kernels: extraction, finalize, vote, dice, halo add."""
import ctypes as C
import os

import numpy as np
import pytest
import torch

import medicalsemseg_b200 as mss
from medicalsemseg_b200 import _lib, inferer
from medicalsemseg_b200.grid import make_grid
from oracle import dice as odice
from oracle import vote as ovote
from tests.golden.cases import VOTE_CASES, make_vote_maps

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")


def extract_all(vol, roi, overlap, cval, use_tma, batch=3):
    plan = inferer.get_plan(tuple(vol.shape[2:]), roi, overlap, vol.device, vol.shape[0])
    imp = torch.ones(plan.grid.roi, device=vol.device)
    st = inferer.Stitcher(plan, imp, fuse=_lib.FUSE_LOGITS, sw_batch=batch, use_tma=use_tma)
    vol = inferer._tma_ready(vol, plan.grid, cval)  # as the inferer hands it over: W padded to a multiple of 4
    patches, centers = [], []
    for first in range(0, st.total, batch):
        n = min(batch, st.total - first)
        p, c = st.extract(vol, first, n, cval)
        patches.append(p)
        centers.append(c)
    return plan.grid, torch.cat(patches), torch.cat(centers)


@pytest.mark.parametrize("shape,roi,overlap,cval", [
    ((1, 1, 40, 36, 44), 16, 0.25, 0.0),        # W % 4 == 0: TMA path
    ((2, 3, 33, 40, 48), (16, 24, 32), 0.5, 0.0),
    ((1, 1, 128, 112, 104), 96, 0.5, 0.0),      # the 96^3 box geometry, clamped starts 32/16/8
    ((1, 2, 41, 37, 50), (24, 16, 32), 0.5, 0.0),  # W % 4 != 0: plain path
    ((1, 1, 10, 40, 21), 16, 0.25, -0.697),     # padded: plain path with cval
    ((1, 1, 30, 30, 30), (15, 10, 7), 0.3, 0.0),  # roi_w % 4 != 0
    ((1, 2, 40, 36, 52), (16, 16, 32), 0.5, 0.0),  # W starts 0,16,20: all aligned
    ((1, 2, 40, 36, 56), (16, 16, 24), 0.5, 0.0),  # W starts 0,12,24,32
    ((1, 4, 48, 40, 155), (32, 32, 96), 0.5, 0.0),  # BraTS-like: W padded to 156, clamped start 59 -> shift 3
    ((1, 1, 24, 20, 44), (16, 16, 16), 0.5, 0.0),  # W starts 0,8,16,24,28
    ((1, 1, 24, 20, 88), (16, 16, 48), 0.3, 0.0),  # interval 33: W starts 0,33,40 -> shifts 1 and 0
    ((1, 3, 20, 20, 60), (16, 16, 16), 0.6, 0.0),  # interval 6: shifts 0,2 and the clamped 44
    ((1, 1, 20, 20, 64), (16, 16, 20), 0.45, 0.0),  # interval 11: every shift 0..3
])
@pytest.mark.parametrize("use_tma", [1, 0, 2, 3])  # auto / shifted-vector / scalar / volume-stationary rows kernel (where it applies)
def test_extract_matches_slicing(shape, roi, overlap, cval, use_tma):
    rs = np.random.RandomState(11)
    vol = torch.from_numpy(rs.standard_normal(shape).astype(np.float32)).cuda()
    g, patches, centers = extract_all(vol, roi, overlap, cval, use_tma)
    pads = []
    for a in (2, 1, 0):
        diff = g.image_size[a] - g.orig_size[a]
        pads.extend([diff // 2, diff - diff // 2])
    padded = torch.nn.functional.pad(vol, pads, value=cval)
    i = 0
    for b in range(shape[0]):
        for n in range(g.n_windows):
            s = g.window_start(n)
            want = padded[b, :, s[0]:s[0] + g.roi[0], s[1]:s[1] + g.roi[1], s[2]:s[2] + g.roi[2]]
            assert torch.equal(patches[i], want), (b, n, s)
            want_c = torch.tensor(g.centers(n)).float()  # double division rounded once, engine/utils.py:126-130
            assert torch.equal(centers[i].cpu(), want_c)
            i += 1
    assert i == patches.shape[0]


@pytest.mark.parametrize("name", sorted(VOTE_CASES))
def test_majority_vote_bit_exact(name):
    case = VOTE_CASES[name]
    maps = make_vote_maps(case)
    want = ovote.majority_vote(maps, case["k"])
    got = mss.majority_vote([torch.from_numpy(m).cuda() for m in maps], case["k"])
    assert got.dtype == torch.uint8 and np.array_equal(got.cpu().numpy(), want)
    assert np.array_equal(np.load(os.path.join(GOLD, f"vote_{name}.npz"))["voted"], got.cpu().numpy())
    # reference call convention: tuple of float64 arrays (nibabel get_fdata), int64 result
    new = mss.get_new_label(tuple(m.astype(np.float64) for m in maps), len(maps), case["k"])
    assert new.dtype == np.int64 and np.array_equal(new.astype(np.uint8), want)


def test_majority_vote_unaligned_and_large():
    rs = np.random.RandomState(5)
    for n in (1, 15, 17, 1000003):
        maps = [rs.randint(0, 5, size=n + 1).astype(np.uint8) for _ in range(5)]
        dev = [torch.from_numpy(m).cuda()[1:] for m in maps]  # odd base address: scalar path
        want = ovote.majority_vote([m[1:] for m in maps], 5)
        assert np.array_equal(mss.majority_vote(dev, 5).cpu().numpy(), want)
    v = 240 * 240 * 155  # BASELINE.json configs[3] shape
    maps = [rs.randint(0, 3, size=v).astype(np.uint8) for _ in range(5)]
    got = mss.majority_vote([torch.from_numpy(m).cuda() for m in maps], 3).cpu().numpy()
    assert np.array_equal(got, ovote.majority_vote_rule(maps, 3))


@pytest.mark.parametrize("k,n,ldt", [(14, 100003, "u8"), (14, 100003, "f32"), (3, 17, "u8"), (16, 4096, "f32"),
                                     (2, 1 << 22, "u8"), (14, 512 * 512 * 50, "u8")])
def test_dice_counts_bit_exact(k, n, ldt):
    rs = np.random.RandomState(k + n % 97)
    pred = rs.randint(0, k + 2, size=n).astype(np.uint8)   # values >= k must be counted nowhere
    lab = rs.randint(0, k + 1, size=n).astype(np.uint8)
    lab_t = torch.from_numpy(lab).cuda() if ldt == "u8" else torch.from_numpy(lab.astype(np.float32)).cuda()
    got = mss.dice_counts(torch.from_numpy(pred).cuda(), lab_t, k)
    want = odice.dice_counts(pred, lab, k)
    assert got.dtype == torch.int64 and np.array_equal(got.cpu().numpy(), want)
    assert np.allclose(mss.dice_from_counts(got), odice.dice_from_counts(want), equal_nan=True)
    # accumulating into an existing buffer adds
    again = mss.dice_counts(torch.from_numpy(pred).cuda(), lab_t, k, out=got)
    assert np.array_equal(again.cpu().numpy(), 2 * want)


def test_dice_meter_follows_reference_nan_rules():
    k = 4
    meter = mss.DiceMeter(k)
    pred = torch.tensor([0, 1, 1, 2, 2, 2], dtype=torch.uint8).cuda()
    lab = torch.tensor([0, 1, 2, 2, 2, 0], dtype=torch.float32).cuda().reshape(1, 1, 1, 2, 3)
    meter.update(pred, lab)
    meter.update(pred, lab)
    means, m = meter.class_means()
    want, want_m = odice.eval_meters(np.stack([odice.dice_from_counts(odice.dice_counts(pred.cpu().numpy(), lab.cpu().numpy(), k))] * 2))
    assert np.allclose(means, want, equal_nan=True) and m == pytest.approx(want_m)
    assert np.isnan(means[3])


def test_halo_add():
    lib = _lib.load()
    a = torch.randn(6, 40, device="cuda")
    b = torch.randn(6, 48, device="cuda")
    want = a.clone()
    want[:, :32] += b[:, :32]
    rc = lib.mss_halo_add(a.data_ptr(), 40, b.data_ptr(), 48, 6, 32, torch.cuda.current_stream().cuda_stream)
    assert rc == 0 and torch.equal(a, want)
    a2 = torch.randn(5, 7, device="cuda")
    b2 = torch.randn(5, 7, device="cuda")
    want2 = a2 + b2
    assert lib.mss_halo_add(a2.data_ptr(), 7, b2.data_ptr(), 7, 5, 7, torch.cuda.current_stream().cuda_stream) == 0
    assert torch.equal(a2, want2)


def test_logits_to_labels_probs_and_ties():
    rs = np.random.RandomState(3)
    logits = torch.from_numpy(rs.standard_normal((2, 5, 9, 10, 13)).astype(np.float32)).cuda()
    logits[0, 1, 0, 0, 0] = logits[0, 3, 0, 0, 0] = 7.0      # exact tie: first maximum wins
    logits[1, :, 2, 3, 4] = float("nan")                        # NaN rows: softmax is all-NaN, np.argmax answers 0
    logits[1, 3, 4, 4, 4] = float("nan")
    logits[1, 2, 5, 5, 5] = float("inf")
    st = mss.InferStats()
    labels, probs = mss.logits_to_labels(logits, return_probs=True, stats=st)
    want = np.argmax(torch.softmax(logits, 1).cpu().numpy(), axis=1).astype(np.uint8)
    assert np.array_equal(labels.cpu().numpy(), want)
    assert labels[0, 0, 0, 0].item() == 1 and labels[1, 2, 3, 4].item() == 0 and labels[1, 4, 4, 4].item() == 0
    ok = torch.isfinite(logits).all(1, keepdim=True).expand_as(logits)
    assert torch.allclose(probs[ok], torch.softmax(logits, 1)[ok], rtol=1e-5, atol=1e-7)
    assert st.near_ties >= 1


@pytest.mark.parametrize("m", [1, 2, 3, 4, 5, 6, 7, 8, 9, 12, 15])
@pytest.mark.parametrize("k", [2, 3, 14, 16])
def test_majority_vote_all_ensemble_sizes(m, k):
    """M <= 8 runs the bit-sliced kernel, above that the scalar one; labels 16..255 (legal, they match no class)
    force the per-chunk scalar path inside the sliced kernel."""
    rs = np.random.RandomState(1000 * m + k)
    n = 32 * 1031 + 19  # chunks plus a scalar tail
    base = rs.randint(0, k, n).astype(np.uint8)
    maps = []
    for _ in range(m):
        noise = rs.randint(0, 16, n).astype(np.uint8)
        mp = np.where(rs.random_sample(n) < 0.35, noise, base)
        mp[rs.randint(0, n, 40)] = rs.randint(16, 256, 40)  # sprinkle wide labels
        maps.append(np.ascontiguousarray(mp.astype(np.uint8)))
    got = mss.majority_vote([torch.from_numpy(a).cuda() for a in maps], k).cpu().numpy()
    assert np.array_equal(got, ovote.majority_vote(maps, k))


def test_dice_counts_wide_labels_and_uniform():
    rs = np.random.RandomState(9)
    n = 32 * 70001 + 7
    k = 14
    pred = rs.randint(0, k, n).astype(np.uint8)
    lab = np.where(rs.random_sample(n) < 0.7, pred, rs.randint(0, k, n)).astype(np.uint8)
    pred[rs.randint(0, n, 500)] = rs.randint(16, 256, 500)
    lab[rs.randint(0, n, 500)] = rs.randint(16, 256, 500)
    got = mss.dice_counts(torch.from_numpy(pred).cuda(), torch.from_numpy(lab).cuda(), k).cpu().numpy()
    assert np.array_equal(got, odice.dice_counts(pred, lab, k))
    # float labels that are not integers / negative / huge belong to no class
    labf = lab.astype(np.float32)
    labf[::1001] = 2.5
    labf[5::1003] = -1.0
    labf[7::1009] = 1e9
    want_lab = lab.copy()
    want_lab[::1001] = 255
    want_lab[5::1003] = 255
    want_lab[7::1009] = 255
    got = mss.dice_counts(torch.from_numpy(pred).cuda(), torch.from_numpy(labf).cuda(), k).cpu().numpy()
    assert np.array_equal(got, odice.dice_counts(pred, want_lab, k))
    # one class everywhere: every thread's packed counters see the maximum load
    full = torch.full((512 * 512 * 200,), 3, dtype=torch.uint8, device="cuda")
    got = mss.dice_counts(full, full, k).cpu().numpy()
    assert got[0, 3] == full.numel() and got[1, 3] == full.numel() and got[2, 3] == full.numel() and got.sum() == 3 * full.numel()


def test_extract_ahead_equals_per_batch():
    rs = np.random.RandomState(2)
    vol = torch.from_numpy(rs.standard_normal((2, 2, 40, 36, 44)).astype(np.float32)).cuda()
    plan = inferer.get_plan((40, 36, 44), 16, 0.5, vol.device, 2)
    imp = torch.ones(plan.grid.roi, device=vol.device)
    per_window = 4 * 2 * 16 ** 3
    got = {}
    for budget in (per_window, 7 * per_window, 1 << 30):  # one batch per launch, ragged groups, everything at once
        st = inferer.Stitcher(plan, imp, fuse=_lib.FUSE_LOGITS, sw_batch=3, extract_bytes=budget)
        ps, cs, firsts = [], [], []
        for first, n, p, c in st.batches(vol, 0.0):
            assert p.is_contiguous() and p.shape[0] == n
            ps.append(p.clone()), cs.append(c.clone()), firsts.append(first)
        assert firsts == list(range(0, st.total, 3))
        got[budget] = (torch.cat(ps), torch.cat(cs))
    ref = got[per_window]
    for budget, (p, c) in got.items():
        assert torch.equal(p, ref[0]) and torch.equal(c, ref[1])


from oracle import resample as oresample  # noqa: E402
from tests.golden.cases import RESAMPLE_CASES, make_label_map  # noqa: E402


@pytest.mark.parametrize("name", sorted(RESAMPLE_CASES))
def test_resample_3d_bit_exact(name):
    from medicalsemseg_b200.resample import resample_3d
    case = RESAMPLE_CASES[name]
    img = make_label_map(case)
    got = resample_3d(torch.from_numpy(img).cuda(), case["target"]).cpu().numpy()
    assert np.array_equal(got, np.load(os.path.join(GOLD, f"resample_{name}.npz"))["out"])
    assert np.array_equal(got, oresample.resample_3d(img, case["target"]))


@pytest.mark.parametrize("oz", [3, 5, 15, 16, 17, 20, 30, 31, 100, 147, 257, 513, 1030])
def test_resample_3d_output_widths(oz):
    """Every row length class of the stream kernel (chunk size 4 / 8 / 16, rows per period 1 .. 16, rows shorter than a
    chunk, periods longer than a CTA -> rows kernel) vs scipy's zoom, incl. sizes where scipy writes its constant 0."""
    from medicalsemseg_b200.resample import resample_3d
    rs = np.random.RandomState(oz)
    for in_shape, tgt in (((7, 9, max(2, int(oz * 1.37))), (11, 5, oz)), ((5, 6, max(2, oz // 2 + 1)), (5, 13, oz)),
                          ((3, 4, oz), (9, 4, oz))):
        img = rs.randint(1, 14, in_shape).astype(np.uint8)
        got = resample_3d(torch.from_numpy(img).cuda(), tgt).cpu().numpy()
        assert np.array_equal(got, oresample.resample_3d(img, tgt)), (in_shape, tgt)
    imgs = rs.randint(1, 14, (3, 6, 5, 23)).astype(np.uint8)  # a batch: rows cross volume boundaries inside a period
    got = resample_3d(torch.from_numpy(imgs).cuda(), (4, 7, oz)).cpu().numpy()
    for b in range(3):
        assert np.array_equal(got[b], oresample.resample_3d(imgs[b], (4, 7, oz)))


def test_resample_3d_batched_and_large():
    from medicalsemseg_b200.resample import resample_3d
    rs = np.random.RandomState(77)
    imgs = rs.randint(0, 14, (3, 37, 41, 29)).astype(np.uint8)
    got = resample_3d(torch.from_numpy(imgs).cuda(), (64, 30, 48)).cpu().numpy()
    for b in range(3):
        assert np.array_equal(got[b], oresample.resample_3d(imgs[b], (64, 30, 48)))
    big = rs.randint(0, 14, (200, 200, 96)).astype(np.uint8)   # BTCV-like: resampled grid back to the original one
    got = resample_3d(torch.from_numpy(big).cuda(), (512, 512, 147)).cpu().numpy()
    assert np.array_equal(got, oresample.resample_3d(big, (512, 512, 147)))


from oracle import transforms as otransforms  # noqa: E402
from tests.golden.cases import INTENSITY_CASES, make_ct_volume  # noqa: E402


def _ulp_close(got, want, ulps=1):
    got, want = np.asarray(got, np.float32), np.asarray(want, np.float32)
    tol = ulps * np.spacing(np.maximum(np.abs(want), np.float32(1e-30)).astype(np.float32))
    return bool(np.all(np.abs(got.astype(np.float64) - want.astype(np.float64)) <= tol))


@pytest.mark.parametrize("name", sorted(INTENSITY_CASES))
def test_cubed_intensity_scaler_vs_reference_fixture(name):
    from medicalsemseg_b200 import transforms as T
    c = INTENSITY_CASES[name]
    fx = np.load(os.path.join(GOLD, f"intensity_{name}.npz"))["out"]
    vol = torch.from_numpy(make_ct_volume(c)).cuda()
    got = T.scale_intensity_range(vol, c["a_min"], c["a_max"], c["b_min"], c["b_max"], c["clip"], cubed=True,
                                  float64=True).cpu().numpy()
    # float64 intermediates like the reference under NumPy >= 2; cbrtf (CUDA) vs cbrtf (glibc) may differ by 1 ulp
    assert np.max(np.abs(got - fx)) <= 4e-7 * max(1.0, float(np.abs(fx).max()))  # 1 ulp of cbrtf, scaled
    assert np.mean(got == fx) > 0.5
    got32 = T.scale_intensity_range(vol, c["a_min"], c["a_max"], c["b_min"], c["b_max"], c["clip"], cubed=True,
                                    float64=False).cpu().numpy()
    assert np.max(np.abs(got32 - fx)) <= 6e-7 * max(1.0, float(np.abs(fx).max()))


@pytest.mark.parametrize("a_min,a_max,b_min,b_max,clip", [(-1000, 1000, 0.0, 1.0, True), (-175, 250, 0.0, 1.0, True),
                                                          (-500, 1500, -1.0, 1.0, False), (0, 1, None, None, False),
                                                          (5, 5, 0.5, 1.0, True), (-1000, 1000, None, 1.0, True)])
def test_fixed_range_scaler_and_normalise_bit_exact(a_min, a_max, b_min, b_max, clip):
    from medicalsemseg_b200 import transforms as T
    rs = np.random.RandomState(3)
    vol = (rs.standard_normal((2, 11, 13, 17)) * 700 - 200).astype(np.float32)
    vol[0, 0, 0, :5] = [np.nan, np.inf, -np.inf, 0.0, -0.0]
    want = otransforms.scale_intensity_range(vol, a_min, a_max, b_min, b_max, clip)
    got = T.scale_intensity_range(torch.from_numpy(vol).cuda(), a_min, a_max, b_min, b_max, clip).cpu().numpy()
    assert np.array_equal(got, want, equal_nan=True)
    want_n = otransforms.normalize_intensity(want, 0.1943, 0.2786)
    got_n = T.normalize_intensity(torch.from_numpy(want).cuda(), 0.1943, 0.2786).cpu().numpy()
    assert np.array_equal(got_n, want_n, equal_nan=True)


def test_statistics_based_intensity_transforms():
    from medicalsemseg_b200 import transforms as T
    rs = np.random.RandomState(4)
    vol = (rs.standard_normal((2, 16, 18, 20)) * 300).astype(np.float32)
    vol[rs.random_sample(vol.shape) < 0.3] = 0.0
    dv = torch.from_numpy(vol).cuda()
    got = T.normalize_intensity(dv, nonzero=True, channel_wise=True).cpu().numpy()
    want = otransforms.normalize_intensity(vol, nonzero=True, channel_wise=True)
    assert np.allclose(got, want, rtol=1e-5, atol=1e-5) and np.array_equal(got == 0, want == 0)
    for q in (5, 50, 95, 99.5):
        assert T.percentile(dv, q) == pytest.approx(float(np.percentile(vol, q)), rel=1e-6, abs=1e-6)
    got = T.scale_intensity_range_percentiles(dv, 5, 95, 0.0, 1.0, clip=True).cpu().numpy()
    want = otransforms.scale_intensity_range_percentiles(vol, 5, 95, 0.0, 1.0, clip=True)
    assert np.allclose(got, want, rtol=1e-5, atol=1e-6)


def test_test_time_intensity_chain_is_one_launch_equivalent():
    from types import SimpleNamespace
    from medicalsemseg_b200 import transforms as T
    rs = np.random.RandomState(5)
    vol = (rs.standard_normal((1, 20, 20, 24)) * 700 - 200).astype(np.float32)
    dv = torch.from_numpy(vol).cuda()
    cfg = SimpleNamespace(t_fixed_ct_intensity=True, t_ct_min=-1000, t_ct_max=1000, t_normalize=True, t_norm_mean=0.1943,
                          t_norm_std=0.2786)  # data/dataset_builder.py:333-368 with utils/arguments.py defaults
    want = otransforms.normalize_intensity(otransforms.scale_intensity_range(vol, -1000, 1000, 0.0, 1.0, True), 0.1943, 0.2786)
    assert np.array_equal(T.test_time_intensity(dv, cfg).cpu().numpy(), want)
    cfg2 = SimpleNamespace(t_cubed_ct_intensity=True, t_ct_min=-1000, t_ct_max=1000, t_normalize=True, t_norm_mean=0.1943,
                           t_norm_std=0.2786)
    want2 = otransforms.normalize_intensity(
        otransforms.scale_cubed_intensity_range(vol, -1000, 1000, 0.0, 1.0, True, dtype=np.float64), 0.1943, 0.2786)
    got2 = T.test_time_intensity(dv, cfg2).cpu().numpy()
    assert np.allclose(got2, want2, rtol=0, atol=1e-6)
    assert T.test_time_intensity(dv, SimpleNamespace()) is dv


def test_halo_add_nd_strided_boxes():
    """One launch adds any (strided) face of an accumulator: the boxes block.exchange_halos hands over."""
    from medicalsemseg_b200.slab import cuda_halo_add
    acc = torch.randn(2, 3, 10, 12, 16, device="cuda")
    other = torch.randn(2, 3, 10, 12, 16, device="cuda")
    for box in [(slice(None), slice(None), slice(2, 9), slice(3, 8), slice(4, 12)),     # interior box, W offset aligned
                (slice(None), slice(None), slice(0, 10), slice(0, 12), slice(5, 11)),    # unaligned W range
                (slice(0, 1), slice(1, 3), slice(7, 10), slice(None), slice(None)),      # a D face
                (slice(None), slice(None), slice(None), slice(8, 12), slice(0, 16))]:    # an H face
        a = acc.clone()
        want = a.clone()
        want[box] += other[box]
        cuda_halo_add(a[box], other[box])              # strided source (what a peer read looks like)
        assert torch.equal(a, want)
        b = acc.clone()
        cuda_halo_add(b[box], other[box].contiguous())  # contiguous source (what NCCL delivers)
        assert torch.equal(b, want)
    v = torch.randn(40, device="cuda")
    w = torch.randn(40, device="cuda")
    want = v + w
    cuda_halo_add(v, w)
    assert torch.equal(v, want)


def _blobby(rs, shape, k):
    from scipy import ndimage
    f = ndimage.gaussian_filter(rs.standard_normal(shape), 2.5)
    return np.digitize(f, np.quantile(f, np.linspace(0, 1, k + 1)[1:-1])).astype(np.uint8)


@pytest.mark.parametrize("shape,k,seed", [((24, 28, 20), 4, 0), ((40, 33, 37), 6, 1), ((16, 50, 12), 3, 2), ((31, 17, 64), 5, 3)])
def test_hausdorff95_bit_exact_vs_monai_restated(shape, k, seed):
    """Surface kernel + exact EDT passes + NumPy percentile vs oracle.hausdorff (MONAI 0.8 restated with scipy):
    integer squared distances make the float64 result bit-identical."""
    from oracle import hausdorff as oh
    rs = np.random.RandomState(seed)
    gt = _blobby(rs, shape, k)
    pred = gt.copy()
    flip = rs.random_sample(shape) < 0.06
    pred[flip] = rs.randint(0, k, int(flip.sum()))
    pred[pred == k - 1] = 0                       # a class missing from the prediction: inf / nan by NumPy's rules
    want = oh.hausdorff_distance(pred, gt, k + 1)  # class k is absent from both maps: NaN
    got = mss.hausdorff_distance(torch.from_numpy(pred).cuda(), torch.from_numpy(gt).cuda(), k + 1)
    assert got.dtype == np.float64 and np.array_equal(got, want, equal_nan=True)
    assert np.isnan(got[k])
    # the plain maximum (percentile=None) and the directed variant follow the same code path
    for kw in (dict(percentile=None), dict(directed=True), dict(include_background=False)):
        assert np.array_equal(mss.hausdorff_distance(torch.from_numpy(pred).cuda(), torch.from_numpy(gt).cuda(), k + 1, **kw),
                              oh.hausdorff_distance(pred, gt, k + 1, **kw), equal_nan=True)
    assert mss.mean_hausdorff(got) == pytest.approx(oh.mean_hausdorff(want), nan_ok=True)
    # float32 labels as the reference's loaders deliver them
    gotf = mss.hausdorff_distance(torch.from_numpy(pred).cuda(), torch.from_numpy(gt.astype(np.float32)).cuda()[None, None], k + 1)
    assert np.array_equal(gotf, want, equal_nan=True)


@pytest.mark.parametrize("shape", [(12, 14, 37), (9, 20, 64), (6, 7, 155), (5, 33, 8), (20, 5, 19)])
def test_mask_edges_packed_kernel_vs_erosion(shape):
    """mss_mask_edges (8 voxels per thread on packed bytes, word loads at every alignment) vs binary_erosion XOR mask on
    the cropped box (scipy, border_value 0, thin axes squeezed away - MONAI get_mask_edges): whole volumes, boxes at odd
    offsets, boxes one voxel thick, rows shorter than a run."""
    from scipy import ndimage
    from medicalsemseg_b200.hausdorff import _edges
    rs = np.random.RandomState(sum(shape))
    lab = _blobby(rs, shape, 3)
    lab[rs.random_sample(shape) < 0.05] = 2
    dev = torch.from_numpy(lab).cuda()
    boxes = [((0, 0, 0), shape)]
    for _ in range(12):
        lo = tuple(int(rs.randint(0, n)) for n in shape)
        hi = tuple(int(rs.randint(l + 1, n + 1)) for l, n in zip(lo, shape))
        boxes.append((lo, hi))
    boxes.append(((1, 1, 1), (2, shape[1], shape[2])))             # one plane
    boxes.append(((0, 2, 3), (shape[0], 3, shape[2] - 1)))         # one row per plane
    boxes.append(((0, 0, shape[2] - 1), (shape[0], shape[1], shape[2])))  # one column
    for lo, hi in boxes:
        for c in (0, 1, 2):
            m = lab[lo[0]:hi[0], lo[1]:hi[1], lo[2]:hi[2]] == c
            sq = np.squeeze(m)
            if sq.ndim == 0:
                continue  # 0-d erosion: not defined by the reference
            want = (ndimage.binary_erosion(sq) ^ sq).reshape(m.shape)
            got = _edges(dev, c, lo, hi).cpu().numpy()
            assert np.array_equal(got.astype(bool), want) and got.max() <= 1, (lo, hi, c)


def test_squared_edt_matches_scipy_and_properties():
    from scipy import ndimage
    from medicalsemseg_b200.hausdorff import squared_edt
    rs = np.random.RandomState(4)
    for shape, dens in [((33, 20, 41), 0.01), ((8, 64, 8), 0.2), ((50, 50, 50), 0.0005)]:
        feat = rs.random_sample(shape) < dens
        h0 = torch.from_numpy(np.where(feat, 0, 1 << 29).astype(np.int32)).cuda()
        got = squared_edt(h0).cpu().numpy()
        want = ndimage.distance_transform_edt(~feat)
        assert np.array_equal(np.sqrt(got.astype(np.float64)), want)
        got8 = squared_edt(torch.from_numpy(feat.astype(np.uint8)).cuda()).cpu().numpy()  # row scan from the mask first
        assert np.array_equal(got8, got)
    # identical surfaces: Hausdorff distance 0 for every present class
    lab = torch.from_numpy(_blobby(rs, (30, 30, 30), 4)).cuda()
    assert np.array_equal(mss.hausdorff_distance(lab, lab, 4), np.zeros(4))


@pytest.mark.parametrize("shape,k", [((1, 14, 20, 24, 28), 14), ((2, 3, 16, 18, 21), 3), ((1, 5, 9, 10, 13), 5)])
def test_dice_ce_loss_matches_monai_restated(shape, k):
    """Fused softmax + Dice/CE sums vs oracle.losses (MONAI DiceCELoss restated with torch on the CPU): a float
    reduction, so the bar is a relative tolerance (1e-5), not bit equality."""
    from medicalsemseg_b200 import losses as L
    from oracle import losses as ol
    rs = np.random.RandomState(k)
    logits = torch.from_numpy((rs.standard_normal(shape) * 3).astype(np.float32))
    labels = torch.from_numpy(rs.randint(0, k, (shape[0], 1) + shape[2:]).astype(np.float32))
    want, parts = ol.dice_ce_loss(logits, labels)
    got, gparts = L.dice_ce_loss(logits.cuda(), labels.cuda())
    assert got == pytest.approx(want, rel=1e-5) and gparts["dice"] == pytest.approx(parts["dice"], rel=1e-5)
    assert gparts["ce"] == pytest.approx(parts["ce"], rel=1e-5)
    got8, _ = L.dice_ce_loss(logits.cuda(), labels.to(torch.uint8).cuda())       # uint8 label maps
    assert got8 == pytest.approx(want, rel=1e-5)
    for kw in (dict(squared_pred=False), dict(include_background=False), dict(lambda_dice=0.5, lambda_ce=2.0)):
        assert L.dice_ce_loss(logits.cuda(), labels.cuda(), **kw)[0] == pytest.approx(ol.dice_ce_loss(logits, labels, **kw)[0], rel=1e-5)
    # on the stitcher's W-pitched logits view (no copy): W = 13 lives in a pitch of 16
    if shape[-1] % 4:
        buf = torch.zeros(shape[:-1] + ((shape[-1] + 3) // 4 * 4,), device="cuda")
        buf[..., :shape[-1]] = logits.cuda()
        assert L.dice_ce_loss(buf[..., :shape[-1]], labels.cuda())[0] == pytest.approx(want, rel=1e-5)


def test_dice_ce_loss_full_size_identities():
    from medicalsemseg_b200 import losses as L
    k, d, h, w = 14, 128, 128, 200
    gen = torch.Generator(device="cuda").manual_seed(3)
    labels = torch.randint(0, k, (1, 1, d, h, w), device="cuda", generator=gen, dtype=torch.uint8)
    onehot = torch.nn.functional.one_hot(labels[0, 0].long(), k).permute(3, 0, 1, 2)[None].float()
    sure = onehot * 40.0   # softmax saturates to the one-hot: Dice -> 0 (up to the smoothing), CE -> 0
    loss, parts = L.dice_ce_loss(sure, labels)
    assert abs(parts["dice"]) < 1e-6 and abs(parts["ce"]) < 1e-6 and abs(loss) < 2e-6
    flat = torch.zeros_like(sure)  # uniform softmax 1/K: CE = log K exactly
    _, parts = L.dice_ce_loss(flat, labels)
    assert parts["ce"] == pytest.approx(np.log(k), rel=1e-6)
    s = L.dice_ce_sums(flat, labels)[0]
    assert s[2 * k:3 * k].sum() == d * h * w and s[k:2 * k] == pytest.approx(np.full(k, d * h * w / k ** 2), rel=1e-6)


@pytest.mark.parametrize("ldt", ["u8", "f32"])
@pytest.mark.parametrize("n", [100003, 32 * 4096, 17])
def test_dice_counts_batched_equals_per_volume(ldt, n):
    """One launch for a batch of label maps gives the per-volume counts of separate launches (odd n: the volumes of the
    batch lose their 16-byte alignment and the scalar path must take over)."""
    rs = np.random.RandomState(n % 1000)
    b, k = 5, 14
    pred = rs.randint(0, k + 2, size=(b, n)).astype(np.uint8)
    lab = rs.randint(0, k + 1, size=(b, n)).astype(np.uint8)
    lab_t = torch.from_numpy(lab).cuda() if ldt == "u8" else torch.from_numpy(lab.astype(np.float32)).cuda()
    got = mss.dice_counts_batched(torch.from_numpy(pred).cuda(), lab_t, k).cpu().numpy()
    for i in range(b):
        assert np.array_equal(got[i], odice.dice_counts(pred[i], lab[i], k)), i
    meter, single = mss.DiceMeter(k), mss.DiceMeter(k)
    meter.update_batch(torch.from_numpy(pred).cuda(), lab_t)
    for i in range(b):
        single.update(torch.from_numpy(pred[i]).cuda(), lab_t[i])
    assert np.allclose(meter.class_means()[0], single.class_means()[0], equal_nan=True)


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs in one process")
def test_two_devices_in_one_process():
    """The dynamic shared-memory opt-in of the accumulate / extract / resample kernels is per device: stitching on cuda:0
    and then on cuda:1 from the same process must work (round-1 advisor finding: a process-wide `static bool`)."""
    from oracle import sliding_window as osw
    from oracle.predictors import ArithmeticPredictor
    rs = np.random.RandomState(3)
    vol = torch.from_numpy(rs.standard_normal((1, 1, 40, 36, 44)).astype(np.float32))
    pred = ArithmeticPredictor(5)
    ref = osw.sliding_window_inference(vol, None, (16, 16, 16), 4, pred, overlap=0.5, mode="gaussian", tuple_input=False)
    for dev in ("cuda:0", "cuda:1"):
        out = mss.sliding_window_inference(vol.to(dev), None, (16, 16, 16), 4, pred, overlap=0.5, mode="gaussian",
                                           mss_tuple_input=False)
        assert out.device == torch.device(dev) and torch.equal(out.cpu(), ref)
        lab = mss.resample_3d(mss.sliding_window_infer(vol.to(dev), pred, (16, 16, 16), 0.5, "gaussian")[0], (50, 30, 61))
        assert lab.device == torch.device(dev)
