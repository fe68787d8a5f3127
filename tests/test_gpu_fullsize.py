"""Size-independent properties at BASELINE.json's FULL sizes (the oracle would need minutes to hours there):
partition of unity of the normalised blend, linearity in the predictor output, fused == unfused, launch-split
invariance, vote idempotence / permutation invariance, Dice count identities, resample identities.

The predictors are cheap elementwise maps of the patch, so the stitched result has a closed form in terms of the input
volume - any mis-addressed window, missed voxel or wrong weight count shows up as a deviation from it."""
import numpy as np
import pytest
import torch

import medicalsemseg_b200 as mss

pytestmark = pytest.mark.gpu

CFG2 = dict(shape=(1, 1, 512, 512, 200), k=14, n_windows=400)      # BASELINE.json configs[1]
CFG3 = dict(shape=(1, 1, 512, 512, 1024), k=2, n_windows=2100)     # configs[2] geometry (K cut to 2 to bound memory)
CFG4 = dict(shape=(1, 4, 240, 240, 155), k=3, n_windows=48)        # configs[3]


class AffineOfPatch:
    """logits[:, c] = a_c * (sum of the patch's channels) + b_c: every window that covers a voxel predicts the same
    value there, so the normalised blend must return exactly that value (up to the rounding of sum(w*x)/sum(w))."""

    def __init__(self, k):
        self.a = [0.25 * (c + 1) * (-1) ** c for c in range(k)]
        self.b = [0.125 * c - 0.5 for c in range(k)]

    def __call__(self, x):
        s = x.sum(dim=1)
        return torch.stack([s * a + b for a, b in zip(self.a, self.b)], dim=1)

    def expected(self, vol):
        s = vol.sum(dim=1)
        return torch.stack([s * a + b for a, b in zip(self.a, self.b)], dim=1)


@pytest.mark.parametrize("cfg", [CFG2, CFG4, CFG3], ids=["cfg2", "cfg4", "cfg3_geometry"])
def test_blend_is_a_partition_of_unity_at_full_size(cfg):
    gen = torch.Generator(device="cuda").manual_seed(5)
    vol = torch.randn(cfg["shape"], device="cuda", generator=gen)
    pred = AffineOfPatch(cfg["k"])
    st = mss.InferStats()
    out = mss.sliding_window_inference(vol, None, 96, 8, pred, overlap=0.5, mode="gaussian", mss_tuple_input=False, mss_stats=st)
    assert st.n_windows == cfg["n_windows"]
    want = pred.expected(vol)
    assert out.shape == want.shape
    err = (out - want).abs().max().item()
    assert err <= 2e-5 * max(1.0, want.abs().max().item()), err
    # fused labels == argmax of the closed form wherever the top-2 gap is not tiny
    labels = mss.sliding_window_infer(vol, pred, 96, 0.5, "gaussian", sw_batch_size=8)
    top2 = want.topk(2, dim=1).values
    clear = (top2[:, 0] - top2[:, 1]) > 1e-4 * top2.abs().amax(dim=1).clamp_min(1e-6)
    assert torch.equal(labels[clear], want.argmax(dim=1).to(torch.uint8)[clear])
    assert clear.float().mean().item() > 0.99
    del out, want, labels
    torch.cuda.empty_cache()


def test_launch_split_and_fusion_invariance_cfg2():
    """The same 400 windows applied by 1 launch or by many (accumulator read-modify-write) give bit-identical logits;
    the fused label path equals the argmax of those logits."""
    gen = torch.Generator(device="cuda").manual_seed(6)
    vol = torch.randn(CFG2["shape"], device="cuda", generator=gen)

    def noisy(x):  # window-dependent logits: position-hashed, so overlapping windows genuinely disagree
        s = x[:, 0]
        return torch.stack([torch.sin(s * (1.0 + 0.37 * c)) + 0.01 * c + s.mean(dim=(1, 2, 3), keepdim=True) * 0.1 * c
                            for c in range(5)], dim=1)

    one = mss.sliding_window_inference(vol, None, 96, 8, noisy, overlap=0.5, mode="gaussian", mss_tuple_input=False)
    st = mss.InferStats()
    many = mss.sliding_window_inference(vol, None, 96, 8, noisy, overlap=0.5, mode="gaussian", mss_tuple_input=False,
                                        mss_group_bytes=1 << 30, mss_stats=st)
    assert st.n_accumulate_calls > 3
    assert torch.equal(one, many)
    st2 = mss.InferStats()
    labels = mss.sliding_window_infer(vol, noisy, 96, 0.5, "gaussian", sw_batch_size=8, stats=st2)
    want = one.argmax(dim=1).to(torch.uint8)
    mism = (labels != want)
    assert int(mism.sum()) <= st2.near_ties  # only counted near-ties may differ (raw sums vs divided sums)
    assert not st2.accumulator_allocated     # the fused path never materialised the fp32 volume


def test_vote_properties_full_size():
    v = 240 * 240 * 155
    gen = torch.Generator(device="cuda").manual_seed(7)
    base = torch.randint(0, 3, (v,), device="cuda", generator=gen, dtype=torch.uint8)
    maps = [torch.where(torch.rand(v, device="cuda", generator=gen) < 0.2,
                        torch.randint(0, 3, (v,), device="cuda", generator=gen, dtype=torch.uint8), base) for _ in range(5)]
    voted = mss.majority_vote(maps, 3)
    for perm in ([4, 3, 2, 1, 0], [2, 0, 4, 1, 3]):   # the vote does not depend on the order of the folds
        assert torch.equal(mss.majority_vote([maps[i] for i in perm], 3), voted)
    assert torch.equal(mss.majority_vote([base] * 5, 3), base)          # unanimous ensembles return the map itself
    assert int(mss.majority_vote([base], 3).sum()) == 0                  # a single fold never out-votes background
    stack = torch.stack(maps)
    for c in (1, 2):  # a class with >= 3 of 5 votes always wins
        sure = (stack == c).sum(0) >= 3
        assert torch.all(voted[sure] == c)
    assert torch.all(voted[(stack == 0).sum(0) >= 4] == 0)


def test_dice_count_identities_full_size():
    v, k = 512 * 512 * 200, 14
    gen = torch.Generator(device="cuda").manual_seed(8)
    pred = torch.randint(0, k, (v,), device="cuda", generator=gen, dtype=torch.uint8)
    lab = torch.where(torch.rand(v, device="cuda", generator=gen) < 0.6, pred,
                      torch.randint(0, k, (v,), device="cuda", generator=gen, dtype=torch.uint8))
    c = mss.dice_counts(pred, lab, k).cpu().numpy()
    assert c[1].sum() == v and c[2].sum() == v and np.all(c[0] <= np.minimum(c[1], c[2]))
    assert c[0].sum() == int((pred == lab).sum())
    assert np.array_equal(c[1], torch.bincount(pred.long(), minlength=k).cpu().numpy())
    assert np.array_equal(c[2], torch.bincount(lab.long(), minlength=k).cpu().numpy())
    same = mss.dice_counts(pred, pred, k).cpu().numpy()
    assert np.array_equal(same[0], same[1]) and np.array_equal(same[1], same[2])
    assert np.allclose(mss.dice_from_counts(same), 1.0)
    cf = mss.dice_counts(pred, lab.float(), k).cpu().numpy()     # float32 labels (the reference's loaders) count the same
    assert np.array_equal(cf, c)


def test_resample_identities_full_size():
    from medicalsemseg_b200.resample import resample_3d, zoom_index_table
    gen = torch.Generator(device="cuda").manual_seed(9)
    lab = torch.randint(0, 14, (512, 512, 200), device="cuda", generator=gen, dtype=torch.uint8)
    assert torch.equal(resample_3d(lab, (512, 512, 200)), lab)                       # identity zoom
    up = resample_3d(lab, (512, 512, 400))
    iz = torch.from_numpy(zoom_index_table(200, 400).astype(np.int64)).cuda()
    want = lab[:, :, iz.clamp_min(0)] * (iz >= 0).to(torch.uint8)
    assert torch.equal(up, want)                                                      # separable index rule along z
    down = resample_3d(lab, (256, 300, 147))
    ix, iy, iz = (torch.from_numpy(zoom_index_table(a, b).astype(np.int64)).cuda() for a, b in ((512, 256), (512, 300), (200, 147)))
    want = lab[ix.clamp_min(0)][:, iy.clamp_min(0)][:, :, iz.clamp_min(0)]
    want = want * ((ix >= 0)[:, None, None] & (iy >= 0)[None, :, None] & (iz >= 0)[None, None, :]).to(torch.uint8)
    assert torch.equal(down, want)


# ---------------------------------------------------------------------------------------------------------------------
# BASELINE.json configs at their real geometry against outputs of the reference's own engine/utils.py (tests/golden/
# make_golden.py --sections fullsize, run in the build container): bit-identical stitched logits, labels identical
# except counted near-ties
# ---------------------------------------------------------------------------------------------------------------------
import hashlib
import json
import os

from oracle.predictors import ArithmeticPredictor
from tests.golden.cases import FULLSIZE_LABEL_STRIDE, FULLSIZE_SAMPLE_STRIDE, FULLSIZE_SW_CASES
from tests.gpu_helpers import cuda_inputs

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _sha(t):
    return hashlib.sha256(np.ascontiguousarray(t.contiguous().cpu().numpy()).tobytes()).hexdigest()


@pytest.mark.parametrize("name", sorted(FULLSIZE_SW_CASES))
def test_fullsize_geometry_bit_identical_to_the_reference_run(name):
    c = FULLSIZE_SW_CASES[name]
    gold = json.load(open(os.path.join(GOLD, "manifest.json")))["fullsize"][name]
    fx = np.load(os.path.join(GOLD, f"sw_{name}.npz"))
    vol, affine = cuda_inputs(c)
    k = c["k"]

    # (1) one accumulate launch over all windows (fused normalise), the reference's signature and predictor convention
    pred = ArithmeticPredictor(k)
    st = mss.InferStats()
    out = mss.sliding_window_inference(vol, affine, c["roi"], c["sw_batch"], pred, overlap=c["overlap"], mode=c["mode"],
                                       mss_stats=st)
    assert list(out.shape) == gold["shape"] and len(pred.calls) == gold["n_predictor_calls"]
    sample = out.contiguous().view(-1)[::FULLSIZE_SAMPLE_STRIDE].cpu().numpy()
    assert np.array_equal(sample, fx["sample"])
    assert _sha(out) == gold["sha256"], "stitched logits differ from the reference run"

    # (2) many launches (accumulator read-modify-write between them): bit-identical to (1), hence to the reference
    st2 = mss.InferStats()
    out2 = mss.sliding_window_inference(vol, affine, c["roi"], c["sw_batch"], ArithmeticPredictor(k), overlap=c["overlap"],
                                        mode=c["mode"], mss_stats=st2,
                                        mss_group_bytes=min(1 << 30, 4 * k * 96 ** 3 * st.n_windows // 4))
    assert st2.n_accumulate_calls > st.n_accumulate_calls
    assert torch.equal(out, out2)
    del out2

    # (3) fused labels (the accumulate kernel's argmax of raw sums) vs the reference's softmax -> np.argmax -> uint8:
    # identical except voxels whose top-2 gap is below 1e-5 (counted); the reference labels at every differing voxel are
    # checked through the stored strided sample and the near-tie mask computed from the (bit-identical) logits
    st3 = mss.InferStats()
    labels = mss.sliding_window_infer(vol, ArithmeticPredictor(k), c["roi"], c["overlap"], c["mode"], sw_batch_size=c["sw_batch"],
                                      affine=affine, stats=st3)[0]
    assert st3.accumulator_allocated == (st3.n_accumulate_calls > 1)  # one launch never materialises the fp32 volume
    # the reference's labels = plain first-max argmax of the (bit-identical) logits, except the handful of voxels where float32
    # softmax rounding merges a near-tie (stored by make_golden.py: index + the reference's label there)
    ref = out[0].argmax(dim=0).to(torch.uint8)
    idx = torch.from_numpy(fx["softmax_diff_index"]).cuda()
    assert idx.numel() == gold["softmax_vs_plain_argmax_mismatch"] <= 4
    plain = ref.clone()
    ref.view(-1)[idx] = torch.from_numpy(fx["softmax_diff_label"]).cuda()
    assert _sha(ref) == gold["labels_sha256"] and _sha(plain) == gold["labels_plain_argmax_sha256"]
    assert np.array_equal(ref.view(-1)[::FULLSIZE_LABEL_STRIDE].cpu().numpy(), fx["labels_sample"])
    bad = labels != ref
    n_bad = int(bad.sum())
    if n_bad:
        top = out[0].topk(2, dim=0).values
        gap = (top[0] - top[1]) / torch.maximum(top[0].abs(), top[1].abs()).clamp_min(1e-37)
        assert bool((gap[bad] < 1e-5).all()), "label mismatch outside the near-tie tolerance"
        assert n_bad <= st3.near_ties
    else:
        assert _sha(labels) == gold["labels_sha256"]
    # (4) labels from the stand-alone finalize kernel on the stitched logits: the same rule on divided sums
    again = mss.logits_to_labels(out.contiguous())[0]
    assert torch.equal(again, plain)
    del out, labels, plain, again, ref
    torch.cuda.empty_cache()
