"""GPU parity: the CUDA stitching path (through the C ABI) against the oracle on the same seeded inputs,
and against the fixtures the reference itself produced (tests/golden)."""
import json
import os

import numpy as np
import pytest
import torch

import medicalsemseg_b200 as mss
from medicalsemseg_b200 import importance as I
from oracle import monai08 as M
from oracle import sliding_window as osw
from oracle.predictors import ArithmeticPredictor
from tests.golden.cases import SW_CASES, make_volume
from tests.gpu_helpers import TOL, assert_labels_match, cuda_inputs, oracle_run, rel_err

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")
SMALL = [n for n in sorted(SW_CASES) if n != "cfg1_geometry"]


@pytest.mark.parametrize("roi", [(96, 96, 96), (16, 16, 16), (24, 16, 32), (8, 12, 20), (5, 7, 9)])
def test_importance_map_bit_exact_with_host_taps(roi):
    want = M.compute_importance_map(roi, mode="gaussian", sigma_scale=0.125)
    got = I.importance_map(roi, "gaussian", 0.125, "cuda").cpu()
    assert torch.equal(got, want)
    ones = I.importance_map(roi, "constant", 0.125, "cuda").cpu()
    assert torch.equal(ones, torch.ones(roi))


def test_importance_map_device_taps_and_v12():
    roi = (96, 96, 96)
    want = M.compute_importance_map(roi, mode="gaussian", sigma_scale=0.125)
    dev = I.importance_map(roi, "gaussian", 0.125, "cuda", taps="device").cpu()
    # float32 erf differs by an ulp between CPU and GPU; erf(a)-erf(b) near 1 amplifies it to ~3e-3 relative per
    # axis in the tails (three factors multiply) - the reason the default evaluates the taps with torch on the host
    assert torch.allclose(dev, want, rtol=2e-2, atol=0)
    c = slice(24, 72)
    assert torch.allclose(dev[c, c, c], want[c, c, c], rtol=2e-5, atol=0)
    v12 = I.importance_map((24, 16, 32), "gaussian", 0.125, "cuda", variant="monai12").cpu()
    assert torch.equal(v12, M.compute_importance_map_v12((24, 16, 32)))
    v12d = I.importance_map((24, 16, 32), "gaussian", 0.125, "cuda", variant="monai12", taps="device").cpu()
    assert torch.allclose(v12d, M.compute_importance_map_v12((24, 16, 32)), rtol=1e-5)


@pytest.mark.parametrize("name", SMALL)
def test_compat_logits_bit_exact_vs_oracle(name):
    """sliding_window_inference (reference signature): same per-window logits, same weights -> identical floats."""
    case = SW_CASES[name]
    ref, ref_pred = oracle_run(case)
    vol, affine = cuda_inputs(case)
    pred = ArithmeticPredictor(case["k"])
    st = mss.InferStats()
    out = mss.sliding_window_inference(vol, affine, case["roi"], case["sw_batch"], pred, overlap=case["overlap"],
                                       mode=case["mode"], cval=case.get("cval", 0.0), mss_stats=st)
    assert out.shape == ref.shape and out.dtype == torch.float32
    assert pred.calls == ref_pred.calls  # same batches, same centre shapes (quirks Q3, Q10)
    got = out.cpu()
    assert rel_err(got, ref) <= TOL
    assert torch.equal(got, ref), f"max abs diff {(got - ref).abs().max().item()}"
    # and the fixture the reference produced in the build container
    fx = np.load(os.path.join(GOLD, f"sw_{name}.npz"))
    gold = json.load(open(os.path.join(GOLD, "manifest.json")))["sliding_window"][name]
    sample = fx["logits"].reshape(-1) if gold["stored"] == "full" else fx["sample"]
    mine = got.contiguous().numpy().reshape(-1)
    mine = mine if gold["stored"] == "full" else mine[::997]
    assert np.allclose(mine, sample, rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("name", SMALL)
@pytest.mark.parametrize("group_bytes", [None, 1])
def test_fused_labels_vs_oracle(name, group_bytes):
    """sliding_window_infer (labels straight from the accumulate kernel); group_bytes=1 forces one
    accumulate launch per predictor batch, i.e. the accumulator read-modify-write path."""
    case = SW_CASES[name]
    ref, _ = oracle_run(case)
    vol, affine = cuda_inputs(case)
    st = mss.InferStats()
    labels = mss.sliding_window_infer(vol, ArithmeticPredictor(case["k"]), case["roi"], case["overlap"], case["mode"],
                                      sw_batch_size=case["sw_batch"], cval=case.get("cval", 0.0), affine=affine,
                                      stats=st, group_bytes=group_bytes)
    assert tuple(labels.shape) == (case["shape"][0],) + tuple(case["shape"][2:])
    mism = assert_labels_match(labels, ref)
    assert mism <= 2  # identical logits -> identical labels unless softmax rounding merges a near-tie
    if group_bytes is None:
        assert st.n_accumulate_calls == 1 and not st.accumulator_allocated
    else:
        assert st.n_accumulate_calls == st.n_predictor_calls
        assert st.accumulator_allocated == (st.n_predictor_calls > 1)


@pytest.mark.parametrize("name", ["basic_b1", "two_volumes", "padded_cval", "constant_mode", "brats_like", "full_axis", "overlap_zero",
                                  "aniso_ragged", "cfg1_geometry"])
def test_rows_kernel_serves_single_launch_volumes(name):
    """K <= 16 (3, 4, 2, 5 and 14 here), every window in one launch, <= 4 window positions along W: the row-staged kernel
    (csrc/accumulate_rows.cu: bulk copies of whole window rows into a shared-memory ring) is the one that runs - same labels
    as the oracle, same near-tie census as the cell kernel's path (one launch per batch)."""
    from medicalsemseg_b200 import _lib
    case = SW_CASES[name]
    ref, _ = oracle_run(case)
    vol, affine = cuda_inputs(case)
    st, st2 = mss.InferStats(), mss.InferStats()
    kw = dict(sw_batch_size=case["sw_batch"], cval=case.get("cval", 0.0), affine=affine)
    labels = mss.sliding_window_infer(vol, ArithmeticPredictor(case["k"]), case["roi"], case["overlap"], case["mode"], stats=st, **kw)
    assert _lib.load().mss_accumulate_last_path() == 2, "expected the row-staged kernel"
    assert assert_labels_match(labels, ref) <= 2
    other = mss.sliding_window_infer(vol, ArithmeticPredictor(case["k"]), case["roi"], case["overlap"], case["mode"], stats=st2,
                                     group_bytes=1, **kw)
    assert st2.n_accumulate_calls == 1 or _lib.load().mss_accumulate_last_path() != 2  # several launches: the cell kernel
    assert torch.equal(labels, other) and st.near_ties == st2.near_ties
    # the same kernel finishes sum / count (engine/utils.py:151) when the caller wants the logits: bit-identical to the oracle
    both, logits = mss.sliding_window_infer(vol, ArithmeticPredictor(case["k"]), case["roi"], case["overlap"], case["mode"],
                                            return_logits=True, **kw)
    if _lib.load().mss_accumulate_last_path() == 2:
        assert torch.equal(logits.cpu(), ref)


@pytest.mark.parametrize("name", ["aniso_ragged", "two_volumes", "padded_cval", "brats_like"])
def test_labels_and_logits_together(name):
    case = SW_CASES[name]
    ref, _ = oracle_run(case)
    vol, affine = cuda_inputs(case)
    labels, logits = mss.sliding_window_infer(vol, ArithmeticPredictor(case["k"]), case["roi"], case["overlap"],
                                              case["mode"], sw_batch_size=case["sw_batch"], cval=case.get("cval", 0.0),
                                              affine=affine, return_logits=True, group_bytes=1)
    assert torch.equal(logits.cpu(), ref)
    assert assert_labels_match(labels, ref) == 0
    again = mss.logits_to_labels(logits.contiguous())
    assert torch.equal(again, labels.contiguous())


def test_plain_tensor_predictor_inferer_object():
    """run_evaluation.py:68-74 convention: SlidingWindowInferer hands the network the plain patch tensor (quirk Q6)."""
    case = SW_CASES["constant_mode"]
    ref, ref_pred = oracle_run(case, tuple_input=False)
    vol, _ = cuda_inputs(case)
    pred = ArithmeticPredictor(case["k"])
    inferer = mss.SlidingWindowInferer(roi_size=case["roi"], sw_batch_size=case["sw_batch"], overlap=case["overlap"],
                                       mode=case["mode"], cval=0.0)
    out = inferer(inputs=vol, network=pred)
    assert pred.calls == ref_pred.calls and all(c[1] is None for c in pred.calls)
    assert torch.equal(out.cpu(), ref)


def test_half_precision_logits_accumulate_in_fp32():
    """Under autocast the predictor returns fp16/bf16; engine/utils.py:147 promotes to fp32 before weighting."""
    case = SW_CASES["overlap_075"]
    for dt in (torch.float16, torch.bfloat16):
        class Half(ArithmeticPredictor):
            def __call__(self, x, *a, **k):
                return super().__call__(x, *a, **k).to(dt)
        vol = torch.from_numpy(make_volume(case))
        affine = torch.tensor([[1.5, 1.5, 2.0]], dtype=torch.float32)
        ref = osw.sliding_window_inference(vol, affine, case["roi"], case["sw_batch"], Half(case["k"]),
                                           overlap=case["overlap"], mode=case["mode"])
        out = mss.sliding_window_inference(vol.cuda(), affine.cuda(), case["roi"], case["sw_batch"], Half(case["k"]),
                                           overlap=case["overlap"], mode=case["mode"])
        assert torch.equal(out.cpu(), ref)


def test_non_constant_padding_modes():
    case = dict(SW_CASES["padded_cval"])
    vol, affine = cuda_inputs(case)
    for pm in ("reflect", "replicate", "circular"):
        ref = osw.sliding_window_inference(vol.cpu(), affine.cpu(), case["roi"], case["sw_batch"],
                                           ArithmeticPredictor(case["k"]), overlap=case["overlap"], mode=case["mode"],
                                           padding_mode=pm)
        out = mss.sliding_window_inference(vol, affine, case["roi"], case["sw_batch"], ArithmeticPredictor(case["k"]),
                                           overlap=case["overlap"], mode=case["mode"], padding_mode=pm)
        assert torch.equal(out.cpu(), ref), pm


def test_cfg1_geometry_128_cube_roi96():
    """BASELINE.json configs[0] stitching geometry (128^3, roi 96^3, overlap .25, K=14, N=8) incl. the TMA extract."""
    case = SW_CASES["cfg1_geometry"]
    ref, _ = oracle_run(case)
    vol, affine = cuda_inputs(case)
    st = mss.InferStats()
    out = mss.sliding_window_inference(vol, affine, case["roi"], case["sw_batch"], ArithmeticPredictor(case["k"]),
                                       overlap=case["overlap"], mode=case["mode"], mss_stats=st)
    assert st.n_windows == 8
    assert torch.equal(out.cpu(), ref)
    labels = mss.sliding_window_infer(vol, ArithmeticPredictor(case["k"]), 96, case["overlap"], "gaussian",
                                      sw_batch_size=case["sw_batch"], affine=affine)
    assert assert_labels_match(labels, ref) == 0
    gold = json.load(open(os.path.join(GOLD, "manifest.json")))["sliding_window"]["cfg1_geometry"]
    fx = np.load(os.path.join(GOLD, "sw_cfg1_geometry.npz"))
    assert np.allclose(out.cpu().contiguous().numpy().reshape(-1)[::997], fx["sample"], rtol=1e-5, atol=1e-6)
    assert gold["shape"] == list(out.shape)


def test_argument_errors_match_reference():
    vol = torch.zeros(1, 1, 20, 20, 20, device="cuda")
    with pytest.raises(AssertionError, match="overlap must be >= 0 and < 1."):
        mss.sliding_window_inference(vol, None, 16, 1, lambda x: x[0], overlap=1.0)
    with pytest.raises(ValueError):
        mss.sliding_window_inference(vol, None, (16, 16), 1, lambda x: x[0])
    with pytest.raises(ValueError):
        mss.sliding_window_inference(vol, None, 16, 1, lambda x: x[0], padding_mode="nearest")
    with pytest.raises(ValueError):  # predictor returning the wrong spatial shape
        mss.sliding_window_inference(vol, None, 16, 1, lambda x: x[0][..., :8])


@pytest.mark.parametrize("name,world,axis", [("aniso_ragged", 2, None), ("basic_b1", 3, 0), ("overlap_075", 3, 1),
                                             ("brats_like", 2, 2), ("cfg1_geometry", 2, None)])
def test_slab_partition_emulated_ranks(name, world, axis):
    """Every rank's work of the multi-GPU slab path (owned-window boxes, buffer origins, raw accumulation, halo add
    kernel, finalise with the GLOBAL weight count) executed rank after rank on one GPU; NCCL send/recv is replaced by
    a device copy.  Logits within 1e-5 of the oracle (summation order differs across the cut), labels exact outside
    near-ties."""
    from medicalsemseg_b200 import slab
    from medicalsemseg_b200.grid import make_grid
    case = SW_CASES[name]
    ref, _ = oracle_run(case, tuple_input=False)
    vol, _ = cuda_inputs(case)
    grid = make_grid(tuple(case["shape"][2:]), case["roi"], case["overlap"])
    part = slab.partition(grid, world, axis)
    ax = part.axis
    sts = [slab.local_pass(vol, ArithmeticPredictor(case["k"]), grid, part, r, case["mode"], sw_batch_size=case["sw_batch"],
                           group_bytes=None if r % 2 else 1) for r in range(world)]
    for r in range(1, world):
        lo, hi = part.halo(r - 1)
        src = slab._region(sts[r - 1].acc, ax, lo - part.buf_lo[r - 1], hi - part.buf_lo[r - 1]).contiguous()
        slab.cuda_halo_add(slab._region(sts[r].acc, ax, lo - part.buf_lo[r], hi - part.buf_lo[r]), src)
    labels = torch.empty((case["shape"][0],) + grid.image_size, dtype=torch.uint8, device="cuda")
    logits = torch.empty((case["shape"][0], case["k"]) + grid.image_size, device="cuda")
    for r in range(world):
        norm = torch.empty_like(sts[r].acc)
        own = slab.finalize_owned(sts[r], part, r, logits_out=norm)
        idx = [slice(None)] * 4
        idx[1 + ax] = slice(part.own_lo[r], part.own_hi[r])
        labels[tuple(idx)] = own
        lidx = [slice(None)] * 5
        lidx[2 + ax] = slice(part.own_lo[r], part.own_hi[r])
        logits[tuple(lidx)] = slab._region(norm, ax, part.own_lo[r] - part.buf_lo[r], part.own_hi[r] - part.buf_lo[r])[..., :grid.image_size[2] if ax != 2 else None]
    assert rel_err(logits.cpu(), ref) <= TOL
    assert_labels_match(labels, ref)


@pytest.mark.parametrize("name,world,dims", [("aniso_ragged", 4, (1, 2, 2)), ("basic_b1", 4, (2, 2, 1)),
                                             ("overlap_075", 8, (2, 2, 2)), ("brats_like", 8, (2, 2, 2)),
                                             ("overlap_075", 4, (1, 1, 4)), ("cfg1_geometry", 8, (2, 2, 2))])
def test_block_partition_emulated_ranks(name, world, dims):
    """The 3-D block partition (medicalsemseg_b200/block.py) rank after rank on one GPU: owned-window boxes cut along
    several axes, raw accumulation, axis-by-axis halo reduction (device copies stand in for NCCL send/recv, same box
    schedule as block.exchange_halos), finalise with the GLOBAL weight count."""
    from medicalsemseg_b200 import block, slab
    from medicalsemseg_b200.grid import make_grid
    case = SW_CASES[name]
    ref, _ = oracle_run(case, tuple_input=False)
    vol, _ = cuda_inputs(case)
    grid = make_grid(tuple(case["shape"][2:]), case["roi"], case["overlap"])
    part = block.block_partition(grid, world, dims)
    assert sum(part.n_windows(r) for r in range(world)) == grid.n_windows
    sts = [block.local_pass(vol, ArithmeticPredictor(case["k"]), grid, part, r, case["mode"], sw_batch_size=case["sw_batch"],
                            group_bytes=None if r % 2 else 1) for r in range(world)]
    live_lo = {r: [0, 0, 0] for r in range(world)}
    live_hi = {r: [h - l for l, h in zip(*part.box(r, "buf"))] for r in range(world)}
    for a in (2, 1, 0):
        p1 = part.axes[a]
        for i in range(1, p1.world):  # ascending along the axis: a forwarding chain sees its predecessor's additions
            for r in range(world):
                c = list(part.coords(r))
                if c[a] != i:
                    continue
                c[a] -= 1
                prev = part.rank_of(c)
                hlo, hhi = p1.halo(i - 1)
                if hhi <= hlo:
                    continue
                blo, bhi = list(live_lo[r]), list(live_hi[r])
                blo[a], bhi[a] = hlo - part.box(r, "buf")[0][a], hhi - part.box(r, "buf")[0][a]
                plo, phi = list(live_lo[prev]), list(live_hi[prev])
                plo[a], phi[a] = hlo - part.box(prev, "buf")[0][a], hhi - part.box(prev, "buf")[0][a]
                src = block._box_view(sts[prev].acc, plo, phi).contiguous()
                slab.cuda_halo_add(block._box_view(sts[r].acc, blo, bhi), src)
        for r in range(world):
            (bl, _), (ol, oh) = part.box(r, "buf"), part.box(r, "own")
            live_lo[r][a], live_hi[r][a] = ol[a] - bl[a], oh[a] - bl[a]
    labels = torch.empty((case["shape"][0],) + grid.image_size, dtype=torch.uint8, device="cuda")
    logits = torch.empty((case["shape"][0], case["k"]) + grid.image_size, device="cuda")
    for r in range(world):
        norm = torch.empty_like(sts[r].acc)
        own = block.finalize_owned(sts[r], part, r, logits_out=norm)
        ol, oh = part.box(r, "own")
        bl, _ = part.box(r, "buf")
        block._box_view(labels, ol, oh).copy_(own)
        block._box_view(logits, ol, oh).copy_(block._box_view(norm, [ol[x] - bl[x] for x in range(3)],
                                                              [oh[x] - bl[x] for x in range(3)]))
    assert rel_err(logits.cpu(), ref) <= TOL
    assert_labels_match(labels, ref)


from oracle.predictors import PositionalPredictor  # noqa: E402
from tests.golden.cases import NNUNET_CASES, make_nnunet_volume  # noqa: E402


@pytest.mark.parametrize("name", sorted(NNUNET_CASES))
@pytest.mark.parametrize("sw_batch", [1, 3])
def test_nnunet_tiled_predictor_bit_exact(name, sw_batch):
    """nnU-Net-style stitching policy + mirror TTA kernels vs the fixtures produced by the reference's own
    _internal_predict_3D_3Dconv_tiled / _internal_maybe_mirror_and_pred_3D."""
    from medicalsemseg_b200 import nnunet as N
    c = NNUNET_CASES[name]
    fx = np.load(os.path.join(os.path.dirname(__file__), "golden", f"nnunet_{name}.npz"))
    vol = torch.from_numpy(make_nnunet_volume(c)).cuda()
    st = mss.InferStats()
    with torch.no_grad():
        seg, probs = N.predict_3D_tiled(vol, PositionalPredictor(c["k"], c["patch"]), c["patch"], c["step"], c["mirror"],
                                        tuple(c["axes"]), c["gaussian"], nonlin=lambda t: t, sw_batch_size=sw_batch, stats=st,
                                        all_in_gpu=c.get("all_in_gpu", False))
    # all_in_gpu=True: the reference's half-precision branch (half importance map / aggregated results / counts / division)
    assert probs.dtype == (torch.float16 if c.get("all_in_gpu") else torch.float32) and seg.dtype == torch.uint8
    assert torch.equal(probs.float().cpu(), torch.from_numpy(fx["probs"]))
    assert np.array_equal(seg.cpu().numpy(), fx["seg"])
    assert st.gpu_launches > 0


def test_mirror_kernels_match_torch_flip():
    from medicalsemseg_b200 import _lib
    import ctypes as C
    lib = _lib.load()
    stream = torch.cuda.current_stream().cuda_stream
    for shape in [(2, 3, 8, 12, 16), (1, 2, 7, 9, 10), (1, 1, 16, 16, 15)]:   # W % 4 != 0 takes the scalar kernel
        x = torch.randn(shape, device="cuda")
        dims = _lib.I3(*shape[2:])
        preds, masks = [], []
        for m in range(8):
            out = torch.empty_like(x)
            _lib.check(lib.mss_flip_copy(x.data_ptr(), out.data_ptr(), shape[0] * shape[1], dims, m, stream), "flip")
            flips = [a for a, bit in ((4, 1), (3, 2), (2, 4)) if m & bit]
            assert torch.equal(out, torch.flip(x, flips) if flips else x)
            preds.append(torch.randn(shape, device="cuda"))
            masks.append(m)
        want = torch.zeros(shape, device="cuda")
        for m, p in zip(masks, preds):
            flips = [a for a, bit in ((4, 1), (3, 2), (2, 4)) if m & bit]
            want += 1 / 8 * (torch.flip(p, flips) if flips else p)
        got = torch.empty(shape, device="cuda")
        ptrs = (C.c_void_p * 8)(*[p.data_ptr() for p in preds])
        _lib.check(lib.mss_mirror_merge(ptrs, (C.c_int32 * 8)(*masks), 8, 0.125, got.data_ptr(), shape[0] * shape[1], dims,
                                        stream), "merge")
        assert torch.equal(got, want)
