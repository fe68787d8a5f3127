"""Flat multi-GPU partition on ONE GPU: the ranks run one after the other on the same device (their accumulators are
ordinary tensors, so mss_finalize_gather reads 'peers' through plain device pointers) and the assembled result is compared
with the CPU oracle and with the single-GPU path.  Real peer memory / NCCL: tests/multigpu_parity.py under torchrun."""
import numpy as np
import pytest
import torch

import medicalsemseg_b200 as mss
from medicalsemseg_b200 import flat
from medicalsemseg_b200.grid import make_grid
from medicalsemseg_b200.importance import importance_map
from oracle import sliding_window as osw
from oracle.predictors import ArithmeticPredictor
from tests.gpu_helpers import rel_err

pytestmark = pytest.mark.gpu

CASES = [
    dict(shape=(1, 1, 70, 40, 52), roi=(16, 16, 16), overlap=0.5, k=5, worlds=(2, 3, 8)),
    dict(shape=(1, 2, 64, 33, 47), roi=(24, 16, 16), overlap=0.25, k=3, worlds=(2, 5)),      # W % 4 != 0, unaligned starts
    dict(shape=(1, 1, 40, 36, 44), roi=(16, 16, 16), overlap=0.75, k=4, worlds=(4, 7)),      # 4+ windows per voxel and axis
    dict(shape=(1, 1, 128, 128, 128), roi=(96, 96, 96), overlap=0.25, k=14, worlds=(2, 8)),  # cfg1 geometry: 8 windows, 1 per rank
]


@pytest.mark.parametrize("ci", range(len(CASES)))
def test_flat_partition_emulated_ranks_match_oracle(ci):
    c = CASES[ci]
    rs = np.random.RandomState(70 + ci)
    vol = torch.from_numpy(rs.standard_normal(c["shape"]).astype(np.float32))
    pred = ArithmeticPredictor(c["k"])
    ref = osw.sliding_window_inference(vol, None, c["roi"], 3, pred, overlap=c["overlap"], mode="gaussian", tuple_input=False)
    want = osw.labels_from_logits(ref)
    near = osw.top2_relative_gap(ref) < 1e-5
    single = mss.sliding_window_infer(vol.cuda(), pred, c["roi"], c["overlap"], "gaussian", sw_batch_size=3)[0].cpu().numpy()
    grid = make_grid(c["shape"][2:], c["roi"], c["overlap"])
    imp = importance_map(grid.roi, "gaussian", 0.125, torch.device("cuda", 0))
    for world in c["worlds"]:
        part = flat.flat_partition(grid, world)
        sts = [flat.local_pass(vol, pred, grid, part, r, "gaussian", sw_batch_size=3, group_bytes=(1 if r % 2 else None))
               for r in range(world)]
        assert sum(st.total for st in sts) == grid.n_windows
        accs = [st.acc for st in sts]
        labels = torch.cat([flat.finalize_owned(grid, part, r, accs, imp)[0] for r in range(world)], dim=1)[0].cpu().numpy()
        logits = torch.cat([flat.finalize_owned(grid, part, r, accs, imp, return_logits=True)[1] for r in range(world)], dim=2)
        assert labels.shape == want.shape
        assert int(((labels != want) & ~near).sum()) == 0 and int(((labels != single) & ~near).sum()) == 0
        assert rel_err(logits.cpu(), ref) <= 1e-5  # partial sums of consecutive window ranges are added: association only


def test_flat_two_ranks_cfg4_geometry_labels_equal_single_gpu():
    """BraTS geometry (K = 3, odd W, clamped unaligned start 59) cut in two: same labels as one GPU outside near-ties."""
    gen = torch.Generator(device="cuda").manual_seed(11)
    vol = torch.randn((1, 4, 240, 240, 155), device="cuda", generator=gen)
    pred = ArithmeticPredictor(3)
    st = mss.InferStats()
    single, slog = mss.sliding_window_infer(vol, pred, 96, 0.5, "gaussian", sw_batch_size=4, return_logits=True, stats=st)
    grid = make_grid((240, 240, 155), 96, 0.5)
    imp = importance_map(grid.roi, "gaussian", 0.125, torch.device("cuda", 0))
    part = flat.flat_partition(grid, 2)
    accs = [flat.local_pass(vol, pred, grid, part, r, "gaussian", sw_batch_size=4).acc for r in range(2)]
    labels = torch.cat([flat.finalize_owned(grid, part, r, accs, imp)[0] for r in range(2)], dim=1)
    top = slog[0].topk(2, dim=0).values
    gap = (top[0] - top[1]) / torch.maximum(top[0].abs(), top[1].abs()).clamp_min(1e-37)
    bad = labels[0] != single[0]
    assert bool((gap[bad] < 1e-5).all())
