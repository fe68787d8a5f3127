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
