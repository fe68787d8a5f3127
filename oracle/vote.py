"""CPU restatement of the reference's ensemble majority vote (TEST INFRASTRUCTURE).

Follows ``/root/reference/majority_vote.py:23-37``.  Checked bit-for-bit against
the reference's own ``get_class_votes``/``get_new_label`` (AST-extracted and
executed by ``tests/golden/make_golden.py``) in ``tests/test_oracle_golden.py``.
"""
from __future__ import annotations

from typing import Sequence

import numpy as np


def class_votes(label_maps: Sequence[np.ndarray], n_classes: int) -> np.ndarray:
    """majority_vote.py:23-33: ``votes[c] = #{f : map_f == c}`` for ``1 <= c < n_classes``;
    the background row is never counted and is set to the constant 1 (quirk Q7).
    Labels ``>= n_classes`` (or non-integers) match no row and are ignored."""
    shape = np.asarray(label_maps[0]).shape
    votes = np.zeros((n_classes,) + tuple(shape), dtype=np.uint64)
    for c in range(1, n_classes):
        for m in label_maps:
            votes[c] += (np.asarray(m) == c).astype(np.uint8)
    votes[0] += 1
    return votes


def majority_vote(label_maps: Sequence[np.ndarray], n_classes: int) -> np.ndarray:
    """majority_vote.py:35-37 + the uint8 cast at :83: first-max argmax over the votes."""
    return np.argmax(class_votes(label_maps, n_classes), axis=0).astype(np.uint8)


def majority_vote_rule(label_maps: Sequence[np.ndarray], n_classes: int) -> np.ndarray:
    """Closed form of the same rule: the lowest-index foreground class holding
    the maximal count wins if that count is >= 2, otherwise background."""
    maps = np.stack([np.asarray(m) for m in label_maps])
    best = np.zeros(maps.shape[1:], dtype=np.uint8)
    best_n = np.ones(maps.shape[1:], dtype=np.int64)  # background's fixed single vote
    for c in range(1, n_classes):
        n = (maps == c).sum(axis=0)
        win = n > best_n
        best[win] = c
        best_n[win] = n[win]
    return best
