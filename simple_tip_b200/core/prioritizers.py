"""CTM / CAM orderings — reference: src/core/prioritizers.py:7-59 (host)."""
from typing import Generator

import numpy as np


def ctm(scores: np.ndarray) -> Generator[int, None, None]:
    """Coverage-total method: indexes by decreasing score (`np.argsort(-scores)`)."""
    assert len(scores.shape) == 1
    yield from np.argsort(-scores)


def cam(scores: np.ndarray, profiles: np.ndarray) -> Generator[int, None, None]:
    """Coverage-additional method: greedily pick the sample covering the most still-uncovered
    profile entries (first index on ties, like np.argmax); once nothing new can be covered,
    the remaining samples follow by decreasing score."""
    scores = scores.copy()
    prof = profiles.reshape((profiles.shape[0], -1)).copy()
    gain = prof.sum(axis=1).flatten()
    todo = prof.shape[1]
    taken = np.zeros(scores.shape[0], dtype=bool)
    while todo > 0:
        pick = int(np.argmax(gain))
        fresh = gain[pick]
        if fresh == 0:
            break
        yield pick
        taken[pick] = True
        cols = prof[pick].nonzero()[0]
        todo -= fresh
        gain = gain - prof[:, cols].sum(axis=1)
        prof[:, cols] = 0
    floor = np.min(scores) - 1
    scores[taken] = floor - 1
    for i in np.argsort(-scores):
        if scores[i] < floor:
            break
        yield i
