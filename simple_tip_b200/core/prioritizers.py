"""CTM / CAM orderings — reference: src/core/prioritizers.py:7-59.

`ctm` / `cam` keep the reference's signatures (dense boolean profiles, host NumPy).
`cam_from_buckets` is the same ordering for k-multisection profiles given in their compact form
(the bucket ids of `KMNC.buckets`): the greedy loop runs on the GPU (libb200tip `tip_cam_buckets`)
without ever building the N x D x k array, which is 41 GB at k = 1000."""
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


def cam_from_buckets(scores: np.ndarray, bucket, sections: int) -> Generator[int, None, None]:
    """`cam(scores, profiles)` for the one-hot-per-neuron profile whose compact form is
    bucket[n, d] in {-1, 0..sections-1} (NumPy array or torch CUDA tensor, int16 / int32, e.g. from
    `KMNC.buckets(..., device_out=True)`).  Yields exactly the indexes the reference's generator yields
    on the dense profile (prioritizers.py:16-59): greedy picks from the GPU, then the remaining samples
    by `np.argsort(-scores)` on the host (NumPy's own sort, so ties fall as in the reference)."""
    import torch

    from .. import _lib
    from .. import engine as E

    scores = np.asarray(scores).copy()
    assert scores.ndim == 1
    dev = E.require_cuda()
    lib = _lib.load()
    b = bucket if isinstance(bucket, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(bucket))
    if b.dtype not in (torch.int16, torch.int32):
        b = b.to(torch.int32)
    b = b.to(dev).contiguous()
    n, d = b.shape
    assert scores.shape[0] == n, "one score per sample"
    if n > 0 and d > 0:
        gain = ((b >= 0) & (b < int(sections))).sum(dim=1, dtype=torch.int32).contiguous()   # valid cells per sample
        covered = torch.zeros((d * int(sections) + 31) // 32, dtype=torch.int32, device=dev)
        newlist = torch.empty(2 * d, dtype=torch.int32, device=dev)
        order = torch.zeros(n, dtype=torch.int32, device=dev)
        state = torch.zeros(4, dtype=torch.int32, device=dev)
        bdt = _lib.TIP_I16 if b.dtype == torch.int16 else _lib.TIP_I32
        picks, done, rounds = 0, False, 64
        while not done and picks < n:
            _lib.check(lib.tip_cam_buckets(E._p(b), bdt, n, d, int(sections), E._p(gain), E._p(covered), E._p(newlist),
                                           E._p(order), E._p(state), min(rounds, n - picks), E._stream()),
                       "tip_cam_buckets")
            st = state.cpu().numpy()
            picks, done = int(st[0]), bool(st[1])
            rounds = min(4 * rounds, 1024)       # a sync every few hundred rounds is noise
        greedy = order[:picks].cpu().numpy().astype(np.int64)
    else:
        greedy = np.zeros(0, dtype=np.int64)
    yield from (int(i) for i in greedy)
    if n == 0:
        return
    taken = np.zeros(n, dtype=bool)
    taken[greedy] = True
    floor = np.min(scores) - 1
    scores[taken] = floor - 1
    for i in np.argsort(-scores):
        if scores[i] < floor:
            break
        yield int(i)
