"""CTM / CAM orderings — reference: src/core/prioritizers.py:7-59.

`ctm` / `cam` keep the reference's signatures; `cam`'s greedy loop runs on the GPU over the bit-packed
profile (`cam_from_bits`, one persistent cooperative kernel), which is also the entry point for profiles that
never left HBM (`NAC/NBC/SNAC/TKNC.packed`).  `cam_from_buckets` is the same ordering for k-multisection profiles given in their compact form
(the bucket ids of `KMNC.buckets`): the greedy loop runs on the GPU (libb200tip `tip_cam_buckets`)
without ever building the N x D x k array, which is 41 GB at k = 1000."""
from typing import Generator

import numpy as np


def ctm(scores: np.ndarray) -> Generator[int, None, None]:
    """Coverage-total method: indexes by decreasing score (`np.argsort(-scores)`)."""
    assert len(scores.shape) == 1
    yield from np.argsort(-scores)


LAST_GREEDY_PICKS = 0      # greedy rounds of the most recent CAM call (bench / diagnostics)
LAST_GREEDY_MS = 0.0       # device time of its greedy loop (CUDA events around the launches)


def _tail_by_score(scores: np.ndarray, greedy: np.ndarray):
    """prioritizers.py:47-59: the samples the greedy loop did not yield, by `np.argsort(-scores)` — NumPy's
    own (unstable) sort on the host, so ties fall exactly as in the reference."""
    n = scores.shape[0]
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


def cam_from_bits(scores: np.ndarray, bits, rounds_per_launch: int = 4096) -> Generator[int, None, None]:
    """`cam(scores, profiles)` for a boolean profile given bit-packed in HBM: bits [N, words] int32 CUDA
    tensor (any consistent packing — `NAC/NBC/SNAC/TKNC.packed()` or `pack_profiles`).  The greedy loop
    (prioritizers.py:24-45) is ONE persistent cooperative kernel (libb200tip `tip_cam_bits`): per round an
    arg-max of the gains, new = profile[pick] & ~covered, gain -= popcount(profile & new)."""
    import torch

    from .. import _lib
    from .. import engine as E

    scores = np.asarray(scores).copy()
    assert scores.ndim == 1
    lib = _lib.load()
    n, words = int(bits.shape[0]), int(bits.shape[1]) if bits.ndim == 2 else 0
    assert scores.shape[0] == n, "one score per sample"
    greedy = np.zeros(0, dtype=np.int64)
    if n > 0 and words > 0:
        dev = bits.device
        bits = bits.contiguous()
        gain = torch.empty(n, dtype=torch.int32, device=dev)
        covered = torch.zeros(2 * words, dtype=torch.int32, device=dev)
        cand = torch.zeros(4 * _lib.CAM_MAX_BLOCKS, dtype=torch.int32, device=dev)
        order = torch.zeros(n, dtype=torch.int32, device=dev)
        state = torch.zeros(4, dtype=torch.int32, device=dev)
        picks, done, first = 0, False, 1
        rounds = max(2, int(rounds_per_launch) & ~1)              # even: the covered set is double-buffered by parity
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
        while not done and picks < n:
            _lib.check(lib.tip_cam_bits(E._p(bits), n, words, E._p(gain), E._p(covered), E._p(cand), E._p(order),
                                        E._p(state), rounds, first, E._stream()), "tip_cam_bits")
            first = 0
            ev1.record()
            st = state.cpu().numpy()
            picks, done = int(st[0]), bool(st[1])
        greedy = order[:picks].cpu().numpy().astype(np.int64)
        global LAST_GREEDY_MS
        LAST_GREEDY_MS = float(ev0.elapsed_time(ev1))
    global LAST_GREEDY_PICKS
    LAST_GREEDY_PICKS = int(greedy.shape[0])
    yield from (int(i) for i in greedy)
    yield from _tail_by_score(scores, greedy)


def pack_profiles(profiles):
    """Dense boolean profiles [N, ...] (NumPy or torch, host or device) -> bit-packed [N, ceil(F/32)] int32
    CUDA tensor (libb200tip `tip_pack_bool`)."""
    import torch

    from .. import _lib
    from .. import engine as E

    dev = E.require_cuda()
    if isinstance(profiles, torch.Tensor):
        p = profiles.reshape(profiles.shape[0], -1).to(dev)
        p = (p != 0).to(torch.uint8).contiguous() if p.dtype != torch.bool else p.contiguous()
    else:
        flat = np.ascontiguousarray(np.asarray(profiles).reshape((profiles.shape[0], -1)))
        p = torch.from_numpy(flat.view(np.uint8) if flat.dtype == np.bool_ else (flat != 0).view(np.uint8)).to(dev)
    n, f = int(p.shape[0]), int(p.shape[1])
    words = (f + 31) // 32
    bits = torch.empty((n, max(words, 0)), dtype=torch.int32, device=dev)
    if n and f:
        _lib.check(_lib.load().tip_pack_bool(E._p(p), n, f, E._p(bits), E._stream()), "tip_pack_bool")
    return bits


def cam(scores: np.ndarray, profiles: np.ndarray) -> Generator[int, None, None]:
    """Coverage-additional method (prioritizers.py:16-59): greedily pick the sample covering the most
    still-uncovered profile entries (first index on ties, like np.argmax); once nothing new can be
    covered, the remaining samples follow by decreasing score.  Same signature as the reference (dense
    boolean profiles of any shape [N, ...]); the profile is bit-packed on the GPU and the greedy loop runs
    there (`cam_from_bits`)."""
    scores = np.asarray(scores)
    assert len(scores.shape) == 1
    yield from cam_from_bits(scores, pack_profiles(profiles))


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
    global LAST_GREEDY_PICKS
    LAST_GREEDY_PICKS = int(greedy.shape[0])
    yield from (int(i) for i in greedy)
    yield from _tail_by_score(scores, greedy)
