"""CPU oracle for the test-input-prioritisation hot path (TEST INFRASTRUCTURE ONLY).

This module is a NumPy restatement of the reference's L1 arithmetic
(`/root/reference/src/core`).  It exists to *check* the CUDA path; nothing in the
product package (`simple_tip_b200/`) may import it.  Allowed importers: `tests/`,
`__graft_entry__.smoke()`, and `bench.py`'s `cpu_baseline` / `--impl reference` leg.

Pinning (see tests/test_oracle_golden.py and oracle/make_golden.py):
  * DSA, KMNC, DeepGini, APFD: pinned against outputs of the reference classes themselves,
    executed in the authoring container and committed under tests/golden/.
  * LSA: the reference's own `LSA`/`StableGaussianKDE` classes cannot run on SciPy >= 1.10
    (`stable_kde.py:50` assigns a read-only property) and the pair arithmetic lives in
    the un-vendored scipy==1.4.1 (`requirements.txt:10`).  The golden vectors for LSA were
    produced by the reference classes running on top of a restated scipy-1.4.1
    `gaussian_kde` base (oracle/ref_harness.py) and cross-checked against
    scipy 1.18 `gaussian_kde.evaluate` (1e-13 rel).  The scipy part is therefore
    "parity pinned to a restatement", which DESIGN.md states explicitly.

Every function cites the reference lines it follows.
"""
from __future__ import annotations

import math
from concurrent.futures import ThreadPoolExecutor
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np


# --------------------------------------------------------------------------------------
# NumPy's pairwise summation order (what `np.add.reduce` does along a contiguous axis).
# The CUDA re-rank reproduces this order so that distances are bit-identical.
# --------------------------------------------------------------------------------------
def pairwise_sum_model(a: np.ndarray):
    """Scalar model of NumPy's `*_pairwise_sum` (umath loops): returns the same bits as
    `np.add.reduce(a)` for a contiguous 1-D float32/float64 array.  Slow; tests only."""
    t = a.dtype.type
    n = a.shape[0]
    if n < 8:
        res = t(0.0)
        for i in range(n):
            res = t(res + a[i])
        return res
    if n <= 128:
        r = [a[j] for j in range(8)]
        i = 8
        while i < n - (n % 8):
            for j in range(8):
                r[j] = t(r[j] + a[i + j])
            i += 8
        res = t(t(t(r[0] + r[1]) + t(r[2] + r[3])) + t(t(r[4] + r[5]) + t(r[6] + r[7])))
        while i < n:
            res = t(res + a[i])
            i += 1
        return res
    n2 = n // 2
    n2 -= n2 % 8
    return t(pairwise_sum_model(a[:n2]) + pairwise_sum_model(a[n2:]))


# --------------------------------------------------------------------------------------
# helpers (surprise.py:62-87, 136-183)
# --------------------------------------------------------------------------------------
def flatten_rows(layers) -> np.ndarray:
    """surprise.py:168-177 / neuron_coverage.py:25-28: samples x everything-else."""
    if isinstance(layers, np.ndarray):
        return layers if layers.ndim == 2 else layers.reshape(layers.shape[0], -1)
    return np.concatenate([np.reshape(l, (l.shape[0], -1)) for l in layers], axis=1)


def subsample_indexes(n: int, subsampling, seed: int) -> Optional[np.ndarray]:
    """surprise.py:72-86: `RandomState(seed).choice(arange(n), k, replace=False)`."""
    if subsampling == 1.0:
        return None
    if isinstance(subsampling, int) and subsampling > 0:
        k = min(subsampling, n)
    elif 0 < subsampling < 1:
        k = int(subsampling * n)
    else:
        raise ValueError("bad subsampling")
    return np.random.RandomState(seed).choice(np.arange(n), k, replace=False)


# --------------------------------------------------------------------------------------
# DSA (surprise.py:523-651)
# --------------------------------------------------------------------------------------
def _closest(from_rows: np.ndarray, to_rows: np.ndarray):
    """surprise.py:633-651 with identical NumPy expressions (so identical bits)."""
    diff = from_rows[:, None] - to_rows
    norms = np.linalg.norm(diff, axis=2)
    del diff
    return np.min(norms, axis=1), np.argmin(norms, axis=1)


def dsa_oracle(
    train: np.ndarray,
    train_pred: np.ndarray,
    test: np.ndarray,
    test_pred: np.ndarray,
    badge_size: int = 10,
    subsampling=1.0,
    subsampling_seed: int = 0,
    threads: int = 1,
) -> Dict[str, np.ndarray]:
    """Distance-based surprise adequacy exactly as surprise.py:530-631 computes it.

    Returns dsa (float64, as `np.empty(...)` + scatter, surprise.py:576,611), and for
    index-parity checks also dist_a / dist_b (input dtype) and idx_a = position of the
    stage-1 winner in the (sub-sampled) training array.
    """
    train = flatten_rows(train)
    train_pred = np.asarray(train_pred).astype(np.int64)
    sel = subsample_indexes(train.shape[0], subsampling, subsampling_seed)
    if sel is not None:
        train, train_pred = train[sel], train_pred[sel]
    test = flatten_rows(test)
    test_pred = np.asarray(test_pred).astype(np.int64)
    num_classes = int(train_pred.max()) + 1
    class_rows = [np.argwhere(train_pred == c).flatten() for c in range(num_classes)]

    m = test.shape[0]
    dsa = np.full(m, np.nan, dtype=np.float64)
    dist_a = np.zeros(m, dtype=train.dtype)
    dist_b = np.zeros(m, dtype=train.dtype)
    idx_a = np.full(m, -1, dtype=np.int64)

    tasks = []
    for c in range(num_classes):
        rows = np.argwhere(test_pred == c).flatten()
        if rows.size == 0:
            continue
        for badge in np.array_split(rows, math.ceil(rows.shape[0] / badge_size)):
            tasks.append((badge, c))

    def run(task):
        rows, c = task
        same = train[class_rows[c]]                      # surprise.py:619
        da, pos = _closest(test[rows], same)             # :620
        winners = same[pos]                              # :648
        mask = np.ones(train.shape[0], dtype=bool)       # :622-626
        mask[class_rows[c]] = False
        db, _ = _closest(winners, train[mask])           # :627-629
        return rows, da, db, class_rows[c][pos]

    if threads > 1:
        with ThreadPoolExecutor(max_workers=threads) as ex:   # surprise.py:599 uses 5
            results = list(ex.map(run, tasks))
    else:
        results = [run(t) for t in tasks]
    for rows, da, db, ia in results:
        with np.errstate(divide="ignore", invalid="ignore"):
            dsa[rows] = da / db                          # :595, widened on store (:611)
        dist_a[rows], dist_b[rows], idx_a[rows] = da, db, ia
    return {"dsa": dsa, "dist_a": dist_a, "dist_b": dist_b, "idx_a": idx_a}


# --------------------------------------------------------------------------------------
# LSA (surprise.py:399-495, stable_kde.py:20-101, scipy==1.4.1 gaussian_kde semantics)
# --------------------------------------------------------------------------------------
class KdeFit:
    """State after `StableGaussianKDE.__init__` (stable_kde.py:20-53)."""

    def __init__(self, dataset_dxn: np.ndarray):
        self.dataset = np.atleast_2d(np.asarray(dataset_dxn)).astype(np.float64)   # :22
        self.d, self.n = self.dataset.shape
        if not self.dataset.size > 1:
            raise ValueError("`dataset` input should have multiple elements.")
        self.factor = float(np.power(self.n, -1.0 / (self.d + 4)))      # scott, scipy kde
        cov = np.atleast_2d(np.cov(self.dataset, rowvar=1, bias=False,
                                   aweights=np.ones(self.n) / self.n))        # :31-33 (weights = 1/n)
        self.prepare_failed = False
        # stable_kde.py:55-77 — note: the diagonal is REPLACED, not incremented.
        increment = 1e-10
        while np.any(np.linalg.eigh(cov * self.factor ** 2)[0] <= 0):
            np.fill_diagonal(cov, increment)
            if increment > 1e-5:
                self.prepare_failed = True
                break
            increment += increment
        if self.prepare_failed:
            self.inv_cov = None
            return
        try:
            inv = np.linalg.inv(cov)                                     # :43
        except np.linalg.LinAlgError:
            self.prepare_failed = True
            self.inv_cov = None
            return
        self.covariance = cov * self.factor ** 2                          # :49
        self.inv_cov = inv / self.factor ** 2                             # :50
        np.linalg.cholesky(self.covariance * 2 * np.pi)                   # :51 (raises if not PD)

    def evaluate(self, points_dxm: np.ndarray, block: int = 256) -> np.ndarray:
        """scipy 1.4.1 `gaussian_kde.evaluate` -> `_stats.gaussian_kernel_estimate`:
        whiten with cholesky(inv_cov), sum_i w_i * exp(-|p_i-q_j|^2/2) * norm,
        norm = (2 pi)^(-d/2) * prod(diag(W)), w_i = 1/n.  float64 throughout.
        Blocked over test points; the inner sum uses NumPy's summation order rather than
        the Cython loop's strictly sequential one (difference ~1e-16 relative)."""
        points = np.atleast_2d(np.asarray(points_dxm)).astype(np.float64)
        m = points.shape[1]
        if self.prepare_failed:
            return np.zeros(m)                                            # stable_kde.py:99-100
        w = np.linalg.cholesky(self.inv_cov)
        p = self.dataset.T @ w
        q = points.T @ w
        norm = math.pow(2 * math.pi, -self.d / 2.0)
        for i in range(self.d):
            norm *= w[i, i]
        weight = 1.0 / self.n
        out = np.zeros(m)
        pn = np.einsum("ij,ij->i", p, p)
        for s in range(0, m, block):
            qb = q[s:s + block]
            # explicit differences would be O(block*n*d) memory; expand but keep float64
            d2 = pn[None, :] + np.einsum("ij,ij->i", qb, qb)[:, None] - 2.0 * (qb @ p.T)
            np.maximum(d2, 0.0, out=d2)
            out[s:s + block] = np.sum((np.exp(-d2 / 2.0) * norm) * weight, axis=1)
        return out

    def evaluate_exact(self, points_dxm: np.ndarray) -> np.ndarray:
        """Same as evaluate() but with explicit residuals (no norm expansion), i.e. the
        literal loop body of gaussian_kernel_estimate; O(n*d) per test point."""
        points = np.atleast_2d(np.asarray(points_dxm)).astype(np.float64)
        m = points.shape[1]
        if self.prepare_failed:
            return np.zeros(m)
        w = np.linalg.cholesky(self.inv_cov)
        p = self.dataset.T @ w
        q = points.T @ w
        norm = math.pow(2 * math.pi, -self.d / 2.0)
        for i in range(self.d):
            norm *= w[i, i]
        out = np.zeros(m)
        for j in range(m):
            r = p - q[j]
            arg = np.einsum("ij,ij->i", r, r)
            out[j] = np.sum((np.exp(-arg / 2.0) * norm) * (1.0 / self.n))
        return out


def lsa_removed_columns(train: np.ndarray, max_features=300) -> List[int]:
    """surprise.py:425-434."""
    if max_features < 1:
        num = min(max_features * train.shape[1], train.shape[1])
    else:
        num = min(max_features, train.shape[1])
    return [int(x) for x in np.argsort(np.var(train, axis=0))[:-num]]


def lsa_oracle(train: np.ndarray, test: np.ndarray, max_features=300, exact: bool = False) -> np.ndarray:
    """LSA(train)(test) as in surprise.py:399-495 (max_features path)."""
    train = flatten_rows(train)
    test = flatten_rows(test)
    removed = lsa_removed_columns(train, max_features)
    if removed:
        train = np.delete(train, removed, axis=1)
        test = np.delete(test, removed, axis=1)
    kde = KdeFit(train.transpose())
    dens = kde.evaluate_exact(test.transpose()) if exact else kde.evaluate(test.transpose())
    with np.errstate(divide="ignore"):
        return -np.log(dens)                                              # surprise.py:495


def pc_lsa_oracle(train, train_pred, test, test_pred, max_features=300) -> np.ndarray:
    """MultiModalSA.build_by_class(x, y, lambda x, y: LSA(x)) (surprise.py:234-371,
    handler_surprise.py:26)."""
    train, test = flatten_rows(train), flatten_rows(test)
    train_pred, test_pred = np.asarray(train_pred), np.asarray(test_pred)
    out = np.full(test_pred.shape, -np.inf, dtype=np.float64)
    for c in np.unique(test_pred):
        if not np.any(train_pred == c):
            raise ValueError(f"No modal found for modal id {c}. Check your discriminator")
        out[test_pred == c] = lsa_oracle(train[train_pred == c], test[test_pred == c], max_features)
    return out


# --------------------------------------------------------------------------------------
# DeepGini (deepgini.py:31-35; uncertainty-wizard==0.2.0 MaxSoftmax = argmax / max)
# --------------------------------------------------------------------------------------
def deepgini_oracle(p: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
    pred = np.argmax(p, axis=1)
    gini = 1 - np.sum(p * p, axis=1)
    return pred, gini


# --------------------------------------------------------------------------------------
# KMNC (neuron_coverage.py:65-94, sum_score :8-22)
# --------------------------------------------------------------------------------------
def kmnc_thresholds(mins: Sequence[np.ndarray], maxs: Sequence[np.ndarray], sections: int):
    lo = np.concatenate([np.asarray(l).flatten() for l in mins])
    hi = np.concatenate([np.asarray(l).flatten() for l in maxs])
    jumps = (hi - lo) / sections                                          # :76
    return lo, jumps, [lo + jumps * i for i in range(sections + 1)]       # :79


def kmnc_oracle(mins, maxs, sections: int, activations) -> Tuple[np.ndarray, np.ndarray]:
    """Dense (scores, profiles) exactly as KMNC.__call__ builds them."""
    _, _, thresh = kmnc_thresholds(mins, maxs, sections)
    act = flatten_rows(list(activations)) if not isinstance(activations, np.ndarray) else flatten_rows(activations)
    prof = np.zeros((act.shape[0], act.shape[1], sections), dtype=bool)
    for i in range(sections):
        prof[..., i] = np.logical_and(thresh[i] <= act, act < thresh[i + 1])  # :91-93
    total = int(np.prod(prof[0].shape))
    dt = np.int16 if total <= np.iinfo(np.int16).max else (np.int32 if total <= np.iinfo(np.int32).max else np.int64)
    return np.sum(prof.reshape(prof.shape[0], -1), axis=1, dtype=dt), prof


def kmnc_buckets_oracle(mins, maxs, sections: int, activations) -> Tuple[np.ndarray, np.ndarray]:
    """Compact form of the same predicate: bucket[n,d] = the i with
    thresh[i] <= a < thresh[i+1] (or -1), hits[n,d] = number of i satisfying it
    (0 or 1 for monotone thresholds), without materialising N x D x k."""
    _, _, thresh = kmnc_thresholds(mins, maxs, sections)
    act = flatten_rows(list(activations)) if not isinstance(activations, np.ndarray) else flatten_rows(activations)
    bucket = np.full(act.shape, -1, dtype=np.int32)
    hits = np.zeros(act.shape, dtype=np.int32)
    for i in range(sections):
        mask = np.logical_and(thresh[i] <= act, act < thresh[i + 1])
        bucket[mask] = i
        hits += mask
    return bucket, hits


# --------------------------------------------------------------------------------------
# NAC / NBC / SNAC / TKNC (neuron_coverage.py:52-62, 97-173) and their fit
# (src/dnn_test_prio/aggregate_statistics.py:12-67 on welford==0.2.5)
# --------------------------------------------------------------------------------------
def _sum_score(prof: np.ndarray) -> np.ndarray:
    total = int(np.prod(prof[0].shape))
    dt = np.int16 if total <= np.iinfo(np.int16).max else (np.int32 if total <= np.iinfo(np.int32).max else np.int64)
    return np.sum(prof.reshape(prof.shape[0], -1), axis=1, dtype=dt)                  # :8-22


def _cat(layers):
    return np.concatenate([np.asarray(l).flatten() for l in layers])


def nac_oracle(threshold, activations):
    prof = flatten_rows(list(activations)) > threshold                                # :60-61
    return _sum_score(prof), prof


def nbc_oracle(mins, maxs, stds, scaler, activations):
    lo = _cat(mins) - scaler * _cat(stds)                                             # :113-114
    hi = _cat(maxs) + scaler * _cat(stds)
    act = flatten_rows(list(activations))
    prof = np.zeros((act.shape[0], act.shape[1], 2), dtype=bool)
    prof[..., 0] = act <= lo                                                          # :129-130
    prof[..., 1] = act >= hi
    return _sum_score(prof), prof


def snac_oracle(maxs, stds, scaler, activations):
    hi = _cat(maxs) + scaler * _cat(stds)                                             # :142
    prof = flatten_rows(list(activations)) >= hi                                      # :146-147
    return _sum_score(prof), prof


def tknc_oracle(top_neurons: int, activations):
    """neuron_coverage.py:159-173.  np.argsort is unstable: for exactly tied activations at the k-th rank the
    marked neuron is implementation-defined (use tie-free inputs when comparing)."""
    per_layer = []
    for layer in activations:
        flat = layer.reshape((layer.shape[0], -1))
        top = np.argsort(flat, axis=1)[..., -top_neurons:]
        mark = np.zeros_like(flat, dtype=bool)
        np.put_along_axis(mark, top, True, axis=1)
        per_layer.append(mark)
    prof = np.concatenate(per_layer, axis=1)
    return _sum_score(prof), prof


def stats_oracle(badges):
    """AggregateStatisticsCollector.track per badge, then get() (aggregate_statistics.py:37-67): running
    np.minimum / np.maximum and, per layer, welford==0.2.5's `Welford`: initialised with the first sample
    (mean = x, s = 0), then `add` per sample — count += 1; delta = x - m; m += delta / count;
    s += delta * (x - m) — in the dtype NumPy gives the state (float32 for float32 activations);
    std = sqrt(s / (count - 1)).  badges: list of badges, each a list of layer arrays [n_b, ...]."""
    mins = maxs = ms = ss = None
    count = 0
    for badge in badges:
        badge = [np.asarray(l) for l in badge]
        if mins is None:
            mins = [l[0] for l in badge]
            maxs = [l[0] for l in badge]
            ms = [np.mean(np.expand_dims(l[0], 0), axis=0) for l in badge]
            ss = [np.var(np.expand_dims(l[0], 0), axis=0, ddof=0) * 1 for l in badge]
            count = 1
            badge = [l[1:] for l in badge]
        mins = [np.minimum(mins[i], np.min(badge[i], axis=0)) for i in range(len(badge))] if badge[0].shape[0] else mins
        maxs = [np.maximum(maxs[i], np.max(badge[i], axis=0)) for i in range(len(badge))] if badge[0].shape[0] else maxs
        for r in range(badge[0].shape[0]):
            count += 1
            for i, l in enumerate(badge):
                delta = l[r] - ms[i]
                ms[i] += delta / count
                ss[i] += delta * (l[r] - ms[i])
    stds = [np.sqrt(s / (count - 1)) if count > 1 else np.sqrt(np.full(s.shape, np.nan)) for s in ss]
    return mins, maxs, stds


def cam_oracle(scores: np.ndarray, profiles: np.ndarray) -> np.ndarray:
    """Coverage-additional order on a dense boolean profile, restated from prioritizers.py:16-59."""
    scores = np.asarray(scores).copy()
    prof = np.asarray(profiles).reshape((profiles.shape[0], -1)).copy()
    gain = np.sum(prof, axis=1).flatten()                                             # :22
    remaining = prof.shape[1]
    yielded = np.zeros(scores.shape[0], dtype=bool)
    order = []
    while remaining > 0 and prof.shape[0] > 0:
        nxt = int(np.argmax(gain))                                                    # :26 (first index on ties)
        fresh = gain[nxt]
        if fresh == 0:                                                                # :30-31
            break
        order.append(nxt)
        yielded[nxt] = True
        cols = prof[nxt].nonzero()[0]
        remaining -= fresh
        gain = gain - np.sum(prof[:, cols], axis=1)                                   # :38-39
        prof[:, cols] = 0
    if scores.shape[0]:
        floor = np.min(scores) - 1                                                    # :48-52
        scores[yielded] = floor - 1
        for x in np.argsort(-scores):
            if scores[x] < floor:
                break
            order.append(int(x))
    return np.array(order, dtype=np.int64)


# --------------------------------------------------------------------------------------
# APFD / CTM (apfd.py:8-19, prioritizers.py:7-13, eval_apfd_table.py:86,101)
# --------------------------------------------------------------------------------------
def apfd_oracle(is_fault: np.ndarray, order) -> float:
    assert is_fault.ndim == 1
    ordered = is_fault[order]
    pos = np.where(ordered == 1)[0]
    k = np.count_nonzero(is_fault)
    n = is_fault.shape[0]
    return 1 - (np.sum(pos + 1) / (k * n)) + (1 / (2 * n))


def ctm_order(scores: np.ndarray) -> np.ndarray:
    return np.argsort(-scores)


def cam_from_buckets_oracle(scores: np.ndarray, bucket: np.ndarray, sections: int) -> np.ndarray:
    """Coverage-additional order (prioritizers.py:16-59) for a one-hot-per-neuron profile given by
    its compact form bucket[n, d] in {-1, 0..sections-1} (KMNC, neuron_coverage.py:82-94), without
    building the N x D x k array: `num_coverable` (:22) starts as the number of valid cells of a
    sample; a pick (`np.argmax`, first index on ties, :26) newly covers the cells (d, bucket[pick, d])
    not covered yet; every sample loses one per newly covered cell it shares (:38-39); the loop ends
    when the best gain is 0 (:30-31) or everything coverable is covered (:44-45); the rest follows
    by `np.argsort(-scores)` over the not-yet-yielded samples (:47-59)."""
    scores = np.asarray(scores).copy()
    bucket = np.asarray(bucket)
    n, d = bucket.shape
    covered = np.zeros((d, sections), dtype=bool)
    gain = (bucket >= 0).sum(axis=1).astype(np.int64)
    remaining = d * sections
    taken = np.zeros(n, dtype=bool)
    order = []
    cols = np.arange(d)
    while True:
        pick = int(np.argmax(gain))
        fresh = int(gain[pick])
        if fresh == 0:
            break
        order.append(pick)
        taken[pick] = True
        cell = bucket[pick]
        new = (cell >= 0) & ~covered[cols, np.maximum(cell, 0)]
        remaining -= fresh
        nd = cols[new]
        gain = gain - (bucket[:, nd] == cell[nd][None, :]).sum(axis=1)
        covered[nd, cell[nd]] = True
        if remaining == 0:
            break
    floor = np.min(scores) - 1
    scores[taken] = floor - 1
    for i in np.argsort(-scores):
        if scores[i] < floor:
            break
        order.append(int(i))
    return np.asarray(order, dtype=np.int64)


# --------------------------------------------------------------------------------------
# Seeded synthetic traces for the BASELINE.json configurations
# --------------------------------------------------------------------------------------
def synth_clusters(n_train: int, n_test: int, d: int, classes: int, seed: int,
                   dtype=np.float32, flip: float = 0.1, spread: float = 2.0):
    """Gaussian class clusters: centre ~ N(0, spread^2 I), rows = centre[label] + N(0, I);
    predicted test label = true label with `flip` share replaced by a random class."""
    rng = np.random.default_rng(seed)
    centres = rng.normal(0.0, spread, size=(classes, d))
    ytr = rng.integers(0, classes, size=n_train)
    yte = rng.integers(0, classes, size=n_test)
    xtr = (centres[ytr] + rng.normal(size=(n_train, d))).astype(dtype)
    xte = (centres[yte] + rng.normal(size=(n_test, d))).astype(dtype)
    pte = yte.copy()
    flips = rng.random(n_test) < flip
    pte[flips] = rng.integers(0, classes, size=int(flips.sum()))
    return xtr, ytr.astype(np.int64), xte, pte.astype(np.int64), yte.astype(np.int64)


def synth_softmax(n: int, c: int, seed: int, dtype=np.float32):
    """C1: softmax(N(0, 3^2) logits); 'true' label = argmax(logits + N(0,1))."""
    rng = np.random.default_rng(seed)
    logits = rng.normal(0.0, 3.0, size=(n, c))
    e = np.exp(logits - logits.max(axis=1, keepdims=True))
    p = (e / e.sum(axis=1, keepdims=True)).astype(dtype)
    truth = np.argmax(logits + rng.normal(size=(n, c)), axis=1)
    return p, truth


def synth_relu(n: int, d: int, seed: int, n_stat: int = 50000, dtype=np.float32):
    """C4: ReLU(N(0,1)) activations; min/max from a separate draw."""
    rng = np.random.default_rng(seed)
    stat = np.maximum(rng.normal(size=(min(n_stat, 4096), d)), 0).astype(dtype)
    mins, maxs = stat.min(axis=0), stat.max(axis=0)
    act = np.maximum(rng.normal(size=(n, d)), 0).astype(dtype)
    return act, mins, maxs
