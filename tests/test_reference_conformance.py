"""The reference's own unit tests (reference tests/test_surprise.py, test_prioritizers.py,
test_apfd.py, test_timer.py — cited per test), restated against the drop-in overlay `src.core`.
Host-only behaviour runs on CPU; anything that scores on the device carries the gpu marker."""
import time

import numpy as np
import pytest

from src.core.apfd import apfd_from_order
from src.core.prioritizers import cam, ctm
from src.core.surprise import (DSA, LSA, MDSA, MLSA, MultiModalSA, SurpriseCoverageMapper, _by_class_discriminator,
                               _class_predictions, _flatten_predictions, _KmeansDiscriminator)
from src.core.timer import Timer


# ---- reference tests/test_surprise.py:18-56 (host) ------------------------------------------
@pytest.mark.parametrize("acts, preds", [
    ([[0.1, 0.2, 0.3], [0.4, 0.5, 0.6]], [0, 1]),
    ([[0.1, 0.2, 0.3], [0.4, 0.5, 0.6], [0.4, 0.5, 0.6]], [0, 1, 1])])
def test_by_class_discriminator(acts, preds):
    ids = _by_class_discriminator(np.array(acts), np.array(preds))
    assert ids.shape == (len(preds),) and np.all(ids == np.array(preds))


@pytest.mark.parametrize("preds, num_classes, message", [
    ([0.5, 0.5], 2, "Predictions must be integers"),
    ([-1, 5, 7], 2, "Class predictions must be >= 0"),
    ([0, 2, 6], 6, "must be < num_classes"),
    ([[0, 0, 0, 1]], 2, "must be one-dimensional")])
def test_class_prediction_assertions(preds, num_classes, message):
    with pytest.raises(AssertionError) as e:
        _class_predictions(preds, num_classes=num_classes)
    assert message in str(e.value)


@pytest.mark.parametrize("value", [np.array([0, 2, 3, 5, 0.1, -5]), [0, 2, 3, 5, 0.1, -5]])
def test_flatten_predictions(value):
    assert np.all(np.array([0, 2, 3, 5, 0.1, -5]) == _flatten_predictions(value))
    assert _flatten_predictions(None) is None


# ---- reference tests/test_surprise.py:59-94 (host, known answers) ----------------------------
@pytest.mark.parametrize("buckets, limit, overflow, sa, expected", [
    (3, 1, False, [0.1, 0.2, 0.8], [[True, False, False], [True, False, False], [False, False, True]]),
    (3, 1, True, [0.1, 0.2, 0.8], [[True, False, False], [True, False, False], [False, True, False]]),
    (3, 1, True, [0.1, 0.2, 1.1], [[True, False, False], [True, False, False], [False, False, True]])])
def test_surprise_coverage_mapper(buckets, limit, overflow, sa, expected):
    profile = SurpriseCoverageMapper(buckets, limit, overflow).get_coverage_profile(np.array(sa))
    assert profile.shape == (3, 3) and profile.dtype == bool and np.all(profile == np.array(expected))


def test_surprise_coverage_mapper_golden(golden):
    g = golden("prioritizers_reference.npz")
    assert np.array_equal(SurpriseCoverageMapper(10, 2.5).get_coverage_profile(g["scm.values"]), g["scm.profile"])
    assert np.array_equal(SurpriseCoverageMapper(10, 2.5, overflow_bucket=True).get_coverage_profile(g["scm.values"]),
                          g["scm.profile_overflow"])
    for i in range(4):
        assert np.array_equal(np.array(list(ctm(g[f"cam{i}.scores"]))), g[f"cam{i}.ctm"]), i


@pytest.mark.gpu
def test_prioritizers_golden(golden):
    g = golden("prioritizers_reference.npz")
    for i in range(4):
        scores, prof = g[f"cam{i}.scores"], g[f"cam{i}.profiles"]
        assert np.array_equal(np.array(list(cam(scores, prof.copy()))), g[f"cam{i}.order"]), i
        assert np.array_equal(np.array(list(ctm(scores))), g[f"cam{i}.ctm"]), i


# ---- reference tests/test_prioritizers.py:27-64 (DeepGini-paper example) ---------------------
@pytest.mark.gpu
def test_cam_ctm_paper_example():
    profiles = np.array([[1, 1, 0, 0, 0, 0, 0, 0], [1, 1, 1, 1, 0, 0, 0, 0], [0, 0, 0, 0, 1, 1, 1, 0],
                         [0, 0, 0, 0, 0, 0, 1, 1]], dtype=bool)
    scores = profiles.sum(axis=1)
    order = list(cam(scores, profiles))
    assert order[0] == 1 and order[1] == 2 and set(order) == {0, 1, 2, 3} and len(order) == 4
    assert order[2] == 3                     # only sample adding coverage (column 7)
    assert list(ctm(scores))[0] == 1
    rng = np.random.default_rng(0)
    big = rng.random((200, 300)) < 0.05
    seen = list(cam(big.sum(axis=1), big))
    assert sorted(seen) == list(range(200))


# ---- reference tests/test_apfd.py:7-18 --------------------------------------------------------
@pytest.mark.parametrize("order, fault, expected", [
    ([0, 1, 2], [True, True, True], (1 - 6 / 9 + 1 / 6)),
    ([0, 1, 2], [True, False, False], (1 - 1 / 3 + 1 / 6)),
    ([0, 1, 2], [False, False, True], (1 - 3 / 3 + 1 / 6)),
    ([2, 1, 0], [False, False, True], (1 - 1 / 3 + 1 / 6)),
    ([2, 1, 0], [True, False, False], (1 - 3 / 3 + 1 / 6))])
def test_apfd_sanity(order, fault, expected):
    assert apfd_from_order(np.array(fault), order) == expected


def test_apfd_golden(golden):
    g = golden("gini_apfd_reference.npz")
    for i in range(4):
        assert apfd_from_order(g[f"apfd{i}.fault"], g[f"apfd{i}.order"]) == float(g[f"apfd{i}.value"])
    with pytest.raises(AssertionError):
        apfd_from_order(np.zeros((2, 2)), [0, 1])


# ---- reference tests/test_timer.py -------------------------------------------------------------
def test_timer():
    t = Timer()
    t.start()
    time.sleep(0.05)
    t.stop()
    assert 0.2 > t.get() >= 0.05
    with t:
        with pytest.warns(RuntimeWarning):
            t.get()
        with pytest.raises(RuntimeError):
            t.start()
    with pytest.raises(RuntimeError):
        t.stop()
    assert Timer(start=True)._running_since is not None


# ---- reference tests/test_surprise.py:97-119 (device) ------------------------------------------
@pytest.mark.gpu
def test_multi_modal_sa_routing():
    rng = np.random.RandomState(42)
    acts = rng.random((10000, 10))
    labels = rng.randint(0, 3, size=10000)
    sa = MultiModalSA.build_by_class(acts, labels, lambda x, y: LSA(x))
    assert sa.modal_sa.keys() == {0, 1, 2} and sa.modal_sa[0].__class__ == LSA
    t_acts, t_labels = rng.random((1000, 10)), rng.randint(0, 3, size=1000)
    surprises = sa(t_acts, t_labels)
    assert surprises.shape == (1000,) and np.sum(surprises == -np.inf) == 0
    for label in range(3):
        direct = sa.modal_sa[label](t_acts[t_labels == label], t_labels[t_labels == label])
        assert np.all(surprises[t_labels == label] == direct)
    with pytest.raises(ValueError, match="No modal found"):
        sa(t_acts[:3], np.array([0, 1, 7]))
    assert sa(t_acts[:0], t_labels[:0]).shape == (0,)


# ---- reference tests/test_surprise.py:133-171 (metamorphic + determinism, device) ---------------
@pytest.mark.gpu
@pytest.mark.parametrize("creator, strictly_positive", [
    pytest.param(lambda x, y: MDSA(x), True, id="MDSA"),
    pytest.param(lambda x, y: LSA(x), False, id="LSA"),
    pytest.param(lambda x, y: DSA(x, y), False, id="DSA")])
def test_sa_plausibility(creator, strictly_positive):
    rng = np.random.RandomState(42)
    acts = rng.random((100, 10))
    labels = rng.randint(0, 3, size=100)
    sa = creator(acts, labels)
    id_sa = sa(acts[:10], labels[:10])
    ood_sa = sa(acts[:10] + 10, labels[:10])
    assert np.all(ood_sa > id_sa)
    if strictly_positive:
        assert np.all(id_sa >= 0) and np.all(ood_sa >= 0)
    assert id_sa.shape == ood_sa.shape == (10,)
    big = np.concatenate([acts for _ in range(100)])
    big_labels = np.concatenate([labels for _ in range(100)])
    first = sa(big, big_labels).reshape((100, -1))
    assert np.all(first == first[0])
    assert np.all(sa(big, big_labels).reshape((100, -1)) == first)


# ---- reference tests/test_surprise.py:122-130, 174-229 (sklearn fits on the host, scoring on the GPU) -------------
@pytest.mark.gpu
def test_mdsa_mlsa_kmeans_paths():
    rng = np.random.RandomState(42)
    acts = rng.random((20000, 10))
    np.testing.assert_allclose(MDSA(acts).covariance_matrix.covariance_, np.cov(acts.T), 0.1)
    three = np.concatenate([rng.random((2000, 10)), rng.random((2000, 10)) + 0.4, rng.random((2000, 10)) + 0.9])
    mlsa = MLSA(three, num_components=3)
    centres = np.array([[0.5] * 10, [0.9] * 10, [1.4] * 10])
    assert len(set(mlsa.gmm.predict(centres))) == 3
    assert np.all(mlsa(centres + 2) > mlsa(centres))
    two = np.concatenate([rng.random((100, 10)), rng.random((100, 10)) + 0.9])
    disc = _KmeansDiscriminator(two, [2, 3, 4])
    assert disc.best_k == 2
    pts = np.array([[0.5] * 10, [1.4] * 10])
    assert len(set(disc(pts, None))) == 2
    mm = MultiModalSA.build_with_kmeans(two, None, lambda x, _: MDSA(x), potential_k=[2, 3, 4])
    assert np.all(mm(pts + 2, None) > mm(pts, None))


# ---- reference tests/test_prioritizers.py:11-135 (paper example under every shape / shuffle, fuzzer) ------------
def _paper_example(seed):
    import random

    rows = [[True, True, True, False, False, True, True, True], [True, True, True, False, False, False, True, True],
            [True, True, True, True, False, False, False, False], [False, False, False, False, True, True, True, True]]
    names = ["A", "B", "C", "D"]
    random.Random(seed).shuffle(rows)
    random.Random(seed).shuffle(names)
    return np.array(rows, dtype=bool), names


@pytest.mark.gpu
def test_cam_paper_example_every_shape_and_shuffle():
    for seed in range(10):
        profile, names = _paper_example(seed)
        scores = np.sum(profile, axis=1)
        assert [names[i] for i in ctm(scores)] in (["A", "B", "C", "D"], ["A", "B", "D", "C"])
        for shape in [(4, 8), (4, 8, 1), (4, 4, 2), (4, 2, 2, 2), (-1, 2, 4)]:
            got = [names[i] for i in cam(scores, np.reshape(profile, shape))]
            assert got in (["A", "D", "C", "B"], ["A", "C", "D", "B"]), (seed, shape, got)


@pytest.mark.gpu
@pytest.mark.parametrize("shape, prob", [((20, 100), 0.1), ((200, 1000), 0.0001), ((2000, 10000), 0.01), ((4000, 40000), 0.01)])
def test_cam_fuzzer_invariants(shape, prob):
    """The consistency checks of the reference's test_cam_fuzzer: a permutation; coverage increments weakly
    decreasing while coverage grows; afterwards scores weakly decreasing."""
    rng = np.random.default_rng(1)
    profile = rng.random(shape) < prob
    scores = np.sum(profile, axis=1)
    order = [i for i in cam(scores, profile.copy())]
    assert sorted(order) == list(range(shape[0]))
    covered = np.zeros(shape[1], dtype=bool)
    last_inc, prev_sum, last_score, tail = np.inf, 0, np.inf, False
    for i in order:
        covered |= profile[i]
        new_sum = int(covered.sum())
        inc = new_sum - prev_sum
        assert inc <= last_inc
        if inc == 0:
            tail = True
        if tail:
            assert inc == 0 and scores[i] <= last_score
            last_score = scores[i]
        last_inc, prev_sum = inc, new_sum
