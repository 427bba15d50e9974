"""Pins the oracle (oracle/np_oracle.py, oracle/tip_oracle.c) against golden vectors that
were produced by the UNMODIFIED reference classes (oracle/make_golden.py) and against the
known answers inside the reference's own tests (cited per test)."""
import ast

import numpy as np
import pytest

from oracle import c_oracle, np_oracle
from tests.conftest import case_names


def _kw(npz, name):
    return ast.literal_eval(str(npz[f"{name}.kw"]))


def test_pairwise_model_is_numpy_order():
    rng = np.random.default_rng(0)
    for dt in (np.float32, np.float64):
        for n in (1, 7, 8, 9, 100, 128, 129, 300, 1000, 2304):
            x = rng.normal(size=n).astype(dt)
            sq = x * x
            assert np_oracle.pairwise_sum_model(sq) == np.add.reduce(sq)
            assert c_oracle.pairwise_sumsq(x) == np.add.reduce(sq)
            assert np.sqrt(c_oracle.pairwise_sumsq(x)) == np.linalg.norm(x[None, None, :], axis=2)[0, 0]


def test_dsa_oracles_match_reference(golden):
    g = golden("dsa_reference.npz")
    for name in case_names(g, "dsa"):
        kw = _kw(g, name)
        xtr, ytr, xte, pte = (g[f"{name}.{k}"] for k in ("xtr", "ytr", "xte", "pte"))
        got = np_oracle.dsa_oracle(xtr, ytr, xte, pte, **kw)
        assert np.array_equal(got["dsa"], g[f"{name}.dsa"], equal_nan=True), name
        assert np.array_equal(got["dist_a"], g[f"{name}.dist_a"]), name
        assert np.array_equal(got["dist_b"], g[f"{name}.dist_b"]), name
        # C port: brute force in NumPy's summation order -> same bits, same winners
        tr = np_oracle.flatten_rows(xtr)
        sel = np_oracle.subsample_indexes(tr.shape[0], kw.get("subsampling", 1.0), 0)
        ytr_s = ytr
        if sel is not None:
            tr, ytr_s = tr[sel], ytr[sel]
        c = c_oracle.dsa(tr, ytr_s, np_oracle.flatten_rows(xte), pte)
        assert np.array_equal(c["dist_a"], g[f"{name}.dist_a"]), name
        assert np.array_equal(c["dist_b"], g[f"{name}.dist_b"]), name
        assert np.array_equal(c["idx_a"], got["idx_a"]), name
        assert np.array_equal(c["dsa"], g[f"{name}.dsa"], equal_nan=True), name


def test_dsa_golden_contains_exact_ties(golden):
    g = golden("dsa_reference.npz")
    assert (g["f32_ties.dist_a"] == 0).sum() >= 3        # test rows that are train rows
    assert np.unique(g["f32_ties.dist_a"]).size < 15         # integer grid: many exactly tied distances
    assert (g["f64_plausibility_id.dsa"] == 0).all()     # tests/test_surprise.py:151-155


def test_lsa_oracle_matches_reference(golden):
    g = golden("lsa_reference.npz")
    for name, kw in (("plaus", {}), ("cube", {}), ("mf30", {"max_features": 30}), ("far", {})):
        want = g[f"{name}.lsa"]
        got = np_oracle.lsa_oracle(g[f"{name}.xtr"], g[f"{name}.xte"], **kw)
        assert np.array_equal(np.isinf(got), np.isinf(want)), name
        f = np.isfinite(want)
        np.testing.assert_allclose(got[f], want[f], rtol=1e-10, err_msg=name)
        exact = np_oracle.lsa_oracle(g[f"{name}.xtr"], g[f"{name}.xte"], exact=True, **kw)
        np.testing.assert_allclose(exact[f], want[f], rtol=1e-12, err_msg=name)
    assert np.isinf(g["plaus.lsa"][10:]).all() and np.isfinite(g["plaus.lsa"][:10]).all()
    assert np.array_equal(np.array(np_oracle.lsa_removed_columns(g["mf30.xtr"], 30)), g["mf30.removed"])
    # stabilisation gives up -> every density 0 -> +inf (stable_kde.py:40-41,99-100)
    assert bool(g["singular.prepare_failed"])
    got = np_oracle.lsa_oracle(g["singular.xtr"], g["singular.xte"])
    assert np.isinf(got).all() and np.isinf(g["singular.lsa"]).all()
    # per-class routing (surprise.py:317-371)
    got = np_oracle.pc_lsa_oracle(g["mf30.xtr"], g["pclsa.ytr"], g["mf30.xte"], g["pclsa.pte"])
    np.testing.assert_allclose(got, g["pclsa.lsa"], rtol=1e-10)


def test_lsa_oracle_matches_modern_scipy(golden):
    """Independent cross-check of the restated scipy-1.4.1 evaluate: scipy's current
    gaussian_kde (float64 dataset, scott) implements the same estimator."""
    from scipy.stats import gaussian_kde

    g = golden("lsa_reference.npz")
    kde = gaussian_kde(g["cube.xtr"].astype(np.float64).T)
    want = -np.log(kde.evaluate(g["cube.xte"].astype(np.float64).T))
    np.testing.assert_allclose(np_oracle.lsa_oracle(g["cube.xtr"], g["cube.xte"]), want, rtol=1e-9)
    np.testing.assert_allclose(g["cube.lsa"], want, rtol=1e-9)


def test_kmnc_oracle_matches_reference(golden):
    g = golden("kmnc_reference.npz")
    for name in case_names(g, "score"):
        mins, maxs, act = g[f"{name}.mins"], g[f"{name}.maxs"], g[f"{name}.act"]
        cut, k = int(g[f"{name}.cut"]), int(g[f"{name}.sections"])
        layers_min, layers_max = [mins[:cut], mins[cut:]], [maxs[:cut], maxs[cut:]]
        bucket, hits = np_oracle.kmnc_buckets_oracle(layers_min, layers_max, k, [act[:, :cut], act[:, cut:]])
        assert np.array_equal(bucket, g[f"{name}.bucket"]), name
        assert np.array_equal(hits, g[f"{name}.hits"]), name
        assert hits.max() <= 1
        assert np.array_equal(hits.sum(axis=1), g[f"{name}.score"]), name
        if k <= 10:
            score, prof = np_oracle.kmnc_oracle(layers_min, layers_max, k, [act[:, :cut], act[:, cut:]])
            assert np.array_equal(score, g[f"{name}.score"]) and score.dtype == g[f"{name}.score"].dtype
        # C port with the very thresholds NumPy built
        _, _, thresh = np_oracle.kmnc_thresholds(layers_min, layers_max, k)
        tdt = np.result_type(act.dtype, thresh[0].dtype)
        cb, cs = c_oracle.kmnc(act.astype(tdt), np.stack(thresh).astype(tdt))
        assert np.array_equal(cb, g[f"{name}.bucket"]) and np.array_equal(cs, g[f"{name}.score"]), name


def test_kmnc_known_answer_from_reference_tests():
    """tests/test_coverage_metrics.py:38-64 (scores [13,13] then [11,13])."""
    acts = [np.array([[0.1, 0.4, 0.9, 0.4], [0.1, 0.9, 0.9, 0.4]]),
            np.array([[0.3, 0.2, 0.1, 0.6, 0.8], [0.3, 0.9, 0.1, 0.6, 0.8]]),
            np.array([[0.2, 0.3, 0.4, 0.4], [0.2, 0.9, 0.4, 0.4]])]
    mins = [np.array([0] * 4), np.array([0] * 5), np.array([0.1] * 4)]
    maxs = [np.array([1] * 4), np.array([1] * 5), np.array([0.95] * 4)]
    score, prof = np_oracle.kmnc_oracle(mins, maxs, 2, acts)
    assert np.all(score == [13, 13])
    assert np.all(prof[0][:4] == [[True, False], [True, False], [False, True], [True, False]])
    out = [a.copy() for a in acts]
    out[0][0][0], out[1][0][0] = -0.5, 1.5
    assert np.all(np_oracle.kmnc_oracle(mins, maxs, 2, out)[0] == [11, 13])


def test_deepgini_and_apfd_match_reference(golden):
    g = golden("gini_apfd_reference.npz")
    for name in ("c1_f32", "c1_f64", "wide"):
        p = g[f"{name}.p"]
        pred, gini = np_oracle.deepgini_oracle(p)
        assert np.array_equal(pred, g[f"{name}.pred"]) and np.array_equal(gini, g[f"{name}.gini"])
        cp, cg = c_oracle.deepgini(p)
        assert np.array_equal(cp, g[f"{name}.pred"]) and np.array_equal(cg, g[f"{name}.gini"]), name
        if f"{name}.apfd" in g.files:
            fault = pred != g[f"{name}.truth"]
            assert np_oracle.apfd_oracle(fault, np_oracle.ctm_order(gini)) == float(g[f"{name}.apfd"])
    for i in range(4):
        assert np_oracle.apfd_oracle(g[f"apfd{i}.fault"], g[f"apfd{i}.order"]) == float(g[f"apfd{i}.value"])


def test_deepgini_known_answer_from_reference_tests():
    """tests/test_deepgini.py:15-38 (exact equality in float64)."""
    batch = np.array([[0.1, 0.2, 0.3, 0.4], [0.5, 0.1, 0.1, 0.3], [0.25] * 4, [1.0, 0, 0, 0], [0, 1.0, 0, 0]])
    pred, unc = np_oracle.deepgini_oracle(batch)
    assert np.all(pred == [3, 0, 0, 0, 1]) and np.all(unc == np.array([0.7, 0.64, 0.75, 0, 0]))
    cp, cu = c_oracle.deepgini(batch)
    assert np.all(cp == pred) and np.all(cu == unc)


@pytest.mark.parametrize("order, fault, expected", [
    ([0, 1, 2], [True, True, True], (1 - 6 / 9 + 1 / 6)),
    ([0, 1, 2], [True, False, False], (1 - 1 / 3 + 1 / 6)),
    ([0, 1, 2], [False, False, True], (1 - 3 / 3 + 1 / 6)),
    ([2, 1, 0], [False, False, True], (1 - 1 / 3 + 1 / 6)),
    ([2, 1, 0], [True, False, False], (1 - 3 / 3 + 1 / 6)),
])
def test_apfd_known_answers_from_reference_tests(order, fault, expected):
    """tests/test_apfd.py:7-18."""
    assert np_oracle.apfd_oracle(np.array(fault), order) == expected


def test_compact_cam_oracle_matches_reference_cam_on_kmnc_profiles(golden):
    """prioritizers.py:16-59 run by the reference on dense KMNC profiles vs the compact restatement
    that only sees the bucket ids (what the GPU path, tip_cam_buckets, is checked against)."""
    g = golden("cam_kmnc_reference.npz")
    for i in range(4):
        bucket, score, k = g[f"camk{i}.bucket"], g[f"camk{i}.score"], int(g[f"camk{i}.sections"])
        want = g[f"camk{i}.order"]
        got = np_oracle.cam_from_buckets_oracle(score, bucket, k)
        assert np.array_equal(got, want), i
        assert sorted(got.tolist()) == list(range(bucket.shape[0]))      # a permutation of the samples
    # and the dense restatement agrees with both on a fresh case
    rng = np.random.default_rng(5)
    bucket = rng.integers(-1, 6, size=(50, 30)).astype(np.int32)
    prof = np.zeros((50, 30, 6), dtype=bool)
    np.put_along_axis(prof, np.maximum(bucket, 0)[..., None], (bucket >= 0)[..., None], axis=2)
    score = prof.sum(axis=(1, 2))
    assert np.array_equal(np_oracle.cam_oracle(score, prof), np_oracle.cam_from_buckets_oracle(score, bucket, 6))


def test_dense_cam_oracle_matches_reference(golden):
    """prioritizers.py:16-59 as run by the unmodified reference (tests/golden/prioritizers_reference.npz and the
    CAM orders over NAC/NBC/SNAC/TKNC and surprise-coverage profiles in siblings_reference.npz)."""
    g = golden("prioritizers_reference.npz")
    for i in range(4):
        assert np.array_equal(np_oracle.cam_oracle(g[f"cam{i}.scores"], g[f"cam{i}.profiles"]), g[f"cam{i}.order"]), i
    s = golden("siblings_reference.npz")
    for name in sorted({k.split(".")[1] for k in s.files if k.startswith("sib.") and k.endswith(".cam")}):
        assert np.array_equal(np_oracle.cam_oracle(s[f"sib.{name}.score"], s[f"sib.{name}.profile"]), s[f"sib.{name}.cam"]), name
    assert np.array_equal(np_oracle.cam_oracle(s["sc.values"], s["sc.profile"]), s["sc.cam"])


def test_sibling_coverage_oracles_match_reference(golden):
    """NAC / NBC / SNAC / TKNC and the streaming statistics (welford==0.2.5 restated) against the outputs of
    the reference's own classes (oracle/make_golden.py: sibling_cases)."""
    s = golden("siblings_reference.npz")
    train = [s[f"sib.train{i}"] for i in range(3)]
    test = [s[f"sib.test{i}"] for i in range(3)]
    cuts = s["sib.cuts"]
    mins, maxs, stds = np_oracle.stats_oracle([[l[a:b] for l in train] for a, b in zip(cuts[:-1], cuts[1:])])
    for i in range(3):
        for got, key in ((mins[i], "min"), (maxs[i], "max"), (stds[i], "std")):
            want = s[f"sib.{key}{i}"]
            assert got.dtype == want.dtype and np.array_equal(got, want), (key, i)
    checks = {"NAC_0": lambda: np_oracle.nac_oracle(0.0, test), "NAC_0.75": lambda: np_oracle.nac_oracle(0.75, test),
              "TKNC_1": lambda: np_oracle.tknc_oracle(1, test), "TKNC_3": lambda: np_oracle.tknc_oracle(3, test)}
    for sc in (0, 0.5, 1):
        checks[f"NBC_{sc}"] = lambda sc=sc: np_oracle.nbc_oracle(mins, maxs, stds, sc, test)
        checks[f"SNAC_{sc}"] = lambda sc=sc: np_oracle.snac_oracle(maxs, stds, sc, test)
    for name, fn in checks.items():
        score, prof = fn()
        assert score.dtype == s[f"sib.{name}.score"].dtype and np.array_equal(score, s[f"sib.{name}.score"]), name
        assert np.array_equal(prof, s[f"sib.{name}.profile"]), name
    # float64 statistics
    tr64 = [l.astype(np.float64) for l in train[:2]]
    mn, mx, sd = np_oracle.stats_oracle([tr64])
    for i in range(2):
        assert np.array_equal(mn[i], s[f"sib64.min{i}"]) and np.array_equal(mx[i], s[f"sib64.max{i}"])
        assert np.array_equal(sd[i], s[f"sib64.std{i}"])
    sc64, p64 = np_oracle.nbc_oracle(mn, mx, sd, 0.5, [l.astype(np.float64) for l in test[:2]])
    assert np.array_equal(sc64, s["sib64.NBC_0.5.score"]) and np.array_equal(p64, s["sib64.NBC_0.5.profile"])
