"""Import harness for the UNMODIFIED reference (TEST INFRASTRUCTURE ONLY).

Usable only where /root/reference exists (the authoring container): it is what
oracle/make_golden.py uses to produce tests/golden/*.npz, and what
tests/test_oracle_vs_reference.py uses (skipped when the directory is absent, e.g. on the
GPU box).  No reference source is copied; three harness-side shims make it importable on
the container's newer stack:

  1. `np.int = int`        — surprise.py:151,154,158 use the alias removed in NumPy 1.24.
  2. `uncertainty_wizard`  — not installable offline.  A stub module provides
                             `quantifiers.Quantifier`, `quantifiers.MaxSoftmax.calculate`
                             (argmax / max, uncertainty-wizard==0.2.0) and `ProblemType`,
                             so deepgini.py imports unchanged.
  3. `scipy.stats.gaussian_kde` — stable_kde.py:50 assigns `self.inv_cov`, a read-only
                             property since SciPy 1.10, and the class was written against
                             scipy==1.4.1 (requirements.txt:10).  While importing the
                             reference we substitute a restatement of the 1.4.1 class
                             (`Legacy141GaussianKDE`) so that the reference's own
                             `StableGaussianKDE`, `LSA` and `MultiModalSA` run unmodified.
                             Its evaluate() is the literal loop nest of
                             `_stats.gaussian_kernel_estimate` (C port: tip_oracle.c).
  4. `welford`             — welford==0.2.5 (requirements.txt:6) is not installable offline; a stub
                             module restates its `Welford` class (init / add / add_all / var_s) so that
                             the reference's own `AggregateStatisticsCollector`
                             (src/dnn_test_prio/aggregate_statistics.py) runs unmodified
                             (`load_aggregate_statistics`).
"""
from __future__ import annotations

import importlib
import os
import sys
import types

import numpy as np

REFERENCE_ROOT = os.environ.get("TIP_REFERENCE_ROOT", "/root/reference")


def available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "src", "core"))


class Legacy141GaussianKDE:
    """scipy==1.4.1 `scipy.stats.gaussian_kde`, the subset the reference touches
    (constructor, scott bandwidth, `_compute_covariance` hook, `evaluate`)."""

    def __init__(self, dataset, bw_method=None, weights=None):
        self.dataset = np.atleast_2d(np.asarray(dataset))
        if not self.dataset.size > 1:
            raise ValueError("`dataset` input should have multiple elements.")
        self.d, self.n = self.dataset.shape
        if weights is not None:
            self._weights = np.atleast_1d(weights).astype(float)
            self._weights /= sum(self._weights)
            self._neff = 1 / sum(self._weights ** 2)
        self.set_bandwidth(bw_method=bw_method)

    @property
    def weights(self):
        try:
            return self._weights
        except AttributeError:
            self._weights = np.ones(self.n) / self.n
            return self._weights

    @property
    def neff(self):
        try:
            return self._neff
        except AttributeError:
            self._neff = 1 / sum(self.weights ** 2)
            return self._neff

    def scotts_factor(self):
        return np.power(self.neff, -1.0 / (self.d + 4))

    covariance_factor = scotts_factor

    def set_bandwidth(self, bw_method=None):
        if bw_method is not None and bw_method != "scott":
            raise NotImplementedError("harness restates the default (scott) bandwidth only")
        self._compute_covariance()

    def _compute_covariance(self):  # overridden by the reference's StableGaussianKDE
        raise NotImplementedError

    def evaluate(self, points):
        from oracle import c_oracle

        points = np.atleast_2d(np.asarray(points))
        d, m = points.shape
        if d != self.d:
            if d == 1 and m == self.d:
                points = np.reshape(points, (self.d, 1))
                m = 1
            else:
                raise ValueError(f"points have dimension {d}, dataset has dimension {self.d}")
        dtype = np.common_type(self.covariance, points)
        whitening = np.linalg.cholesky(self.inv_cov).astype(dtype, copy=False)
        points_ = np.dot(self.dataset.T, whitening).astype(dtype, copy=False)
        xi_ = np.dot(points.T, whitening).astype(dtype, copy=False)
        norm = float(np.power(2 * np.pi, -self.d / 2.0))
        for i in range(self.d):
            norm *= whitening[i, i]
        # weights are all 1/n for the reference's usage (weights=None)
        return c_oracle.kde_eval(points_, xi_, float(self.weights[0]), float(norm))

    __call__ = evaluate


def _uwiz_stub() -> types.ModuleType:
    uw = types.ModuleType("uncertainty_wizard")
    q = types.ModuleType("uncertainty_wizard.quantifiers")

    class Quantifier:  # uncertainty_wizard.quantifiers.Quantifier (abstract base)
        pass

    class MaxSoftmax(Quantifier):
        @classmethod
        def calculate(cls, nn_outputs):
            idx = np.argmax(nn_outputs, axis=1)
            return idx, np.max(nn_outputs, axis=1)

    class ProblemType:
        CLASSIFICATION = 2
        REGRESSION = 1

    q.Quantifier, q.MaxSoftmax = Quantifier, MaxSoftmax
    uw.quantifiers, uw.ProblemType = q, ProblemType
    return uw


class Welford025:
    """welford==0.2.5 `welford.Welford`, the subset aggregate_statistics.py touches: constructed from a
    (1, ...) array (`init`), `add_all` = sequential `add` per sample, `var_s` = s / (count - 1).  Mean and
    the sum of squared deviations keep the dtype NumPy gives them (float32 for float32 activations)."""

    def __init__(self, elements=None):
        self._shape = None
        self._count, self._m, self._s = 0, None, None
        if elements is not None:
            self.init(elements)

    def init(self, elements):
        self._shape = elements[0].shape
        self._count = elements.shape[0]
        self._m = np.mean(elements, axis=0)
        self._s = np.var(elements, axis=0, ddof=0) * elements.shape[0]

    def add(self, element, backup_flg=True):
        if self._shape is None:
            self._shape = element.shape
            self._m, self._s = np.zeros(element.shape), np.zeros(element.shape)
        self._count += 1
        delta = element - self._m
        self._m += delta / self._count
        self._s += delta * (element - self._m)

    def add_all(self, elements, backup_flg=True):
        for elem in elements:
            self.add(elem, backup_flg=False)

    @property
    def count(self):
        return self._count

    @property
    def mean(self):
        return self._m

    @property
    def var_s(self):
        if self._count <= 0:
            return None
        if self._count <= 1:
            return np.full(self._shape, np.nan)
        return self._s / (self._count - 1)

    @property
    def var_p(self):
        if self._count <= 0:
            return None
        return self._s / self._count


def _welford_stub() -> types.ModuleType:
    w = types.ModuleType("welford")
    w.Welford = Welford025
    return w


def load_aggregate_statistics():
    """The reference's `src.dnn_test_prio.aggregate_statistics` module (unmodified) on top of the welford stub."""
    if not available():
        raise RuntimeError(f"reference not found under {REFERENCE_ROOT}")
    sys.modules.setdefault("welford", _welford_stub())  # shim 4
    saved_path = list(sys.path)
    saved_src = {k: v for k, v in sys.modules.items() if k == "src" or k.startswith("src.")}
    for k in saved_src:
        del sys.modules[k]
    try:
        here = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
        sys.path[:] = [REFERENCE_ROOT] + [p for p in saved_path if os.path.abspath(p or ".") != here]
        import importlib.util as _ilu

        importlib.invalidate_caches()

        # the package __init__ of src.dnn_test_prio is empty, but import the one file only: its siblings need TensorFlow
        timer = importlib.import_module("src.core.timer")
        assert os.path.abspath(timer.__file__).startswith(os.path.abspath(REFERENCE_ROOT))
        spec = _ilu.spec_from_file_location(
            "tip_ref_aggregate_statistics", os.path.join(REFERENCE_ROOT, "src", "dnn_test_prio", "aggregate_statistics.py"))
        mod = _ilu.module_from_spec(spec)
        spec.loader.exec_module(mod)
        return mod
    finally:
        sys.path[:] = saved_path
        for k in [k for k in sys.modules if k == "src" or k.startswith("src.")]:
            del sys.modules[k]
        sys.modules.update(saved_src)


_loaded = {}


def load():
    """Returns a namespace with the reference's L1 modules (surprise, neuron_coverage,
    deepgini, apfd, prioritizers)."""
    if _loaded:
        return types.SimpleNamespace(**_loaded)
    if not available():
        raise RuntimeError(f"reference not found under {REFERENCE_ROOT}")
    if not hasattr(np, "int"):
        np.int = int  # shim 1
    sys.modules.setdefault("uncertainty_wizard", _uwiz_stub())  # shim 2
    sys.modules.setdefault("uncertainty_wizard.quantifiers", sys.modules["uncertainty_wizard"].quantifiers)

    import scipy.stats as st

    saved_kde = st.gaussian_kde
    # The reference must win `import src.core...` even when this repo's overlay `src/` is on
    # sys.path: import it under a private package name pointing at the reference tree.
    saved_path = list(sys.path)
    saved_src = {k: v for k, v in sys.modules.items() if k == "src" or k.startswith("src.")}
    for k in saved_src:
        del sys.modules[k]
    try:
        st.gaussian_kde = Legacy141GaussianKDE  # shim 3
        sys.path[:] = [REFERENCE_ROOT] + [p for p in saved_path
                                         if os.path.abspath(p or ".") != os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))]
        importlib.invalidate_caches()
        for name in ("surprise", "neuron_coverage", "deepgini", "apfd", "prioritizers", "stable_kde"):
            mod = importlib.import_module(f"src.core.{name}")
            assert os.path.abspath(mod.__file__).startswith(os.path.abspath(REFERENCE_ROOT)), mod.__file__
            _loaded[name] = mod
    finally:
        st.gaussian_kde = saved_kde
        sys.path[:] = saved_path
        for k in [k for k in sys.modules if k == "src" or k.startswith("src.")]:
            del sys.modules[k]
        sys.modules.update(saved_src)
    return types.SimpleNamespace(**_loaded)
