"""Neuron-coverage criteria behind the reference's `src.core.neuron_coverage` API.

KMNC (neuron_coverage.py:65-94 of the reference) is scored by libb200tip.so: one streaming
pass computes, per (sample, neuron), the section index i with
`thresh[i] <= a < thresh[i+1]` (thresholds `min + jumps*i` re-evaluated on the fly with NumPy's
rounding) and the per-sample count, i.e. the compact form of the reference's dense
N x D x sections boolean profile.  `KMNC.__call__` expands it to the dense profile only because
the reference API returns one (consumers: `cam`, handler_coverage.py:122-124); `buckets()` is the
compact entry point for large `sections`.  NAC / NBC / SNAC / TKNC keep the reference's contract
with host NumPy (next rows).
"""
from __future__ import annotations

import abc
from typing import List, Tuple

import numpy as np


def sum_score(profiles: np.ndarray) -> np.ndarray:
    """Count of covered profile entries per sample, in the narrowest of int16/32/64 that can
    hold the maximum (neuron_coverage.py:8-22)."""
    assert profiles.dtype == np.bool_
    return np.sum(profiles.reshape((profiles.shape[0], -1)), axis=1, dtype=_score_dtype(np.prod(profiles[0].shape)))


def _score_dtype(max_value: int):
    if max_value <= np.iinfo(np.int16).max:
        return np.int16
    if max_value <= np.iinfo(np.int32).max:
        return np.int32
    return np.int64


def flatten_layers(layers: List[np.ndarray]) -> np.ndarray:
    return np.concatenate([np.reshape(l, (l.shape[0], -1)) for l in layers], axis=1)


class CoverageMethod(abc.ABC):
    def __init__(self):
        pass

    @abc.abstractmethod
    def __call__(self, activations: List[np.ndarray]) -> Tuple[np.ndarray, np.ndarray]:
        ...


class NAC(CoverageMethod):
    """Neuron activation coverage: a > threshold."""

    def __init__(self, cov_threshold: float):
        super().__init__()
        self.cov_threshold = cov_threshold

    def __call__(self, activations):
        prof = flatten_layers(activations) > self.cov_threshold
        return sum_score(prof), prof


class KMNC(CoverageMethod):
    """k-multisection neuron coverage on the GPU."""

    def __init__(self, mins: List[np.ndarray], maxs: List[np.ndarray], sections: int):
        super().__init__()
        self.sections = sections
        lo = np.concatenate([np.asarray(l).flatten() for l in mins])
        hi = np.concatenate([np.asarray(l).flatten() for l in maxs])
        jumps = (hi - lo) / sections                       # neuron_coverage.py:76 (same NumPy expression)
        self._lo, self._jumps = lo + jumps * 0, jumps       # lo promoted to the threshold dtype
        self._dev_stats = None

    @property
    def thresh(self) -> List[np.ndarray]:
        """The reference's threshold list (neuron_coverage.py:79), built on demand."""
        lo = self._lo
        return [lo + self._jumps * i for i in range(self.sections + 1)]

    def buckets(self, activations, device_out: bool = False):
        """(scores int32 [N], bucket ids [N, D] int16/int32; -1 = no section covered).
        Activations may be NumPy arrays or device-resident torch tensors (one or a list);
        device_out=True returns the two results as CUDA tensors (no D2H of the N x D bucket ids)."""
        import torch

        from .. import _lib
        from .. import engine as E

        dev_act = E.device_matrix(activations)
        if dev_act is not None:
            act = dev_act if dev_act.dtype in (torch.float32, torch.float64) else dev_act.to(torch.float32)
            act = act.contiguous()
        elif isinstance(activations, np.ndarray):
            act = activations.reshape((activations.shape[0], -1))
        elif len(activations) == 1:          # one layer: no concatenation copy (keeps a pinned buffer pinned)
            act = np.reshape(activations[0], (activations[0].shape[0], -1))
        else:
            act = flatten_layers(activations)
        stat_dt = self._jumps.dtype
        if stat_dt not in (np.float32, np.float64):
            raise TypeError(f"KMNC statistics must be float32/float64 after NumPy promotion, got {stat_dt}")
        if dev_act is None and act.dtype not in (np.float32, np.float64):
            act = act.astype(np.result_type(act.dtype, stat_dt))
            if act.dtype not in (np.float32, np.float64):
                raise TypeError(f"unsupported activation dtype {act.dtype}")
        act_np_dtype = E.NP_DTYPE[act.dtype] if dev_act is not None else act.dtype
        dev = E.require_cuda()
        lib = _lib.load()
        n, d = act.shape
        assert d == self._lo.shape[0], "activation width does not match the min/max statistics"
        if self._dev_stats is None:
            self._dev_stats = (E.to_device(self._lo.astype(stat_dt), dev), E.to_device(self._jumps, dev))
        lo_dev, jump_dev = self._dev_stats
        a_dev = E.to_device(act, dev)
        small = self.sections <= np.iinfo(np.int16).max
        bucket = torch.empty((n, d), dtype=torch.int16 if small else torch.int32, device=dev)
        score = torch.empty(n, dtype=torch.int32, device=dev)
        _lib.check(lib.tip_kmnc(E._p(a_dev), E.tip_dtype(act_np_dtype), n, d, E._p(lo_dev), E._p(jump_dev),
                                E.tip_dtype(stat_dt), self.sections, E._p(bucket),
                                _lib.TIP_I16 if small else _lib.TIP_I32, E._p(score), E._stream()), "tip_kmnc")
        if device_out:
            return score, bucket
        # D2H into cached pinned buffers (the bucket ids are the bulk of the traffic of this call)
        key = (n, d, bucket.dtype)
        if getattr(self, "_host_key", None) != key:
            self._host_bucket = torch.empty((n, d), dtype=bucket.dtype, pin_memory=True)
            self._host_score = torch.empty(n, dtype=torch.int32, pin_memory=True)
            self._host_key = key
        self._host_bucket.copy_(bucket, non_blocking=True)
        self._host_score.copy_(score, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return self._host_score.numpy().copy(), self._host_bucket.numpy().copy()

    def __call__(self, activations: List[np.ndarray]) -> Tuple[np.ndarray, np.ndarray]:
        score, bucket = self.buckets(activations)
        n, d = bucket.shape
        profiles = np.zeros((n, d, self.sections), dtype=bool)
        np.put_along_axis(profiles, np.maximum(bucket, 0).astype(np.int64)[..., None], (bucket >= 0)[..., None], axis=2)
        return score.astype(_score_dtype(d * self.sections)), profiles


class NBC(CoverageMethod):
    """Neuron boundary coverage: below min - s*std or above max + s*std."""

    def __init__(self, mins, maxs, stds, scaler: float):
        super().__init__()
        lo = np.concatenate([np.asarray(l).flatten() for l in mins])
        hi = np.concatenate([np.asarray(l).flatten() for l in maxs])
        sd = np.concatenate([np.asarray(l).flatten() for l in stds])
        self.min_boundaries, self.max_boundaries = lo - scaler * sd, hi + scaler * sd

    def __call__(self, activations):
        act = flatten_layers(activations)
        prof = np.stack([act <= self.min_boundaries, act >= self.max_boundaries], axis=-1)
        return sum_score(prof), prof


class SNAC(CoverageMethod):
    """Strong neuron activation coverage: at or above max + s*std."""

    def __init__(self, maxs, stds, scaler: float):
        super().__init__()
        hi = np.concatenate([np.asarray(l).flatten() for l in maxs])
        sd = np.concatenate([np.asarray(l).flatten() for l in stds])
        self.max_boundaries = hi + scaler * sd

    def __call__(self, activations):
        prof = flatten_layers(activations) >= self.max_boundaries
        return sum_score(prof), prof


class TKNC(CoverageMethod):
    """Top-k neuron coverage, per layer."""

    def __init__(self, top_neurons: int):
        super().__init__()
        self.top_neurons = top_neurons

    def __call__(self, activations):
        per_layer = []
        for layer in activations:
            flat = layer.reshape((layer.shape[0], -1))
            top = np.argsort(flat, axis=1)[..., -self.top_neurons:]
            mark = np.zeros_like(flat, dtype=bool)
            np.put_along_axis(mark, top, True, axis=1)
            per_layer.append(mark)
        prof = flatten_layers(per_layer)
        return sum_score(prof), prof
