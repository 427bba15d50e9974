"""Streaming min / max / std of activation layers on the GPU — the fit step of KMNC / NBC / SNAC.

Mirror of `/root/reference/src/dnn_test_prio/aggregate_statistics.py:12-67` (`AggregateStatisticsCollector`:
`track(badge)` per batch of the training-set walk, `get()` -> (mins, maxs, stds) as lists of layer-shaped
arrays; handler_coverage.py:33-46 is the caller).  The reference keeps `np.minimum` / `np.maximum` running
arrays and one `welford.Welford` per layer (welford==0.2.5, requirements.txt:6): initialised with the first
sample, then `add_all(badge)` = a sequential loop of `add` over the samples with the mean and the sum of
squared deviations held in the activation dtype; `get()` returns `sqrt(var_s)` = sqrt(s / (count - 1)).
Here every layer's state (mean, m2, min, max per neuron) lives in HBM and one kernel per badge and layer
(libb200tip `tip_stats_update`: one thread per neuron walking down the badge in order) applies exactly that
arithmetic, so the statistics are bit-identical; the final `sqrt(s / (count - 1))` is the same NumPy expression
on the host.  Badges may be NumPy arrays or torch CUDA tensors (straight from a forward hook).
"""
from __future__ import annotations

from typing import List, Tuple

import numpy as np

from .timer import Timer

AggStats = Tuple[List[np.ndarray], List[np.ndarray], List[np.ndarray]]


class AggregateStatisticsCollector:
    """A timed, online calculator of mins, maxs and stds of equally shaped arrays (GPU)."""

    def __init__(self):
        self.is_initialized = False
        self.done = False
        self._state = []          # per layer: dict(shape, dtype, mean, m2, mn, mx)
        self._count = 0
        # the reference times min / max / welford separately; here they are one fused pass, booked on the
        # welford timer (its callers only add the three up: handler_coverage.py:49-90)
        self.min_timer = Timer()
        self.max_timer = Timer()
        self.welford_timer = Timer()

    def _initialize(self, badge) -> None:
        import torch

        from .. import engine as E

        dev = E.require_cuda()
        for layer in badge:
            shape = tuple(layer.shape[1:])
            d = int(np.prod(shape)) if len(shape) else 1
            if isinstance(layer, torch.Tensor):
                dt = torch.float64 if layer.dtype == torch.float64 else torch.float32
            else:
                dt = torch.float64 if np.result_type(layer.dtype, np.float32) == np.float64 else torch.float32
            self._state.append({
                "shape": shape, "dtype": dt, "d": d,
                "mean": torch.zeros(d, dtype=dt, device=dev), "m2": torch.zeros(d, dtype=dt, device=dev),
                "mn": torch.full((d,), float("inf"), dtype=dt, device=dev),
                "mx": torch.full((d,), float("-inf"), dtype=dt, device=dev)})
        self.is_initialized = True

    def track(self, badge) -> None:
        """Pass the next badge of arrays to be included in aggregate metrics."""
        import torch

        from .. import _lib
        from .. import engine as E

        if self.done:
            raise RuntimeError("`get` has been called. calling it multiple times falsifies timer.")
        if not self.is_initialized:
            self._initialize(badge)
        lib = _lib.load()
        n = int(badge[0].shape[0])
        with self.welford_timer:
            for layer, st in zip(badge, self._state):
                if isinstance(layer, torch.Tensor):
                    x = layer.reshape(layer.shape[0], -1).to(st["mean"].device, st["dtype"]).contiguous()
                else:
                    x = E.to_device(np.ascontiguousarray(np.reshape(layer, (layer.shape[0], -1)),
                                                         dtype=E.NP_DTYPE[st["dtype"]]), st["mean"].device)
                assert x.shape[1] == st["d"], "badge layers must keep their shape"
                _lib.check(lib.tip_stats_update(E._p(x), E.tip_dtype(st["dtype"]), x.shape[0], st["d"], self._count,
                                                E._p(st["mean"]), E._p(st["m2"]), E._p(st["mn"]), E._p(st["mx"]),
                                                E._stream()), "tip_stats_update")
            torch.cuda.current_stream().synchronize()
        self._count += n

    def get(self) -> AggStats:
        """Return the aggregated metrics."""
        mins, maxs, stds = [], [], []
        with self.welford_timer:
            for st in self._state:
                shape = st["shape"]
                mins.append(st["mn"].cpu().numpy().reshape(shape))
                maxs.append(st["mx"].cpu().numpy().reshape(shape))
                s = st["m2"].cpu().numpy().reshape(shape)
                if self._count <= 1:                      # welford: var_s is NaN below two samples
                    var = np.full(shape, np.nan)
                else:
                    var = s / (self._count - 1)
                stds.append(np.sqrt(var))
        return mins, maxs, stds
