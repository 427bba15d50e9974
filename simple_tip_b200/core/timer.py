"""Wall-clock accumulator — reference: src/core/timer.py:6-50 (start/stop/get, context
manager, `timed` decorator; RuntimeError on double start / stop-without-start,
RuntimeWarning when read while running)."""
import time
import warnings
from functools import wraps


class Timer:
    def __init__(self, start: bool = False):
        self._total = 0
        self._running_since = None
        if start:
            self.start()

    def start(self):
        if self._running_since is not None:
            raise RuntimeError("Timer is already started")
        self._running_since = time.time()

    def stop(self):
        if self._running_since is None:
            raise RuntimeError("Timer is not started")
        self._total += time.time() - self._running_since
        self._running_since = None

    def get(self):
        if self._running_since is not None:
            warnings.warn("Timer is not stopped", RuntimeWarning)
        return self._total

    def timed(self, fn):
        @wraps(fn)
        def wrapped(*args, **kwargs):
            with self:
                return fn(*args, **kwargs)

        return wrapped

    def __enter__(self):
        self.start()

    def __exit__(self, exc_type, exc_val, exc_tb):
        self.stop()
