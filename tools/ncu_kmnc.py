"""Target for `ncu -k regex:kmnc_ --launch-skip 2 --launch-count 1`: the C4 KMNC kernel on resident data."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from oracle import np_oracle  # noqa: E402
from simple_tip_b200 import _lib  # noqa: E402
from simple_tip_b200 import engine as E  # noqa: E402
from simple_tip_b200.core.neuron_coverage import KMNC  # noqa: E402

act, mins, maxs = np_oracle.synth_relu(10000, 4096, seed=4)
km = KMNC([mins], [maxs], 1000)
lib = _lib.load()
dev = E.require_cuda()
a_dev = E.to_device(act, dev)
km.buckets([act[:64]])           # builds the device statistics
lo, jp = km._dev_stats
bucket = torch.empty((10000, 4096), dtype=torch.int16, device=dev)
score = torch.empty(10000, dtype=torch.int32, device=dev)
ts = []
for _ in range(6):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    lib.tip_kmnc(E._p(a_dev), 0, 10000, 4096, E._p(lo), E._p(jp), 0, 1000, E._p(bucket), 3, E._p(score), E._stream())
    b.record()
    torch.cuda.synchronize()
    ts.append(a.elapsed_time(b))
gb = (act.nbytes + bucket.numel() * 2 + 2 * 4096 * 4 + 40000) / 1e9
print(f"KMNC C4 kernel (+memset): min {min(ts) * 1e3:.1f} us -> {gb / (min(ts) * 1e-3):.0f} GB/s")
