"""Target for `ncu -k regex:pair -s 2 -c 1`: the streaming tensor-core filter at C5's trace width on one GPU
(DSA stage 2, 4096 test x 160k train x 2048-d, 1000 classes — a slice small enough for ncu's replays).
   ncu --set full --clock-control none -k regex:pair.*kernel -s 3 -c 1 -o gpurun_out/prof_c5 python tools/ncu_c5.py"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from oracle import synth_traces as ST  # noqa: E402
from simple_tip_b200 import _lib  # noqa: E402
from simple_tip_b200 import engine as E  # noqa: E402

dev = E.require_cuda()
n_train, n_test, d, classes = 160000, 4096, 2048, 1000
t = torch.empty((n_train, d), dtype=torch.float32, device=dev)
ST.fill(t, 0, d, classes, 5, 0)
cls = torch.arange(n_train, device=dev) % classes
order = torch.argsort(cls, stable=True)
eng = E.NnEngine(t.index_select(0, order), np.arange(classes + 1, dtype=np.int64) * (n_train // classes), order.to(torch.int32))
x = torch.empty((n_test, d), dtype=torch.float32, device=dev)
ST.fill(x, 0, d, classes, 5, 1)
qc = (torch.arange(n_test, device=dev) % classes)
qo = torch.argsort(qc, stable=True)
x, qc = x.index_select(0, qo).contiguous(), qc[qo].to(torch.int32).contiguous()
q_off = np.concatenate([[0], np.cumsum(np.bincount(qc.cpu().numpy(), minlength=classes))]).astype(np.int64)
for _ in range(3):
    eng.search(x, qc, q_off, _lib.RANGE_OTHER_CLASSES)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
eng.search(x, qc, q_off, _lib.RANGE_OTHER_CLASSES)      # profiled launch
b.record()
torch.cuda.synchronize()
print("search (filter + re-rank) ms:", a.elapsed_time(b), "algorithmic TFLOP/s incl. re-rank:",
      2.0 * d * n_test * (n_train - n_train // classes) / (a.elapsed_time(b) * 1e-3) / 1e12)
