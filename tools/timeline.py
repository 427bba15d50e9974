"""Per-tile timeline of block 0 of the resident-query filter kernel (bring-up tool).
Since the quarter-accumulator rewrite the two MMA warps stamp their own half: "issue half h" spans both quarters of
that half (wait for the second quarter's accumulator included), "wait TMEM half h empty" is the first quarter's wait."""
import ctypes as C
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from oracle import np_oracle  # noqa: E402
from simple_tip_b200 import _lib  # noqa: E402
from simple_tip_b200 import engine as E  # noqa: E402
from simple_tip_b200.core.surprise import DSA  # noqa: E402

xtr, ytr, xte, pte, _ = np_oracle.synth_clusters(60000, 10000, 128, 10, seed=2)
sa = DSA(xtr, ytr)
eng = sa._engine
order, q_off = E.class_layout(pte, 10)
x = E.to_device(xte, eng.dev).index_select(0, torch.from_numpy(order).to(eng.dev))
qc = torch.from_numpy(pte[order].astype(np.int32)).to(eng.dev)
lib = _lib.load()
for mode in (_lib.RANGE_OTHER_CLASSES, _lib.RANGE_SAME_CLASS):
    for _ in range(2):
        eng.search(x, qc, q_off, mode)
    tiles = 400
    buf = torch.zeros((tiles, 16), dtype=torch.int64, device=eng.dev)
    lib.tip_debug_timeline(C.c_void_p(buf.data_ptr()), tiles)
    eng.search(x, qc, q_off, mode)
    torch.cuda.synchronize()
    lib.tip_debug_timeline(None, 0)
    t = buf.cpu().numpy()
    n = int((t[:, 3] > 0).sum())
    t = t[:n].astype(np.float64)
    t0 = t[0, 0]
    print(f"mode {mode}: {n} tiles by block 0, total {(t[n-1, 7] - t0):.0f} cycles -> {(t[n-1,7]-t0)/n:.0f} cycles/tile")
    d = lambda a, b: t[:, a] - t[:, b]
    def st(name, v):
        print(f"  {name:42s} mean {v.mean():8.0f}  p50 {np.median(v):8.0f}  p90 {np.percentile(v, 90):8.0f}  max {v.max():8.0f}")
    st("MMA: wait train tile (1-0)", d(1, 0))
    st("MMA: wait TMEM half 0 empty (2-1)", d(2, 1))
    st("MMA: issue half 0 + commit (3-2)", d(3, 2))
    st("MMA: wait TMEM half 1 empty (11-3)", d(11, 3))
    st("MMA: issue half 1 + commits (12-11)", d(12, 11))
    st("MMA: tile period (0[t+1]-0[t])", t[1:, 0] - t[:-1, 0])
    st("EPI h0: wait accumulator (5-4)", d(5, 4))
    st("EPI h0: accumulator ready -> TMEM release (6-5)", d(6, 5))
    st("EPI h0: release -> tile end (7-6)", d(7, 6))
    st("EPI h0: tile period", t[1:, 4] - t[:-1, 4])
    st("commit(tfull0) -> EPI sees it (5-3)", d(5, 3))
    st("EPI release -> MMA h0 of next tile starts ((2)[t+1]-(6)[t])", t[1:, 2] - t[:-1, 6])
    st("TMA: wait stage empty (9-8)", d(9, 8))
    st("TMA: issue (10-9)", d(10, 9))
    print("  first 6 tiles (relative cycles):")
    for i in range(min(6, n)):
        print("   ", " ".join(f"{int(v - t0):7d}" for v in t[i, :13]))
    rel = (t[:, 6] - t[:, 5])
    print("  EPI ready->release per tile:", " ".join(f"{int(v)}" for v in rel[:170]))
