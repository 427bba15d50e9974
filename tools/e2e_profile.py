"""Host-side breakdown of DSA.__call__ (bring-up tool)."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from oracle import np_oracle  # noqa: E402
from simple_tip_b200 import engine as E  # noqa: E402
from simple_tip_b200.core import surprise as S  # noqa: E402

xtr, ytr, xte, pte, _ = np_oracle.synth_clusters(60000, 10000, 128, 10, seed=2)
sa = S.DSA(xtr, ytr)
pin = torch.from_numpy(xte).pin_memory().numpy()
for _ in range(5):
    sa(pin, pte)


def T():
    torch.cuda.synchronize()
    return time.perf_counter()


reps = 20
acc = {}
eng = sa._engine
for _ in range(reps):
    t = [T()]
    target_pred = S._class_predictions(pte); t.append(T())
    x_all = eng.input_buffer(10000, torch.float32)
    x_all.copy_(torch.from_numpy(pin), non_blocking=True); t.append(T())
    order, q_off = E.class_layout(target_pred, 10); t.append(T())
    plan = E.dsa_plan(eng, int(order.size), q_off, x_all.dtype, True, None, n_total=10000); t.append(T())
    np.copyto(plan.idx_host.numpy(), order, casting="unsafe")
    plan.idx.copy_(plan.idx_host, non_blocking=True); t.append(T())
    plan.graph.replay(); t.append(T())
    plan.out_host.copy_(plan.out, non_blocking=True); t.append(T())
    res = plan.out_host.numpy()
    dsa = res[3].copy(); t.append(T())
    names = ["class_predictions", "H2D traces (pinned)", "class_layout", "plan lookup", "H2D order", "graph replay",
             "D2H (pinned)", "numpy finish"]
    for n, d in zip(names, np.diff(t)):
        acc[n] = acc.get(n, 0) + d
for n, v in acc.items():
    print(f"{n:24s} {1e6 * v / reps:8.1f} us")
t0 = T()
for _ in range(reps):
    sa(pin, pte)
print("full __call__", 1e6 * (T() - t0) / reps, "us")
