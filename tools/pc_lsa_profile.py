"""Where does a per-class LSA call (MultiModalSA.build_by_class(LSA), handler_surprise.py:26) spend its time?
Wall-clock per modal with the fast operand pass on / off (bring-up tool, GPU box)."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from oracle import np_oracle  # noqa: E402
from simple_tip_b200.core.surprise import LSA, MultiModalSA  # noqa: E402

xtr, ytr, xte, pte, _ = np_oracle.synth_clusters(60000, 10000, 256, 10, seed=3, spread=1.0)
pc = MultiModalSA.build_by_class(xtr, ytr, lambda x, y: LSA(x))
for mode in ("auto", "slow"):
    if mode == "slow":
        for sa in pc.modal_sa.values():
            sa.kde._engine.fast_ok = False
    for _ in range(3):
        pc(xte, pte)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(5):
        pc(xte, pte)
    torch.cuda.synchronize()
    print(mode, "pc-lsa ms per call:", 1e3 * (time.perf_counter() - t0) / 5)
    for c, sa in list(pc.modal_sa.items())[:3]:
        rows = xte[pte == c]
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(5):
            sa(rows)
        torch.cuda.synchronize()
        print("  class", c, rows.shape, "ms:", 1e3 * (time.perf_counter() - t0) / 5, getattr(sa.kde, "last_operands", None),
              getattr(sa.kde, "last_fast_check", None))
# bench-like variants: bf16-rounded traces, pinned source, L2 flush + CUDA-event timing as bench.py does
rb = lambda a: torch.from_numpy(a).to(torch.bfloat16).to(torch.float32).numpy()
xtr2, xte2 = rb(xtr), rb(xte)
pc2 = MultiModalSA.build_by_class(xtr2, ytr, lambda x, y: LSA(x))
pinned = torch.from_numpy(xte2).pin_memory().numpy()
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for name, src in (("pageable", xte2), ("pinned", pinned)):
    for _ in range(3):
        pc2(src, pte)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(5):
        pc2(src, pte)
    torch.cuda.synchronize()
    wall = 1e3 * (time.perf_counter() - t0) / 5
    ev = []
    for _ in range(5):
        flush.fill_(1)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        pc2(src, pte)
        b.record()
        torch.cuda.synchronize()
        ev.append(a.elapsed_time(b))
    print("bench-like", name, "wall ms", wall, "event ms (after L2 flush)", sum(ev) / 5,
          [getattr(sa.kde, "last_operands", None) for sa in list(pc2.modal_sa.values())[:3]])
from torch.profiler import ProfilerActivity, profile

for sa in pc.modal_sa.values():
    sa.kde._engine.fast_ok = sa.kde._engine.t_pack_f16 is not None
sa = pc.modal_sa[0]
rows = xte[pte == 0]
sa(rows)
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    sa(rows)
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=25))
