"""Target for `ncu -k regex:pair -s <skip> -c <count>`: the LSA log-sum-exp tensor-core launches at C3
(10k x 60k x 256): launch order after warm-up = [fast fp16 pass, three-segment pass].
   ncu --set full --clock-control none --import-source on -k regex:pair.*kernel -s 4 -c 2 -o gpurun_out/prof_lse python tools/ncu_lse.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from oracle import np_oracle  # noqa: E402
from simple_tip_b200 import engine as E  # noqa: E402
from simple_tip_b200.core.surprise import LSA  # noqa: E402

xtr, _, xte, _, _ = np_oracle.synth_clusters(60000, 10000, 256, 10, seed=3, spread=1.0)
sa = LSA(xtr)
kde = sa.kde
eng = kde._engine
q = E.whiten(E.to_device(xte, eng.dev), None, kde._mu_dev, kde._w_dev)
for _ in range(2):                       # 4 launches of warm-up
    if eng.t_pack_f16 is not None:
        eng.log_kernel_sum(q, fast=True)
    eng.log_kernel_sum(q)
torch.cuda.synchronize()
if eng.t_pack_f16 is not None:
    eng.log_kernel_sum(q, fast=True)     # profiled 1
eng.log_kernel_sum(q)                    # profiled 2
torch.cuda.synchronize()
