"""Target for `ncu -k regex:pair_rs_kernel --launch-skip 4 --launch-count 2`: C2 stage-2 then stage-1 filter."""
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
w = eng.gather(eng.search(x, qc, q_off, _lib.RANGE_SAME_CLASS)[1])     # pair_rs launch 1
eng.search(w, qc, q_off, _lib.RANGE_OTHER_CLASSES)                      # 2
eng.search(x, qc, q_off, _lib.RANGE_SAME_CLASS)                         # 3
eng.search(w, qc, q_off, _lib.RANGE_OTHER_CLASSES)                      # 4
torch.cuda.synchronize()
eng.search(w, qc, q_off, _lib.RANGE_OTHER_CLASSES)                      # 5: profiled (stage 2)
eng.search(x, qc, q_off, _lib.RANGE_SAME_CLASS)                         # 6: profiled (stage 1)
torch.cuda.synchronize()
