"""APFD (Feng et al., DeepGini) — reference: src/core/apfd.py:8-19.  O(n) host arithmetic; the
evaluator the parity metric is defined with (eval_apfd_table.py:86,101)."""
from typing import List, Union

import numpy as np


def apfd_from_order(is_fault, index_order: Union[List[int], np.ndarray]) -> float:
    """1 - sum(rank of each fault) / (k*n) + 1/(2n), ranks counted from 1."""
    assert is_fault.ndim == 1, "at the moment, only unique faults are supported"
    hits = np.where(is_fault[index_order] == 1)[0]
    k = np.count_nonzero(is_fault)
    n = is_fault.shape[0]
    return 1 - (np.sum(hits + 1) / (k * n)) + (1 / (2 * n))
