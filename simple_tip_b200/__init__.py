"""simple_tip_b200 — B200-native scoring engine behind the dnn-tip prioritizer API.

`simple_tip_b200.core` mirrors the reference's `src.core` (DSA / LSA / MultiModalSA, KMNC,
DeepGini, apfd_from_order, ctm / cam, Timer); the arithmetic runs in libb200tip.so
(hand-written sm_100a CUDA, C ABI in include/b200tip.h) through ctypes.
"""
__version__ = "0.1.0"
