"""Bring-up timing of the DSA / LSA / KMNC paths (CUDA events); not the contract bench."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from oracle import np_oracle  # noqa: E402
from simple_tip_b200 import _lib  # noqa: E402
from simple_tip_b200 import engine as E  # noqa: E402
from simple_tip_b200.core.neuron_coverage import KMNC  # noqa: E402
from simple_tip_b200.core.surprise import DSA, LSA  # noqa: E402


def timed(fn, n=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return min(ts), float(np.median(ts))


def main():
    print(torch.cuda.get_device_name(0))
    xtr, ytr, xte, pte, _ = np_oracle.synth_clusters(60000, 10000, 128, 10, seed=2)
    t0 = time.time()
    sa = DSA(xtr, ytr)
    torch.cuda.synchronize()
    print(f"DSA fit {time.time() - t0:.3f}s")
    eng = sa._engine
    order, q_off = E.class_layout(pte, 10)
    x = E.to_device(xte, eng.dev).index_select(0, torch.from_numpy(order).to(eng.dev))
    qc = torch.from_numpy(pte[order].astype(np.int32)).to(eng.dev)
    for use_filter in (True,):
        mn, med = timed(lambda: E.dsa_distances(eng, x, qc, q_off, None, use_filter))
        print(f"DSA C2 device-resident (filter={use_filter}): min {mn:.3f} ms median {med:.3f} ms -> {10000 / mn * 1e3:.3e} inputs/s")
    mn, med = timed(lambda: eng.search(x, qc, q_off, _lib.RANGE_SAME_CLASS))
    print(f"  stage-1 search: {mn:.3f} ms")
    w = eng.gather(eng.search(x, qc, q_off, _lib.RANGE_SAME_CLASS)[1])
    mn, med = timed(lambda: eng.search(w, qc, q_off, _lib.RANGE_OTHER_CLASSES))
    print(f"  stage-2 search: {mn:.3f} ms")
    print("  stats (exhaustive rows, candidates):", eng.stats.cpu().numpy())
    E.dsa_distances(eng, x, qc, q_off, None, True)
    for mode, (cnt, idx) in eng.last_cand_cnt_by_mode.items():
        c = cnt.cpu().numpy()
        masks = idx.cpu().numpy()[:, :, 1].view(np.uint32)
        bits = np.array([sum(bin(int(v)).count("1") for v in masks[i, :min(c[i], masks.shape[1])]) for i in range(0, len(c), 7)])
        print(f"  mode {mode}: candidate chunks per query mean {c.mean():.2f} p50 {np.median(c):.0f} p99 {np.percentile(c, 99):.0f} "
              f"max {c.max()};  rows to re-rank per query mean {bits.mean():.1f} p50 {np.median(bits):.0f} p99 {np.percentile(bits, 99):.0f} max {bits.max()}")
    mn, med = timed(lambda: sa(xte, pte), n=5)
    print(f"DSA C2 end-to-end (host numpy in/out): min {mn:.3f} ms median {med:.3f} ms")
    t0 = time.time(); sa(xte, pte); print(f"  wall {1e3 * (time.time() - t0):.3f} ms")
    if os.environ.get("QUICK_DSA_ONLY"):
        return

    xtr, _, xte, _, _ = np_oracle.synth_clusters(60000, 10000, 256, 10, seed=3, spread=1.0)
    t0 = time.time()
    lsa = LSA(xtr)
    torch.cuda.synchronize()
    print(f"LSA fit {time.time() - t0:.3f}s")
    mn, med = timed(lambda: lsa(xte), n=5)
    print(f"LSA C3 end-to-end: min {mn:.3f} ms median {med:.3f} ms -> {10000 / mn * 1e3:.3e} inputs/s")
    kde = lsa.kde
    xd = E.to_device(xte, kde._engine.dev)
    q = E.whiten(xd, None, kde._mu_dev, kde._w_dev)
    mn, med = timed(lambda: E.whiten(xd, None, kde._mu_dev, kde._w_dev))
    print(f"  whiten: {mn:.3f} ms")
    mn, med = timed(lambda: kde._engine.log_kernel_sum(q))
    print(f"  lse kernel path: {mn:.3f} ms")

    act, mins, maxs = np_oracle.synth_relu(10000, 4096, seed=4)
    km = KMNC([mins], [maxs], 1000)
    mn, med = timed(lambda: km.buckets([act]), n=3, warm=1)
    print(f"KMNC C4 end-to-end: min {mn:.3f} ms")
    lib = _lib.load()
    a_dev = E.to_device(act, E.require_cuda())
    lo, jp = km._dev_stats
    bucket = torch.empty((10000, 4096), dtype=torch.int16, device=a_dev.device)
    score = torch.empty(10000, dtype=torch.int32, device=a_dev.device)
    f = lambda: lib.tip_kmnc(E._p(a_dev), 0, 10000, 4096, E._p(lo), E._p(jp), 0, 1000, E._p(bucket), 3, E._p(score), E._stream())
    mn, med = timed(f, n=10)
    gb = (act.nbytes + bucket.numel() * 2 + 2 * 4096 * 4 + 40000) / 1e9
    print(f"KMNC C4 kernel: min {mn * 1e3:.1f} us -> {gb / (mn * 1e-3):.0f} GB/s")


if __name__ == "__main__":
    main()
