"""Secondary configurations of BASELINE.json on one B200 (not the contract bench):
C3 LSA 10k x 60k x 256, C4 KMNC 10k x 4096 x 1000 sections, and the per-GPU slice of C5
(DSA, D = 2048, 1000 classes, N_train/8 = 160k rows) with on-device synthetic traces."""
import json
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
from simple_tip_b200.core.surprise import LSA  # noqa: E402

PEAK_TF, PEAK_GBS = 1689.8, 6568.4
if os.path.exists("MEASURED_PEAKS.json"):
    p = json.load(open("MEASURED_PEAKS.json"))
    PEAK_TF, PEAK_GBS = p["bf16_tflops"], p["hbm_gbs"]


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
    return float(np.median(ts))


out = {}
dev = E.require_cuda()

# ---- C3: LSA ---------------------------------------------------------------------------------
xtr, _, xte, _, _ = np_oracle.synth_clusters(60000, 10000, 256, 10, seed=3, spread=1.0)
r = lambda a: torch.from_numpy(a).to(torch.bfloat16).to(torch.float32).numpy()      # bf16-stored traces
xtr, xte = r(xtr), r(xte)
t0 = time.time()
lsa = LSA(xtr)
torch.cuda.synchronize()
fit_s = time.time() - t0
pin = torch.from_numpy(xte).pin_memory().numpy()
e2e = timed(lambda: lsa(pin))
kde = lsa.kde
xd = E.to_device(xte, dev)
q = E.whiten(xd, None, kde._mu_dev, kde._w_dev)
E.PROFILE = None
dev_ms = timed(lambda: kde._engine.log_kernel_sum(E.whiten(xd, None, kde._mu_dev, kde._w_dev)))
flops = 2.0 * 256 * 10000 * 60000
sub = np.random.default_rng(0).choice(10000, 200, replace=False)
want = np_oracle.lsa_oracle(xtr, xte[sub])
got = lsa(xte)[sub]
err = float(np.max(np.abs(got - want) / np.maximum(1.0, np.abs(want))))
out["c3_lsa"] = {"inputs_per_s_device": 10000 / (dev_ms * 1e-3), "inputs_per_s_e2e": 10000 / (e2e * 1e-3),
                 "ms_device": dev_ms, "ms_e2e": e2e, "fit_s": fit_s, "algorithmic_tflops": flops / (dev_ms * 1e-3) / 1e12,
                 "frac_of_bf16_peak": flops / (dev_ms * 1e-3) / 1e12 / PEAK_TF, "max_rel_err_vs_oracle_200_rows": err}
print(json.dumps({"c3_lsa": out["c3_lsa"]}))

# ---- C4: KMNC --------------------------------------------------------------------------------
act, mins, maxs = np_oracle.synth_relu(10000, 4096, seed=4)
km = KMNC([mins], [maxs], 1000)
pin = torch.from_numpy(act).pin_memory().numpy()
e2e = timed(lambda: km.buckets([pin]), n=3, warm=1)
lib = _lib.load()
a_dev = E.to_device(act, dev)
lo, jp = km._dev_stats
bucket = torch.empty((10000, 4096), dtype=torch.int16, device=dev)
score = torch.empty(10000, dtype=torch.int32, device=dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def k():
    lib.tip_kmnc(E._p(a_dev), 0, 10000, 4096, E._p(lo), E._p(jp), 0, 1000, E._p(bucket), 3, E._p(score), E._stream())


ts = []
for _ in range(10):
    flush.fill_(0)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); k(); b.record(); torch.cuda.synchronize()
    ts.append(a.elapsed_time(b))
ms = float(np.median(ts))
nbytes = act.nbytes + 10000 * 4096 * 2 + 2 * 4096 * 4 + 10000 * 4
out["c4_kmnc"] = {"inputs_per_s_device": 10000 / (ms * 1e-3), "inputs_per_s_e2e": 10000 / (e2e * 1e-3), "us_kernel": ms * 1e3,
                  "algorithmic_GBps": nbytes / (ms * 1e-3) / 1e9, "frac_of_hbm_peak": nbytes / (ms * 1e-3) / 1e9 / PEAK_GBS}
print(json.dumps({"c4_kmnc": out["c4_kmnc"]}))

# ---- C5 per-GPU slice: DSA D = 2048 ---------------------------------------------------------
del a_dev, bucket, flush
n_train, n_test, d, classes = 160000, int(os.environ.get("C5_TESTS", "20000")), 2048, 1000
g = torch.Generator(device=dev).manual_seed(5)
centres = torch.randn((classes, d), generator=g, device=dev) * 0.5
ytr = torch.arange(n_train, device=dev) % classes                       # 160 rows per class on this shard
order = torch.argsort(ytr, stable=True)
t_sorted = (centres[ytr[order]] + torch.randn((n_train, d), generator=g, device=dev))
class_off = np.arange(classes + 1, dtype=np.int64) * (n_train // classes)
eng = E.NnEngine(t_sorted, class_off, order.to(torch.int32))
yte = torch.arange(n_test, device=dev) % classes
x = centres[yte] + torch.randn((n_test, d), generator=g, device=dev)
q_order = torch.argsort(yte, stable=True)
x = x[q_order].contiguous()
q_class = yte[q_order].to(torch.int32).contiguous()
q_off = np.arange(classes + 1, dtype=np.int64) * (n_test // classes)
torch.cuda.synchronize()
E.PROFILE = []
ms = timed(lambda: E.dsa_distances(eng, x, q_class, q_off), n=3, warm=1)
prof = E.PROFILE
E.PROFILE = None
by = {}
for name, fl, e0, e1 in prof:
    by.setdefault(name, []).append((fl, e0.elapsed_time(e1)))
kern = {kk: {"ms": float(np.mean([t for _, t in v])), "algorithmic_tflops": float(np.mean([f for f, _ in v])) / (np.mean([t for _, t in v]) * 1e-3) / 1e12}
        for kk, v in by.items()}
a, b, gid = E.dsa_distances(eng, x, q_class, q_off)
rows = torch.from_numpy(np.random.default_rng(1).choice(n_test, 64, replace=False)).to(dev)
xs, qs = x[rows].contiguous(), q_class[rows].contiguous()
# exhaustive check of a row subset: queries keep their class, one-class-per-query layout not needed for the scan
sub_order = torch.argsort(qs, stable=True)
xs, qs, rows = xs[sub_order].contiguous(), qs[sub_order].contiguous(), rows[sub_order]
cnt = np.bincount(qs.cpu().numpy(), minlength=classes)
sub_off = np.concatenate([[0], np.cumsum(cnt)]).astype(np.int64)
ea, eb, eg = E.dsa_distances(eng, xs, qs, sub_off, None, use_filter=False)
ok = bool(torch.equal(ea, a[rows]) and torch.equal(eb, b[rows]) and torch.equal(eg.long(), gid[rows].long()))
stats = eng.stats.cpu().numpy().tolist()
flops = 2.0 * d * n_test * n_train
out["c5_slice_dsa"] = {"n_train_shard": n_train, "n_test": n_test, "d": d, "classes": classes, "ms_per_pass": ms,
                       "inputs_per_s": n_test / (ms * 1e-3), "algorithmic_tflops_whole_pass": flops / (ms * 1e-3) / 1e12,
                       "kernels": kern, "subset_matches_exhaustive_scan": ok, "stats_exhaustive_rows_candidates": stats,
                       "extrapolated_s_for_100k_tests": ms * 1e-3 * 100000 / n_test}
print(json.dumps({"c5_slice_dsa": out["c5_slice_dsa"]}))
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/bench_extra.json", "w"), indent=1)

# ---- DeepGini bandwidth variants --------------------------------------------
from simple_tip_b200.core.deepgini import DeepGini  # noqa: E402

res = {}
for n, c in ((10_000_000, 10), (1_000_000, 1000)):
    p = torch.rand((n, c), device=dev, dtype=torch.float32)
    p /= p.sum(dim=1, keepdim=True)
    pred = torch.empty(n, dtype=torch.int32, device=dev)
    gini = torch.empty(n, dtype=torch.float32, device=dev)
    f = lambda: lib.tip_deepgini(E._p(p), 0, n, c, E._p(pred), E._p(gini), E._stream())
    ms = timed(f, n=5, warm=2)
    nbytes = n * c * 4 + n * 8
    sub = p[:2000].cpu().numpy()
    wp, wg = np_oracle.deepgini_oracle(sub)
    f()
    torch.cuda.synchronize()
    ok = bool(np.array_equal(pred[:2000].cpu().numpy(), wp) and np.array_equal(gini[:2000].cpu().numpy(), wg))
    res[f"{n}x{c}"] = {"ms": ms, "GBps": nbytes / (ms * 1e-3) / 1e9, "frac_of_hbm_peak": nbytes / (ms * 1e-3) / 1e9 / PEAK_GBS,
                       "bit_exact_vs_oracle_2000_rows": ok}
    del p, pred, gini
out["deepgini"] = res
print(json.dumps({"deepgini": res}))
json.dump(out, open("gpurun_out/bench_extra.json", "w"), indent=1)
