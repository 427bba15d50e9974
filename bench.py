"""bench.py — DSA inputs prioritized / second on B200 (BASELINE.json metric), driver contract.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload c1|c2|c3|c4]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" is one pass of the hot path over one batch of synthetic test traces:
  c2 (default, the configuration the metric is quoted on): DSA, 10 000 test x 60 000 train x
      128-d float32 Gaussian-cluster traces, 10 classes (seed 2).
      With N GPUs the training set is sharded over the ranks (N_train axis, north_star) and the
      test batch grows to 10 000 x N so that per-GPU work is fixed ("scaling": "weak"); the
      per-shard minima are merged with NCCL all-reduces (engine.TrainShardComm).
  c1 / c3 / c4: DeepGini 10k x 10 (+APFD), LSA 10k x 60k x 256 and KMNC 10k x 4096 x 1000 sections,
      single GPU, same JSON shape (each with its own cpu_baseline from the oracle port).

`value` times the device-resident path (test traces already in HBM, result left in HBM);
`e2e` times the reference-facing call `DSA.__call__(numpy, numpy) -> numpy` from pinned host
memory, host<->device copies inside the timed region.  Steps are timed individually with CUDA
events on the launching stream, L2 is flushed (256 MiB write) between steps, the sum over K
steps is max-reduced over ranks.  `--impl reference` times the reference's own NumPy algorithm
(oracle port: same NumPy expressions, same 5-thread badge pool, surprise.py:599) on the host.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

C2 = dict(n_train=60000, n_test=10000, d=128, classes=10, seed=2)
METRIC = "dsa_inputs_prioritized_per_sec"


def _peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return float(p["bf16_tflops"]), float(p["hbm_gbs"]), "measured (MEASURED_PEAKS.json, burst)"
    return 1590.0, 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu, self.rows, self.proc = gpu_index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.gpu), "-lms", "100"], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm = [float(r[1]) for r in self.rows if len(r) >= 8 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) >= 8 and r[2].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 8 for i in range(4) if r[4 + i].lower() == "active"})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


# ----------------------------------------------------------------------------------------------
# reference arm / CPU baseline (the only place bench.py executes oracle/)
# ----------------------------------------------------------------------------------------------
def _ref_sample(xtr, ytr, xte, pte, n_inputs: int, threads: int = 5):
    from oracle import np_oracle

    t0 = time.perf_counter()
    np_oracle.dsa_oracle(xtr, ytr, xte[:n_inputs], pte[:n_inputs], badge_size=10, threads=threads)
    return time.perf_counter() - t0


def cpu_baseline(xtr, ytr, xte, pte, budget_inputs: int = 160):
    """Oracle port of the reference DSA (same NumPy expressions, 5 badge threads) on a bounded
    prefix of the same test set; cost is linear in the number of inputs (independent badges)."""
    from oracle import c_oracle

    dt = _ref_sample(xtr, ytr, xte, pte, budget_inputs)
    out = {"value": budget_inputs / dt, "unit": "inputs/s", "cores": 5, "kind": "port",
           "sample": f"first {budget_inputs} of the {xte.shape[0]} test inputs vs all {xtr.shape[0]} train rows, "
                     f"{dt:.1f} s; 5 badge threads as surprise.py:599; host has {os.cpu_count()} cpus"}
    try:  # stronger, non-reference CPU number for context: C/OpenMP port on all cores
        c_oracle.build()
        n_c = min(xte.shape[0], 2000)
        t0 = time.perf_counter()
        c_oracle.dsa(xtr, ytr, xte[:n_c], pte[:n_c])
        dc = time.perf_counter() - t0
        out["c_port_all_cores"] = {"value": n_c / dc, "cores": c_oracle.max_threads(),
                                   "sample": f"{n_c} inputs, {dc:.1f} s, OpenMP brute force in NumPy's summation order"}
    except Exception as e:  # pragma: no cover
        out["c_port_all_cores"] = {"error": str(e)}
    return out


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import np_oracle

    xtr, ytr, xte, pte, _ = np_oracle.synth_clusters(**C2)
    per_step = 48                                   # bounded sample of the workload per step (~3-4 s)
    for _ in range(args.warmup):
        _ref_sample(xtr, ytr, xte, pte, per_step)
    t = [_ref_sample(xtr, ytr, xte, pte, per_step) for _ in range(args.steps)]
    total = float(np.sum(t))
    value = per_step * args.steps / total
    line = {"metric": METRIC, "value": value, "unit": "inputs/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "impl": "reference",
            "config": {"workload": "C2: DSA 10k test x 60k train x 128-d float32, 10 classes (seed 2)",
                       "step": f"bounded sample: {per_step} test inputs per step vs all 60k train rows"},
            "cpu_baseline": {"value": value, "unit": "inputs/s", "cores": 5, "kind": "port",
                             "sample": f"{per_step} inputs per step; oracle port of surprise.py:558-651 (NumPy, "
                                       f"5 badge threads); host has {os.cpu_count()} cpus"},
            "e2e": {"value": value, "unit": "inputs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


# ----------------------------------------------------------------------------------------------
# our arm
# ----------------------------------------------------------------------------------------------
def run_ours(args):
    import torch

    from oracle import np_oracle   # synthetic trace generator + cpu_baseline only
    from simple_tip_b200 import _lib
    from simple_tip_b200 import engine as E
    from simple_tip_b200.core.surprise import DSA

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    # the contract is ONE JSON line on stdout: anything libraries print there (NCCL's version banner)
    # goes to stderr until the line is ready
    sys.stdout.flush()
    saved_stdout = os.dup(1)
    os.dup2(2, 1)
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    comm = None          # N_train-sharded engine (collectives on the data path)
    dist = None          # process group (timing barrier / max over ranks in every multi-GPU mode)
    # How the global batch of 10000 x N test inputs is spread over N GPUs:
    #   test  — every rank scores its own 10000 inputs against a replicated train set (30.7 MB at C2):
    #           independent units, no data-path collective;
    #   train — the train set is split 1/N per rank, every rank sees all 10000 x N inputs and the
    #           per-shard minima / winner rows are all-reduced (the layout C5 needs: tools/c5_multi.py).
    shard = args.shard if args.shard != "auto" else "test"
    if world > 1:
        import torch.distributed as dist

        if os.environ.get("NCCL_DEBUG", "").upper() in ("VERSION", "INFO"):
            os.environ["NCCL_DEBUG"] = "WARN"      # NCCL would print its banner on stdout; the contract is ONE JSON line

        dist.init_process_group("nccl", device_id=dev)
        if shard == "train":
            comm = E.TrainShardComm()
    assert world == args.gpus, f"--gpus {args.gpus} but WORLD_SIZE={world}"

    cfg = dict(C2)
    cfg["n_test"] = C2["n_test"] * world           # weak scaling: 10000 test inputs (x 60000 / N or x 60000 train rows) per GPU
    xtr, ytr, xte, pte, _ = np_oracle.synth_clusters(**cfg)
    n_global = xte.shape[0]
    if world > 1 and shard == "test":
        sl = slice(rank * C2["n_test"], (rank + 1) * C2["n_test"])
        xte, pte = np.ascontiguousarray(xte[sl]), np.ascontiguousarray(pte[sl])
    sa = DSA(xtr, ytr, comm=comm)
    eng = sa._engine
    n_test = xte.shape[0]

    # device-resident inputs (value) and pinned host inputs (e2e)
    order, q_off = E.class_layout(pte, int(sa.num_classes))
    x_sorted = E.to_device(xte, dev).index_select(0, torch.from_numpy(order).to(dev))
    q_class = torch.from_numpy(pte[order].astype(np.int32)).to(dev)
    xte_pinned = torch.from_numpy(xte).pin_memory()
    xte_host = xte_pinned.numpy()
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)

    plan = None
    if sa.use_graphs:
        # steady-state path of DSA.__call__: the search replayed as CUDA graph(s); sharded runs replay
        # one graph per stage with eager NCCL all-reduces in between
        plan = E.dsa_plan(eng, n_test, q_off, x_sorted.dtype, sa.use_filter, comm)
        plan.load_sorted(x_sorted)

    def step_eager():
        a, b, _ = E.dsa_distances(eng, x_sorted, q_class, q_off, comm)
        return a / b

    def step_device():
        if plan is None:
            return step_eager()
        out = plan.run()
        return out[3] if out.shape[0] == 4 else out[0] / out[1]

    def step_e2e():
        return sa(xte_host, pte)

    xte_dev = E.to_device(xte, dev)

    def step_api_device():      # same public call with the traces already in HBM (torch CUDA tensor in)
        return sa(xte_dev, pte)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, profile=False):
        times = []
        E.PROFILE = [] if profile else None
        for _ in range(steps):
            flush.fill_(1)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            if dist is not None:
                dist.barrier()
            a.record()
            fn()
            b.record()
            torch.cuda.synchronize()
            times.append(a.elapsed_time(b))
        prof, E.PROFILE = E.PROFILE, None
        return times, prof

    for _ in range(max(3, args.warmup)):
        step_device()
        step_e2e()
        step_eager()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    t_dev, _ = timed(step_device, args.steps)
    t_e2e, _ = timed(step_e2e, args.steps)
    step_api_device()
    t_api, _ = timed(step_api_device, args.steps)
    # per-kernel durations (roofline) and the launch census come from the same kernels launched
    # eagerly with CUDA events around the tensor-core launches; graph replays launch the same set
    launches0 = _lib.launch_count()
    _, prof = timed(step_eager, args.steps, profile=True)
    launches = _lib.launch_count() - launches0
    barrier()
    clocks = sampler.stop() if rank == 0 else None

    tot = torch.tensor([sum(t_dev), sum(t_e2e)], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(tot, op=dist.ReduceOp.MAX)
    tot_dev_ms, tot_e2e_ms = [float(v) for v in tot.cpu()]

    if rank == 0:
        value = n_global * args.steps / (tot_dev_ms * 1e-3)
        e2e = n_global * args.steps / (tot_e2e_ms * 1e-3)
        tflops_peak, hbm_peak, peak_src = _peaks()
        # dominant kernel: the stage-2 tcgen05 filter launch (other-class columns)
        roof = None
        if prof:
            by = {}
            for name, flops, ev0, ev1 in prof:
                by.setdefault(name, []).append((flops, ev0.elapsed_time(ev1)))
            name = max(by, key=lambda k: np.mean([t for _, t in by[k]]))
            flops = float(np.mean([f for f, _ in by[name]]))
            ms = float(np.mean([t for _, t in by[name]]))
            traffic = None
            tpath = os.path.join(ROOT, "profiles", "roofline_traffic.json")
            if os.path.exists(tpath):
                traffic = json.load(open(tpath)).get(name)
            roof = {"kernel": name, "bound": "tensor", "achieved": flops / (ms * 1e-3) / 1e12, "peak": tflops_peak,
                    "unit": "TFLOP/s", "frac": flops / (ms * 1e-3) / 1e12 / tflops_peak, "traffic": traffic,
                    "peak_source": peak_src, "ms_per_launch": ms,
                    "algorithmic_flops_per_launch": flops,
                    "other_launches_ms": {k: float(np.mean([t for _, t in v])) for k, v in by.items() if k != name}}
        line = {"metric": METRIC, "value": value, "unit": "inputs/s", "n_gpus": world, "steps": args.steps,
                "warmup": max(3, args.warmup), "ms_per_step": tot_dev_ms / args.steps, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": f"C2: DSA {n_global} test x 60000 train x 128-d float32, 10 classes (seed 2)",
                           "parallelism": (f"N_train sharded over {world} GPU(s), every rank scores all 10000 x {world} "
                                           "inputs; all-reduce of per-shard minima and winner rows" if comm is not None else
                                           f"N_test sharded over {world} GPU(s): 10000 inputs per GPU, train set "
                                           "replicated (30.7 MB), no data-path collective" if world > 1 else
                                           "1 GPU"),
                           "filter": "bf16 tcgen05 candidate filter + exact fp32 re-rank (bit-identical to NumPy)",
                           "l2": "flushed between steps (256 MiB write)", "timing": "per-step CUDA events, summed",
                           "launch": ("CUDA-graph replay of the search" if comm is None else
                                      "per-stage CUDA graphs + eager NCCL all-reduces") if plan is not None else "eager launches"},
                "e2e": {"value": e2e, "unit": "inputs/s", "ms_per_step": tot_e2e_ms / args.steps,
                        "h2d_bytes_per_step": int((xte.nbytes + pte.shape[0] * 4) * (world if comm is None else 1)),
                        # dist_a, dist_b, winner index, dsa as float64, summed over the ranks
                        "d2h_bytes_per_step": int(4 * n_test * 8 * (world if comm is None else 1))},
                # the public call with device-resident traces (rank 0's own time, not max-reduced)
                "api_device_inputs": {"ms_per_step": float(np.mean(t_api)), "unit": "ms",
                                      "note": "DSA.__call__(torch CUDA tensor, labels) -> numpy scores"},
                "gpu_launches": int(launches), "clocks": clocks, "roofline": roof}
        if world == 1 and not args.no_cpu:
            line["cpu_baseline"] = cpu_baseline(xtr, ytr, xte, pte)
        sys.stdout.flush()
        os.dup2(saved_stdout, 1)
        print(json.dumps(line), flush=True)
        os.dup2(2, 1)
    if dist is not None:
        dist.destroy_process_group()


# ----------------------------------------------------------------------------------------------
# secondary workloads (single GPU): same JSON shape, selected with --workload
# ----------------------------------------------------------------------------------------------
def run_secondary(args):
    """c1 DeepGini 10k x 10 (+ APFD parity), c3 LSA 10k x 60k x 256 (bf16-stored traces),
    c4 KMNC 10k x 4096 x 1000 sections — BASELINE.json configs 1, 3, 4 on one B200."""
    import torch

    from oracle import c_oracle, np_oracle   # synthetic generators + cpu_baseline only
    from simple_tip_b200 import _lib
    from simple_tip_b200 import engine as E

    sys.stdout.flush()
    saved_stdout = os.dup(1)
    os.dup2(2, 1)
    torch.cuda.set_device(0)
    dev = torch.device("cuda", 0)
    lib = _lib.load()
    tflops_peak, hbm_peak, peak_src = _peaks()
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)

    def timed(fn, steps):
        out = []
        for _ in range(steps):
            flush.fill_(1)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            fn()
            b.record()
            torch.cuda.synchronize()
            out.append(a.elapsed_time(b))
        return float(np.sum(out))

    wl = args.workload
    extra = {}
    if wl == "c3":
        from simple_tip_b200.core.surprise import LSA

        xtr, _, xte, _, _ = np_oracle.synth_clusters(60000, 10000, 256, 10, seed=3, spread=1.0)
        rb = lambda a: torch.from_numpy(a).to(torch.bfloat16).to(torch.float32).numpy()   # bf16-stored traces
        xtr, xte = rb(xtr), rb(xte)
        sa = LSA(xtr)
        kde = sa.kde
        xd = E.to_device(xte, dev)
        pinned = torch.from_numpy(xte).pin_memory().numpy()
        n_units, metric, dtype = 10000, "lsa_inputs_prioritized_per_sec", "bf16x3 (split-bf16 tensor-core dot, fp32 log-sum-exp, f64 finish)"
        workload = "C3: LSA Gaussian-KDE 10000 test x 60000 train x 256-d, traces stored in bf16 (seed 3)"
        step_device = lambda: kde._engine.log_kernel_sum(E.whiten(xd, None, kde._mu_dev, kde._w_dev))
        step_e2e = lambda: sa(pinned)
        h2d, d2h = int(xte.nbytes), 2 * 10000 * 8
        flops = 2.0 * 256 * 10000 * 60000

        def roofline(ms):
            return {"kernel": "pair_kernel<MODE_LSE> + whiten/pack/merge (whole device step)", "bound": "tensor",
                    "achieved": flops / (ms * 1e-3) / 1e12, "peak": tflops_peak, "unit": "TFLOP/s",
                    "frac": flops / (ms * 1e-3) / 1e12 / tflops_peak, "traffic": None, "peak_source": peak_src,
                    "note": "algorithmic 2*D flop per pair; 3 bf16 MMAs are executed per algorithmic one"}

        def cpu():
            n_s = 600          # ~10 s of single-threaded CPU work
            kf = np_oracle.KdeFit(xtr.T.astype(np.float64))
            w = np.linalg.cholesky(kf.inv_cov)
            p, q = kf.dataset.T @ w, xte[:n_s].astype(np.float64) @ w
            t0 = time.perf_counter()
            c_oracle.kde_eval(p, q, 1.0 / kf.n, 1.0, threads=1)
            dt = time.perf_counter() - t0
            return {"value": n_s / dt, "unit": "inputs/s", "cores": 1, "kind": "port",
                    "sample": f"{n_s} test inputs vs all 60000 train rows, {dt:.1f} s; literal loop nest of scipy 1.4.1 "
                              "gaussian_kernel_estimate (single-threaded, as in the reference)"}
    elif wl == "c4":
        from simple_tip_b200.core.neuron_coverage import KMNC

        act, mins, maxs = np_oracle.synth_relu(10000, 4096, seed=4)
        km = KMNC([mins], [maxs], 1000)
        pinned = torch.from_numpy(act).pin_memory().numpy()
        km.buckets([act[:8]])
        a_dev = E.to_device(act, dev)
        lo, jp = km._dev_stats
        bucket = torch.empty((10000, 4096), dtype=torch.int16, device=dev)
        score = torch.empty(10000, dtype=torch.int32, device=dev)
        n_units, metric, dtype = 10000, "kmnc_inputs_profiled_per_sec", "f32 compare, i16 bucket ids"
        workload = "C4: KMNC 10000 x 4096 ReLU traces, 1000 sections (seed 4); compact bucket ids + scores"
        step_device = lambda: lib.tip_kmnc(E._p(a_dev), 0, 10000, 4096, E._p(lo), E._p(jp), 0, 1000, E._p(bucket), 3,
                                           E._p(score), E._stream())
        step_e2e = lambda: km.buckets([pinned])
        h2d, d2h = int(act.nbytes), 10000 * 4096 * 2 + 10000 * 4
        nbytes = act.nbytes + 10000 * 4096 * 2 + 2 * 4096 * 4 + 10000 * 4

        def roofline(ms):
            return {"kernel": "kmnc_strip_kernel", "bound": "hbm", "achieved": nbytes / (ms * 1e-3) / 1e9, "peak": hbm_peak,
                    "unit": "GB/s", "frac": nbytes / (ms * 1e-3) / 1e9 / hbm_peak, "traffic": None, "peak_source": peak_src}

        def cpu():
            n_s, k_s = 2000, 50    # the dense N x D x k profile of the reference is 41 GB at k=1000
            t0 = time.perf_counter()
            np_oracle.kmnc_oracle([mins], [maxs], k_s, [act[:n_s]])
            dt = time.perf_counter() - t0
            per_section = dt / (n_s * k_s)
            return {"value": 1.0 / (per_section * 1000), "unit": "inputs/s", "cores": 1, "kind": "port",
                    "sample": f"{n_s} inputs at sections={k_s} ({dt:.2f} s), cost scaled linearly to 1000 sections "
                              "(the reference loops over sections, neuron_coverage.py:90-93)"}
    else:   # c1
        from simple_tip_b200.core.apfd import apfd_from_order
        from simple_tip_b200.core.deepgini import DeepGini

        p, truth = np_oracle.synth_softmax(10000, 10, seed=1)
        pinned = torch.from_numpy(p).pin_memory().numpy()
        p_dev = E.to_device(p, dev)
        pred_d = torch.empty(10000, dtype=torch.int32, device=dev)
        gini_d = torch.empty(10000, dtype=torch.float32, device=dev)
        n_units, metric, dtype = 10000, "deepgini_inputs_prioritized_per_sec", "f32"
        workload = "C1: DeepGini on 10000 x 10 softmax outputs (seed 1) + APFD"
        step_device = lambda: lib.tip_deepgini(E._p(p_dev), 0, 10000, 10, E._p(pred_d), E._p(gini_d), E._stream())
        step_e2e = lambda: DeepGini.calculate(pinned)
        h2d, d2h = int(p.nbytes), 10000 * 8
        nbytes = p.nbytes + 10000 * 8
        pred, gini = DeepGini.calculate(p)
        wp, wg = np_oracle.deepgini_oracle(p)
        fault = wp != truth
        extra["apfd"] = {"ours": float(apfd_from_order(fault, np.argsort(-gini))),
                         "oracle": float(np_oracle.apfd_oracle(fault, np.argsort(-wg))),
                         "scores_bit_identical": bool(np.array_equal(gini, wg) and np.array_equal(pred, wp))}

        def roofline(ms):
            return {"kernel": "gini_small_kernel", "bound": "hbm", "achieved": nbytes / (ms * 1e-3) / 1e9, "peak": hbm_peak,
                    "unit": "GB/s", "frac": nbytes / (ms * 1e-3) / 1e9 / hbm_peak, "traffic": None, "peak_source": peak_src,
                    "note": "480 KB problem: launch latency, not bandwidth"}

        def cpu():
            t0 = time.perf_counter()
            for _ in range(200):
                np_oracle.deepgini_oracle(p)
            dt = (time.perf_counter() - t0) / 200
            return {"value": 10000 / dt, "unit": "inputs/s", "cores": 1, "kind": "port",
                    "sample": "whole 10000 x 10 batch, NumPy expressions of deepgini.py:33-34, mean of 200 runs"}

    for _ in range(max(3, args.warmup)):
        step_device()
        step_e2e()
    torch.cuda.synchronize()
    sampler = ClockSampler(0)
    sampler.start()
    l0 = _lib.launch_count()
    t_dev = timed(step_device, args.steps)
    launches = _lib.launch_count() - l0
    t_e2e = timed(step_e2e, args.steps)
    clocks = sampler.stop()
    line = {"metric": metric, "value": n_units * args.steps / (t_dev * 1e-3), "unit": "inputs/s", "n_gpus": 1,
            "steps": args.steps, "warmup": max(3, args.warmup), "ms_per_step": t_dev / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": dtype, "data": "synthetic",
            "config": {"workload": workload, "l2": "flushed between steps (256 MiB write)",
                       "timing": "per-step CUDA events, summed"},
            "e2e": {"value": n_units * args.steps / (t_e2e * 1e-3), "unit": "inputs/s", "ms_per_step": t_e2e / args.steps,
                    "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
            "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline(t_dev / args.steps)}
    line.update(extra)
    if not args.no_cpu:
        line["cpu_baseline"] = cpu()
    sys.stdout.flush()
    os.dup2(saved_stdout, 1)
    print(json.dumps(line), flush=True)
    os.dup2(2, 1)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg (profiling runs)")
    ap.add_argument("--shard", default="auto", choices=["auto", "test", "train"],
                    help="multi-GPU layout of C2: test = N_test sharded, train replicated (default: the train set is "
                         "30.7 MB); train = N_train sharded with all-reduces of minima / winner rows (what C5 needs)")
    ap.add_argument("--workload", default="c2", choices=["c1", "c2", "c3", "c4"],
                    help="c2 (default) = the configuration the headline metric is quoted on; c1/c3/c4 = the other "
                         "single-GPU BASELINE.json configurations, same JSON shape")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    elif args.workload != "c2":
        assert args.gpus == 1, "secondary workloads are single-GPU"
        run_secondary(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
