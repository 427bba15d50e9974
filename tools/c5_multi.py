"""BASELINE config 5 on N GPUs of one box: DSA, 100k test x 1.28M train x 2048-d, 1000 classes,
N_train sharded over the ranks (torchrun).  Traces are generated on the device (seeded), the test
batch is identical on every rank, each rank holds 1/N of every class.
    python -m torch.distributed.run --nproc-per-node 8 tools/c5_multi.py [--tests 100000]
"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from simple_tip_b200 import _lib  # noqa: E402
from simple_tip_b200 import engine as E  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--tests", type=int, default=100000)
ap.add_argument("--train", type=int, default=1280000)
ap.add_argument("--steps", type=int, default=3)
args = ap.parse_args()
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if os.environ.get("NCCL_DEBUG", "").upper() in ("VERSION", "INFO"):
    os.environ["NCCL_DEBUG"] = "WARN"
dist.init_process_group("nccl", device_id=dev)
comm = E.TrainShardComm()
classes, d = 1000, 2048
per_class = args.train // classes                 # 1280 rows per class globally
mine = per_class // world                         # rows of every class on this rank
g = torch.Generator(device=dev).manual_seed(5)    # same stream on every rank -> same centres / tests
centres = torch.randn((classes, d), generator=g, device=dev) * 0.5
yte = torch.arange(args.tests, device=dev) % classes
x = centres[yte] + torch.randn((args.tests, d), generator=g, device=dev)
q_order = torch.argsort(yte, stable=True)
x = x[q_order].contiguous()
q_class = yte[q_order].to(torch.int32).contiguous()
q_off = np.arange(classes + 1, dtype=np.int64) * (args.tests // classes)
gs = torch.Generator(device=dev).manual_seed(1000 + rank)   # shard-specific training rows
cls = torch.arange(classes, device=dev).repeat_interleave(mine)
t_sorted = centres[cls] + torch.randn((classes * mine, d), generator=gs, device=dev)
k = torch.arange(mine, device=dev).repeat(classes)
gid = (cls * per_class + rank + world * k).to(torch.int32)         # original index: class-major, dealt round-robin
class_off = np.arange(classes + 1, dtype=np.int64) * mine
eng = E.NnEngine(t_sorted, class_off, gid)
torch.cuda.synchronize()
dist.barrier()
plan = E.dsa_plan(eng, args.tests, q_off, x.dtype, True, comm)
plan.load_sorted(x)
for _ in range(2):
    plan.run()
times = []
for _ in range(args.steps):
    dist.barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    out = plan.run()
    b.record()
    torch.cuda.synchronize()
    times.append(a.elapsed_time(b))
t = torch.tensor([float(np.median(times))], device=dev, dtype=torch.float64)
dist.all_reduce(t, op=dist.ReduceOp.MAX)
# parity on a subset: the sharded exhaustive scan (no tensor-core filter) must give the same bits
rows = torch.from_numpy(np.random.default_rng(1).choice(args.tests, 64, replace=False)).to(dev).sort().values
xs, qs = x[rows].contiguous(), q_class[rows].contiguous()
cnt = np.bincount(qs.cpu().numpy(), minlength=classes)
sub_off = np.concatenate([[0], np.cumsum(cnt)]).astype(np.int64)
ea, eb, eg = E.dsa_distances(eng, xs, qs, sub_off, comm, use_filter=False)
ok = bool(torch.equal(ea, plan.dist_a[rows]) and torch.equal(eb, plan.dist_b[rows]) and torch.equal(eg.long(), plan.gid[rows].long()))
okt = torch.tensor([1 if ok else 0], device=dev)
dist.all_reduce(okt, op=dist.ReduceOp.MIN)
if rank == 0:
    ms = float(t.item())
    flops = 2.0 * d * args.tests * args.train
    print(json.dumps({"config": f"C5: DSA {args.tests} test x {args.train} train x {d}-d, {classes} classes, N_train sharded over {world} GPUs",
                      "n_gpus": world, "ms_per_pass": ms, "inputs_per_s": args.tests / (ms * 1e-3),
                      "algorithmic_tflops_aggregate": flops / (ms * 1e-3) / 1e12,
                      "frac_of_aggregate_bf16_peak": flops / (ms * 1e-3) / 1e12 / (1689.8 * world),
                      "subset_matches_sharded_exhaustive_scan": bool(okt.item()),
                      "stats_exhaustive_rows_candidates_rank0": eng.stats.cpu().numpy().tolist()}))
dist.destroy_process_group()
