"""BASELINE config 5 at FULL size on N GPUs of one box: DSA, 100k test x 1.28M train x 2048-d, 1000 classes,
N_train sharded over the ranks (torchrun).  Same synthetic traces, engine path and oracle check as bench.py's
`n_train_sharded` block (which runs the 10k-test slice of this configuration at every --gpus N).
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 tools/c5_multi.py [--tests 100000]
"""
import argparse
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import bench  # noqa: E402
from simple_tip_b200 import engine as E  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--tests", type=int, default=100000)
ap.add_argument("--train", type=int, default=1280000)
ap.add_argument("--steps", type=int, default=5)
ap.add_argument("--parity-inputs", type=int, default=32)
args = ap.parse_args()
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
bench.pin_to_gpu_numa_node(local)
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
comm = None
d = None
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
    comm = E.TrainShardComm()
    d = dist
tm = bench.Timer(dev, d)
cfg = dict(bench.C5S)
cfg["n_test"], cfg["n_train"] = args.tests, args.train
block = bench.run_c5_slice(args, tm, dev, rank, world, comm, steps=args.steps, cfg=cfg)
if rank == 0:
    block["workload"] = block["workload"].replace("C5 slice", "C5 (full)" if args.tests == 100000 and args.train == 1280000 else "C5 slice")
    print(json.dumps(block))
if world > 1:
    dist.barrier()
    comm.close()
    dist.destroy_process_group()
