#!/bin/bash
# Multi-GPU job: N = number of GPUs of the box (gpurun --gpus N).  Writes under gpurun_out/.
set +e
N=${1:-2}
O=gpurun_out
mkdir -p $O
if [ "$N" -le 4 ]; then
  echo "== pytest multi =="; timeout 600 python -m pytest tests/test_gpu_multi.py -q -x > $O/r2_pytest_multi_$N.log 2>&1; tail -4 $O/r2_pytest_multi_$N.log
fi
echo "== bench --gpus $N =="
NCCL_DEBUG=INFO timeout 700 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus $N --steps 10 --warmup 3 > $O/r2_bench_${N}gpu.json 2> $O/r2_bench_${N}gpu.err
echo rc=$?; grep -c "comm 0x" $O/r2_bench_${N}gpu.err; grep "^\[bench\]" $O/r2_bench_${N}gpu.err | cut -c1-1800
if [ "$N" -ge 8 ]; then
  echo "== full C5 =="
  timeout 700 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 tools/c5_multi.py --steps 3 > $O/r2_c5_full_${N}gpu.json 2> $O/r2_c5_full_${N}gpu.err
  echo rc=$?; tail -2 $O/r2_c5_full_${N}gpu.err | cut -c1-300; cat $O/r2_c5_full_${N}gpu.json | cut -c1-2500
fi
echo done
