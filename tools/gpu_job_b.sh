#!/bin/bash
set +e
O=gpurun_out
mkdir -p $O
echo "== pytest (all gpu tests, no -x) =="; timeout 900 python -m pytest tests -m gpu -q > $O/r2_pytest_all.log 2>&1; tail -8 $O/r2_pytest_all.log
echo "== smoke =="; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
echo "== bench default =="; timeout 600 python bench.py > $O/r2_bench_full.json 2> $O/r2_bench_full.err; echo rc=$?; tail -2 $O/r2_bench_full.err | cut -c1-300
echo "== reference arm =="; timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > $O/r2_bench_reference.json 2>&1; tail -c 600 $O/r2_bench_reference.json
echo "== launch list =="
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r2_launches_bench_c2.csv python bench.py --no-c5 --no-others --no-cpu --steps 3 --warmup 1 > $O/r2_bench_under_ncu.log 2>&1
timeout 300 ncu --set full --clock-control none -k regex:kmnc_ --launch-skip 2 --launch-count 1 -o $O/r2_prof_kmnc -f python tools/ncu_kmnc.py > $O/r2_ncu_kmnc.log 2>&1
echo done
