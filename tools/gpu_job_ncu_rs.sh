#!/bin/bash
# ncu --set full of the resident filter (stage 2 and stage 1 launches of C2)
set +e
O=gpurun_out; mkdir -p $O
timeout 600 ncu --set full --clock-control none --import-source on -k regex:pair_rs_kernel --launch-skip 5 --launch-count 2 -o $O/r2_prof_pair_rs -f python tools/ncu_target.py > $O/r2_ncu_pair_rs.log 2>&1
tail -3 $O/r2_ncu_pair_rs.log | cut -c1-200
ls -la $O/r2_prof_pair_rs.ncu-rep
