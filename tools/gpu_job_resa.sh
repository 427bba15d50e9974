#!/bin/bash
set +e
B200TIP_LSE_POLY=1 timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_multi.py -q -k 'lsa or mdsa or mlsa or probe' 2>&1 | tail -2
for v in 0 1 0 1; do
B200TIP_LSE_POLY=$v timeout 300 python bench.py --workload c3 --steps 20 --no-cpu 2>/dev/null | python -c "
import json,sys
j=json.loads([l for l in sys.stdin if l.startswith('{')][-1])
print('POLY $v C3 ms', round(j['ms_per_step'],4), 'frac', round(j['roofline']['frac'],3), 'graph', round(j.get('device_call_graph_ms') or 0,4), 'e2e', round(j['e2e']['ms_per_step_median'],4), 'pc', round(j['pc_lsa']['ms_per_step_e2e_median'],3), 'err', j['parity']['max_rel_err'], j['operand_scheme']['check'], 'retries', j.get('stall_retries_so_far'))"
done
