#!/bin/bash
set +e
timeout 900 python -m pytest tests/test_gpu_parity.py -q -k 'lsa or mdsa or mlsa' 2>&1 | tail -2
for rep in 1 2; do
timeout 300 python bench.py --workload c3 --steps 20 --no-cpu 2>/dev/null | python -c "
import json,sys
j=json.loads([l for l in sys.stdin if l.startswith('{')][-1])
print('C3 ms', round(j['ms_per_step'],4), 'frac', round(j['roofline']['frac'],3), 'e2e', round(j['e2e']['ms_per_step'],4), round(j['e2e']['ms_per_step_median'],4), round(j['e2e']['ms_per_step_max'],4), 'eager', j.get('ms_per_step_eager_kernels_only'), 'pcmed', j['pc_lsa'].get('ms_per_step_e2e_median'), 'pc_lsa', round(j['pc_lsa']['ms_per_step_e2e'],3), 'parity', j['parity']['max_rel_err'], j['pc_lsa']['max_rel_err_256_rows'])"
done
