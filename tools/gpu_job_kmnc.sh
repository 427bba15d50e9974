#!/bin/bash
# KMNC grid-shape A/B (bring-up): resident waves 1 / 2 / 4
set +e
O=gpurun_out; mkdir -p $O
timeout 300 python -m pytest tests/test_gpu_parity.py -q -k kmnc 2>&1 | tail -2
for w in 1 2 4; do
  echo "== waves $w =="
  B200TIP_KMNC_WAVES=$w timeout 300 python bench.py --workload c4 --steps 20 --no-cpu 2>/dev/null | python -c "
import json,sys
j=json.loads([l for l in sys.stdin if l.startswith('{')][-1])
print('ms', j['ms_per_step'], 'frac', j['roofline']['frac'], 'after_flush', j.get('ms_per_step_after_flush_write'), 'e2e', j['e2e']['ms_per_step'])"
done
echo "== table test + c2 line =="
timeout 300 python -m pytest tests/test_gpu_parity.py -q -k "fit_time_table or forward_hook" 2>&1 | tail -2
timeout 400 python bench.py --no-c5 --no-cpu --steps 20 2>/dev/null | python -c "
import json,sys
j=json.loads([l for l in sys.stdin if l.startswith('{')][-1])
print('C2', j['ms_per_step'], 'e2e', j['e2e'])
print('table', json.dumps(j.get('fit_time_table')))
for k in ('c3','c4'):
    print(k, j['other_configs'][k]['ms_per_step'], json.dumps(j['other_configs'][k]['e2e']))"
