#!/bin/bash
# resident filter with the A operand in tensor memory: parity, then A/B (bring-up)
set +e
echo "== probe + dsa tests with A in TMEM =="
B200TIP_A_TMEM=1 timeout 600 python -m pytest tests/test_gpu_parity.py -q -k "probe or dsa" 2>&1 | tail -15 | cut -c1-300
for v in 0 1 0 1; do
  B200TIP_A_TMEM=$v timeout 300 python bench.py --no-c5 --no-others --no-cpu --steps 20 2>/dev/null | python -c "
import json,sys
j=json.loads([l for l in sys.stdin if l.startswith('{')][-1])
r=j['roofline']
print('A_TMEM $v C2 ms', round(j['ms_per_step'],4), 'stage2', round(r['ms_per_launch'],4), 'stage1', round(r['other_launches_ms']['nn_filter_same_class'],4), 'frac', round(r['frac'],3), 'e2e', round(j['e2e']['ms_per_step'],4))"
done
