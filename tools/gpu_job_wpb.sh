#!/bin/bash
# exact re-rank: queries per block A/B (bring-up)
set +e
for w in 8 2 1; do
  echo "== WPB $w =="
  B200TIP_RERANK_WPB=$w timeout 300 python -m pytest tests/test_gpu_parity.py -q -x -k "dsa" 2>&1 | tail -1
  for rep in 1 2; do
  B200TIP_RERANK_WPB=$w timeout 300 python bench.py --no-c5 --no-others --no-cpu --steps 20 2>/dev/null | python -c "
import json,sys
j=json.loads([l for l in sys.stdin if l.startswith('{')][-1])
r=j['roofline']
print('WPB $w C2 ms', round(j['ms_per_step'],4), 'stage2', round(r['ms_per_launch'],4), 'e2e', round(j['e2e']['ms_per_step'],4))"
  done
  B200TIP_RERANK_WPB=$w timeout 300 python bench.py --workload c5s --steps 5 --no-cpu 2>/dev/null | python -c "
import json,sys
j=json.loads([l for l in sys.stdin if l.startswith('{')][-1])
print('WPB $w C5s ms', j.get('ms_per_pass'), j.get('parity_ok'))"
done
