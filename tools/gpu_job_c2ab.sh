#!/bin/bash
# C2 chain A/B (bring-up): shared-memory carve-out of the small kernels, per-CTA clocks of both filter launches
set +e
O=gpurun_out; mkdir -p $O
for c in 0 1; do
  echo "== carveout $c: cta clock =="
  B200TIP_CARVEOUT=$c timeout 300 python tools/cta_clock.py 2>&1 | grep -v Warning | cut -c1-700
  for rep in 1 2; do
  B200TIP_CARVEOUT=$c timeout 300 python bench.py --no-c5 --no-others --no-cpu --steps 20 2>/dev/null | python -c "
import json,sys
j=json.loads([l for l in sys.stdin if l.startswith('{')][-1])
r=j['roofline']
print('carveout $c C2 ms', round(j['ms_per_step'],4), 'stage2', round(r['ms_per_launch'],4), 'stage1', r['other_launches_ms'], 'e2e', round(j['e2e']['ms_per_step'],4))"
  done
done
