#!/bin/bash
set +e
O=gpurun_out; mkdir -p $O
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/r2_launches_bench_c3.csv python bench.py --workload c3 --steps 3 --warmup 1 --no-cpu > $O/r2_c3_under_ncu.log 2>&1
python tools/launch_list.py $O/r2_launches_bench_c3.csv | head -14 | cut -c1-160
python - <<'PY'
import csv
rows=list(csv.reader(open('gpurun_out/r2_launches_bench_c3.csv')))
hdr=[i for i,r in enumerate(rows) if r and r[0]=='ID'][0]
h=rows[hdr]; ki=h.index('Kernel Name'); vi=h.index('Metric Value')
seq=[(r[ki].split('(')[0][-50:], float(r[vi].replace(',',''))/1e3) for r in rows[hdr+1:] if len(r)>vi and r[vi]]
idx=[i for i,(k,v) in enumerate(seq) if 'pair_kernel' in k]
i=idx[-1]
for k,v in seq[i-6:i+4]: print(f"{v:9.1f}  {k}")
PY
