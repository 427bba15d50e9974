#!/bin/bash
# One-GPU validation + measurement job (run on the GPU box through gpurun; writes everything under gpurun_out/).
set +e
O=gpurun_out
mkdir -p $O
echo "== pytest (all gpu tests) =="; timeout 900 python -m pytest tests -m gpu -q -x > $O/r2_pytest_all.log 2>&1; tail -6 $O/r2_pytest_all.log
echo "== bench default =="; timeout 600 python bench.py > $O/r2_bench_full.json 2> $O/r2_bench_full.err; echo rc=$?; tail -2 $O/r2_bench_full.err | cut -c1-300
echo "== pc-lsa profile =="; timeout 300 python tools/pc_lsa_profile.py > $O/r2_pclsa.txt 2>&1; head -12 $O/r2_pclsa.txt | cut -c1-200
echo "== A/B: seeds / group re-rank =="
for cfg in "B200TIP_SEEDS=0 B200TIP_RERANK_GROUPS=0" "B200TIP_SEEDS=1 B200TIP_RERANK_GROUPS=0"; do
  env $cfg timeout 200 python bench.py --no-c5 --no-others --no-cpu --steps 20 > $O/ab.json 2> $O/ab.err
  python - "$cfg" <<'PY'
import json,sys
try:
    d=json.load(open("gpurun_out/ab.json")); r=d["roofline"]
    print(sys.argv[1], "ms/step", round(d["ms_per_step"],4), "e2e", round(d["e2e"]["ms_per_step"],4), "stage2", round(r["ms_per_launch"],4), "stage1", r["other_launches_ms"])
except Exception as e:
    print(sys.argv[1], "FAILED", e)
PY
done
echo "== pair2 (cta_group::2) =="
B200TIP_PAIR2=1 timeout 300 python -m pytest tests/test_gpu_parity.py -q -x -k "pair_probe or config5 or lsa_config3 or lsa_random or lsa_golden" > $O/r2_pytest_pair2.log 2>&1; tail -4 $O/r2_pytest_pair2.log
for p in 0 1; do
  B200TIP_PAIR2=$p timeout 300 python bench.py --workload c5s --steps 5 > $O/r2_c5s_pair$p.json 2> $O/r2_c5s_pair$p.err
  B200TIP_PAIR2=$p timeout 300 python bench.py --workload c3 --steps 10 --no-cpu > $O/r2_c3_pair$p.json 2> $O/r2_c3_pair$p.err
  python - $p <<'PY'
import json,sys
p=sys.argv[1]
for f,k in ((f"gpurun_out/r2_c5s_pair{p}.json","ms_per_pass"),(f"gpurun_out/r2_c3_pair{p}.json","ms_per_step")):
    try:
        d=json.load(open(f)); print("pair2="+p, f, k, d[k], d.get("frac_of_n_x_tensor_peak", d.get("roofline",{}).get("frac")), d.get("parity_ok", d.get("parity")))
    except Exception as e:
        print("pair2="+p, f, "FAILED", e)
PY
done
echo "== ncu =="
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r2_launches_bench_c2.csv python bench.py --no-c5 --no-others --no-cpu --steps 2 --warmup 1 > $O/r2_bench_under_ncu.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:pair_rs_kernel --launch-skip 5 --launch-count 2 -o $O/r2_prof_rs -f python tools/ncu_target.py > $O/r2_ncu_rs.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:pair_kernel -s 4 -c 2 -o $O/r2_prof_lse -f python tools/ncu_lse.py > $O/r2_ncu_lse.log 2>&1
timeout 300 ncu --set full --clock-control none -k regex:kmnc_ --launch-skip 2 --launch-count 1 -o $O/r2_prof_kmnc -f python tools/ncu_kmnc.py > $O/r2_ncu_kmnc.log 2>&1
timeout 400 ncu --set full --clock-control none -k regex:pair_kernel -s 3 -c 1 -o $O/r2_prof_c5 -f python tools/ncu_c5.py > $O/r2_ncu_c5.log 2>&1
timeout 300 ncu --set full --clock-control none -k regex:rerank_list -s 5 -c 2 -o $O/r2_prof_rerank -f python tools/ncu_target.py > $O/r2_ncu_rerank.log 2>&1
B200TIP_PAIR2=1 timeout 400 ncu --set full --clock-control none -k regex:pair2_kernel -s 3 -c 1 -o $O/r2_prof_c5_pair2 -f python tools/ncu_c5.py > $O/r2_ncu_c5_pair2.log 2>&1
ls -la $O/*.ncu-rep | tail -8
echo done
