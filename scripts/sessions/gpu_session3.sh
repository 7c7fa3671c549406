#!/bin/bash
# r02 session 3: full GPU suite; config 4 with the new CSC price-out at refactor factors 1 / 8; fused chain limit A/B in the deep
# regime of config 3; launch list of config 4
set -u
O=gpurun_out/r02s3
mkdir -p $O
( time timeout 1500 python -m pytest tests -q -m gpu -x --durations=8 ) > $O/tests_gpu.log 2>&1
echo "gpu tests rc=$?" | tee $O/summary.txt
tail -14 $O/tests_gpu.log
for f in 1 8; do
  timeout 900 python bench.py --workload netlib_like --rows 100000 --cols 100000 --steps 3000 --warmup 20 --refactor-factor $f --cpu-baseline-seconds 12 > $O/bench_c4_f$f.json 2> $O/bench_c4_f$f.err
  echo "bench c4 factor $f rc=$?" | tee -a $O/summary.txt
  python -c "
import json,sys
d=json.load(open('$O/bench_c4_f$f.json'))
print('c4 factor $f:', d['value'], 'pivots/s', d['ms_per_step'], 'ms; refactors', d['run_detail']['refactors_in_region'], 'share', d['run_detail']['refactor_share_of_wall'], 'price', d['roofline']['avg_launch_ms'], d['roofline']['achieved'], 'parity', d.get('parity'))"
done
for fm in 512 1024; do
  MLP_FUSED_MAX=$fm timeout 600 python scripts/deep_curve.py --skip 2500 --segment 500 --max-pivots 4500 > $O/deep_fused$fm.jsonl 2> $O/deep_fused$fm.err
  echo "deep fused_max $fm rc=$?" | tee -a $O/summary.txt
  cut -c1-330 $O/deep_fused$fm.jsonl
done
timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 6000 -c 2500 --csv --log-file $O/launches_c4.csv \
  python bench.py --workload netlib_like --rows 100000 --cols 100000 --steps 300 --warmup 20 --cpu-baseline-seconds 0 > $O/ncu_c4.log 2>&1
echo "ncu launch list c4 rc=$?" | tee -a $O/summary.txt
cat $O/summary.txt
