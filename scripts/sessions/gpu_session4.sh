#!/bin/bash
# r02 session 4: whole GPU suite (no -x), smoke, config 4 with CSC price-out v3, default bench line with extras, ncu of the CSC
# price-out, launch list of config 4, config-4 long run
set -u
O=gpurun_out/r02s4
mkdir -p $O
( time timeout 1500 python -m pytest tests -q -m gpu --durations=8 ) > $O/tests_gpu.log 2>&1
echo "gpu tests rc=$?" | tee $O/summary.txt
tail -20 $O/tests_gpu.log
( timeout 300 python -c "import __graft_entry__ as g; g.smoke()" ) > $O/smoke.log 2>&1
echo "smoke rc=$?" | tee -a $O/summary.txt
cat $O/smoke.log | tail -4
for f in 1 8; do
  timeout 900 python bench.py --workload netlib_like --rows 100000 --cols 100000 --steps 3000 --warmup 20 --refactor-factor $f --cpu-baseline-seconds 12 > $O/bench_c4_f$f.json 2> $O/bench_c4_f$f.err
  echo "bench c4 factor $f rc=$?" | tee -a $O/summary.txt
  python -c "
import json,sys
d=json.load(open('$O/bench_c4_f$f.json'))
print('c4 factor $f:', d['value'], 'pivots/s', d['ms_per_step'], 'ms; refactors', d['run_detail']['refactors_in_region'], 'share', d['run_detail']['refactor_share_of_wall'], 'price', d['roofline']['avg_launch_ms'], d['roofline']['achieved'], 'parity', d.get('parity',{}).get('first_divergence'), d.get('parity',{}).get('pivots_compared'))"
done
( time timeout 900 python bench.py --steps 20 --warmup 5 ) > $O/bench_default.json 2> $O/bench_default.err
echo "bench default rc=$?" | tee -a $O/summary.txt
head -c 6000 $O/bench_default.json; tail -3 $O/bench_default.err
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_price_csc_seg --launch-skip 40 -c 1 -o $O/price_csc_v3 -f \
  python bench.py --workload netlib_like --rows 100000 --cols 100000 --steps 60 --warmup 10 --cpu-baseline-seconds 0 > $O/ncu_price_csc.log 2>&1
echo "ncu csc rc=$?" | tee -a $O/summary.txt
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 2500 -c 1500 --csv --log-file $O/launches_c4.csv \
  python bench.py --workload netlib_like --rows 100000 --cols 100000 --steps 150 --warmup 10 --cpu-baseline-seconds 0 > $O/ncu_c4.log 2>&1
echo "ncu launch list c4 rc=$?" | tee -a $O/summary.txt
timeout 300 python scripts/deep_curve.py --workload netlib_like --m 100000 --n 100000 --refactor-factor 8 --segment 2000 --max-pivots 400000 --max-seconds 90 > $O/deep_c4_f8.jsonl 2> $O/deep_c4_f8.err
echo "deep c4 rc=$?" | tee -a $O/summary.txt
cut -c1-420 $O/deep_c4_f8.jsonl | tail -12
cat $O/summary.txt
