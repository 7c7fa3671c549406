#!/bin/bash
# r02 session 2: full GPU suite, config-3 ms/pivot-vs-k curve to the optimum, config-4 bench line + ncu of the CSC price-out,
# dual-loop lines at 50k, config-2 latency regime
set -u
O=gpurun_out/r02s2
mkdir -p $O
( time timeout 1500 python -m pytest tests -q -m gpu -x --durations=8 ) > $O/tests_gpu.log 2>&1
echo "gpu tests rc=$?" | tee $O/summary.txt
tail -14 $O/tests_gpu.log
timeout 900 python bench.py --workload netlib_like --rows 100000 --cols 100000 --steps 3000 --warmup 20 > $O/bench_c4.json 2> $O/bench_c4.err
echo "bench c4 rc=$?" | tee -a $O/summary.txt
cat $O/bench_c4.json
timeout 600 python scripts/deep_curve.py --segment 500 --max-seconds 240 > $O/deep_curve.jsonl 2> $O/deep_curve.err
echo "deep curve rc=$?" | tee -a $O/summary.txt
tail -4 $O/deep_curve.jsonl
for kind in 1 2; do
  timeout 600 python bench.py --kind $kind --steps 100 --warmup 5 --cpu-baseline-seconds 20 > $O/bench_kind$kind.json 2> $O/bench_kind$kind.err
  echo "bench kind $kind rc=$?" | tee -a $O/summary.txt
  cat $O/bench_kind$kind.json
done
timeout 300 python tests/tools/config2_kernels.py > $O/config2.json 2> $O/config2.err
echo "config2 rc=$?" | tee -a $O/summary.txt
cat $O/config2.json
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_price_csc_seg --launch-skip 60 -c 2 -o $O/price_csc_r02 -f \
  python bench.py --workload netlib_like --rows 100000 --cols 100000 --steps 200 --warmup 20 --cpu-baseline-seconds 0 > $O/ncu_price_csc.log 2>&1
echo "ncu csc rc=$?" | tee -a $O/summary.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 30000 -c 1500 --csv --log-file $O/launches_c4.csv \
  python bench.py --workload netlib_like --rows 100000 --cols 100000 --steps 1200 --warmup 20 --cpu-baseline-seconds 0 > $O/ncu_c4.log 2>&1
echo "ncu launch list c4 rc=$?" | tee -a $O/summary.txt
cat $O/summary.txt
