#!/bin/bash
# r02 session 8: where a refactorization of config 4 spends its time (stage trace with syncs + ncu launch list deep in the run)
set -u
O=gpurun_out/r02s8
mkdir -p $O
MLP_REFACTOR_TRACE=1 timeout 600 python bench.py --workload netlib_like --rows 100000 --cols 100000 --steps 3000 --warmup 20 --cpu-baseline-seconds 0 > $O/bench_c4_trace.json 2> $O/bench_c4_trace.err
echo "trace rc=$?" | tee $O/summary.txt
grep "refactor trace" $O/bench_c4_trace.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 60000 -c 4000 --csv --log-file $O/launches_c4_deep.csv \
  python bench.py --workload netlib_like --rows 100000 --cols 100000 --steps 2500 --warmup 10 --cpu-baseline-seconds 0 > $O/ncu_c4.log 2>&1
echo "ncu launch list c4 rc=$?" | tee -a $O/summary.txt
python scripts/summarize_ncu.py launches $O/launches_c4_deep.csv $O/launches_c4_deep_summary.md
head -50 $O/launches_c4_deep_summary.md
cat $O/summary.txt
