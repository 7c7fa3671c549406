#!/bin/bash
# r02 session 10: refresh tests; where the refactor pivots of a refresh-only run spend their time
set -u
O=gpurun_out/r02s10
mkdir -p $O
( time timeout 900 python -m pytest tests/test_refresh_gpu.py -q -m gpu --durations=5 ) > $O/tests_refresh.log 2>&1
echo "refresh tests rc=$?" | tee $O/summary.txt
tail -25 $O/tests_refresh.log
for le in 100000000 128; do
MLP_LU_EVERY=$le MLP_REFACTOR_TRACE=1 timeout 600 python bench.py --workload netlib_like --rows 100000 --cols 100000 --steps 3000 --warmup 20 --cpu-baseline-seconds 0 > $O/bench_c4_trace_le$le.json 2> $O/bench_c4_trace_le$le.err
grep "refactor trace" $O/bench_c4_trace_le$le.err
done
cat $O/summary.txt
