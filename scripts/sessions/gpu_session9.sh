#!/bin/bash
# r02 session 9: product-form refresh of the core inverse — tests, then config 4 with lu_every 0 / 128 / 512 / never
set -u
O=gpurun_out/r02s9
mkdir -p $O
( time timeout 900 python -m pytest tests/test_refresh_gpu.py tests/test_sparse_gpu.py -q -m gpu -x --durations=5 ) > $O/tests_refresh.log 2>&1
echo "refresh tests rc=$?" | tee $O/summary.txt
tail -25 $O/tests_refresh.log
for le in 0 128 512 100000000; do
  MLP_LU_EVERY=$le MLP_REFACTOR_TRACE=${TRACE:-0} timeout 600 python bench.py --workload netlib_like --rows 100000 --cols 100000 --steps 3000 --warmup 20 --cpu-baseline-seconds 15 > $O/bench_c4_le$le.json 2> $O/bench_c4_le$le.err
  python -c "
import json; d=json.load(open('$O/bench_c4_le$le.json')); print('c4 lu_every $le', d['value'], d['ms_per_step'], 'refactors', d['run_detail']['refactors_in_region'], 'share', d['run_detail']['refactor_share_of_wall'], 'price ms', d['roofline']['avg_launch_ms'], 'parity', d['parity']['first_divergence'], d['parity']['pivots_compared'], 'obj', d['run_detail']['objective_after'], 'launches/pivot', d['gpu_launches']/d['steps'])"
  tail -2 $O/bench_c4_le$le.err
done
MLP_LU_EVERY=512 MLP_REFACTOR_TRACE=1 timeout 600 python bench.py --workload netlib_like --rows 100000 --cols 100000 --steps 3000 --warmup 20 --cpu-baseline-seconds 0 > $O/bench_c4_trace.json 2> $O/bench_c4_trace.err
grep "refactor trace" $O/bench_c4_trace.err
cat $O/summary.txt
