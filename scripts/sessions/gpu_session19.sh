#!/bin/bash
set -u
O=gpurun_out/r02s19
mkdir -p $O
( time timeout 900 python -m pytest tests/test_refresh_gpu.py tests/test_sparse_gpu.py tests/test_incremental_gpu.py tests/test_fullsize_gpu.py tests/test_sharded_gpu.py tests/test_engine_abi_gpu.py tests/test_recalc_gpu.py -q -m gpu -x --durations=4 ) > $O/tests.log 2>&1
echo "tests rc=$?" | tee $O/summary.txt
tail -12 $O/tests.log
for rep in 1 2; do
  timeout 600 python bench.py --workload netlib_like --rows 100000 --cols 100000 --steps 3000 --warmup 20 --cpu-baseline-seconds $([ $rep = 1 ] && echo 15 || echo 0) > $O/bench_c4_$rep.json 2> $O/bench_c4_$rep.err
  python -c "
import json; d=json.load(open('$O/bench_c4_$rep.json')); r=d['run_detail']; print('c4 default rep $rep', round(d['value'],1), round(d['ms_per_step'],4), 'refactors', r['refactors_in_region'], 'refreshes', r['of_them_product_form_refreshes'], 'refac_wall', round(r['refactor_wall_s'],3), 'wall', round(d['e2e']['wall_s'],3), 'parity', (d.get('parity') or {}).get('first_divergence'), (d.get('parity') or {}).get('pivots_compared'))"
  tail -2 $O/bench_c4_$rep.err
done
MLP_REFACTOR_TRACE=2 timeout 600 python bench.py --workload netlib_like --rows 100000 --cols 100000 --steps 3000 --warmup 20 --cpu-baseline-seconds 0 > $O/bench_c4_trace.json 2> $O/bench_c4_trace.err
grep "refactor trace" $O/bench_c4_trace.err
cat $O/summary.txt
