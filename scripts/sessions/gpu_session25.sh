#!/bin/bash
set -u
O=gpurun_out/r02s25
mkdir -p $O
( time timeout 600 python -m pytest tests/test_sparse_gpu.py tests/test_refresh_gpu.py tests/test_incremental_gpu.py "tests/test_fullsize_gpu.py::test_config4_netlib_like_100k_follows_the_oracle" "tests/test_fullsize_gpu.py::test_config4_family_to_the_optimum_follows_the_oracle[None]" -q -m gpu -x ) > $O/tests.log 2>&1
echo "tests rc=$?" | tee $O/summary.txt
tail -4 $O/tests.log
for rep in 1 2; do
  timeout 300 python bench.py --workload netlib_like --rows 100000 --cols 100000 --steps 3000 --warmup 20 --cpu-baseline-seconds 0 > $O/bench_c4_$rep.json 2> $O/bench_c4_$rep.err
  python -c "
import json; d=json.load(open('$O/bench_c4_$rep.json')); r=d['run_detail']; print('c4 rep $rep', round(d['value'],1), round(d['ms_per_step'],4), 'refreshes', r['of_them_product_form_refreshes'], 'refac_wall', round(r['refactor_wall_s'],3), 'wall', round(d['e2e']['wall_s'],3), 'setup', r['setup'])"
done
