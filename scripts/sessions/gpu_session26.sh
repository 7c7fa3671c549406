#!/bin/bash
# r02 session 26: the first half of a dual iteration with one host round trip (mlp_dual_select_ratio): parity suites, then A/B
set -u
O=gpurun_out/r02s26
mkdir -p $O
( time timeout 300 python -m pytest tests/test_parity_gpu.py tests/test_sparse_gpu.py tests/test_sharded_gpu.py tests/test_incremental_gpu.py tests/test_refresh_gpu.py tests/test_deep_gpu.py -q -m gpu -x ) > $O/tests.log 2>&1
echo "tests rc=$?" | tee $O/summary.txt
tail -4 $O/tests.log
for fd in 1 0; do
  MLP_FUSED_DUAL=$fd timeout 200 python bench.py --workload netlib_like --rows 100000 --cols 100000 --steps 3000 --warmup 20 --cpu-baseline-seconds $([ $fd = 1 ] && echo 8 || echo 0) > $O/c4_fd$fd.json 2> $O/c4_fd$fd.err
  python -c "
import json; d=json.load(open('$O/c4_fd$fd.json')); r=d['run_detail']; print('c4 fused_dual=$fd', round(d['value'],1), round(d['ms_per_step'],4), 'refac_wall', round(r['refactor_wall_s'],3), 'parity', (d.get('parity') or {}).get('first_divergence'), (d.get('parity') or {}).get('pivots_compared'))"
  MLP_FUSED_DUAL=$fd timeout 100 python bench.py --kind 1 --steps 400 --warmup 5 --cpu-baseline-seconds $([ $fd = 1 ] && echo 6 || echo 0) --no-extras > $O/k1_fd$fd.json 2> $O/k1_fd$fd.err
  python -c "
import json; d=json.load(open('$O/k1_fd$fd.json')); print('kind1 50k fused_dual=$fd', round(d['value'],1), round(d['ms_per_step'],4), 'parity', (d.get('parity') or {}).get('first_divergence'), (d.get('parity') or {}).get('pivots_compared'))"
done
