#!/bin/bash
# r02 session 20: merged small-K kernels (k_eta_apply, k_unit_eta_t, k_eta_push): full suite, then A/B on config 4 / dual 50k / config 2
set -u
O=gpurun_out/r02s20
mkdir -p $O
( time timeout 1500 python -m pytest tests -q -m gpu --durations=5 ) > $O/tests_gpu.log 2>&1
echo "gpu tests rc=$?" | tee $O/summary.txt
tail -14 $O/tests_gpu.log
for ms in 1 0; do
  MLP_MERGE_SMALL=$ms timeout 600 python bench.py --workload netlib_like --rows 100000 --cols 100000 --steps 3000 --warmup 20 --cpu-baseline-seconds $([ $ms = 1 ] && echo 12 || echo 0) > $O/c4_m$ms.json 2> $O/c4_m$ms.err
  python -c "
import json; d=json.load(open('$O/c4_m$ms.json')); r=d['run_detail']; print('c4 merge=$ms', round(d['value'],1), round(d['ms_per_step'],4), 'refac_wall', round(r['refactor_wall_s'],3), 'launches/pivot', round(d['gpu_launches']/d['steps'],2), 'parity', (d.get('parity') or {}).get('first_divergence'), (d.get('parity') or {}).get('pivots_compared'))"
  MLP_MERGE_SMALL=$ms timeout 300 python bench.py --kind 1 --steps 400 --warmup 5 --cpu-baseline-seconds $([ $ms = 1 ] && echo 10 || echo 0) --no-extras > $O/k1_m$ms.json 2> $O/k1_m$ms.err
  python -c "
import json; d=json.load(open('$O/k1_m$ms.json')); print('kind1 50k merge=$ms', round(d['value'],1), round(d['ms_per_step'],4), 'launches/pivot', d['gpu_launches']/d['steps'], 'parity', (d.get('parity') or {}).get('first_divergence'), (d.get('parity') or {}).get('pivots_compared'))"
  MLP_MERGE_SMALL=$ms timeout 300 python bench.py --rows 1000 --cols 1000 --steps 120 --warmup 5 --cpu-baseline-seconds 0 --no-extras > $O/c2_m$ms.json 2> $O/c2_m$ms.err
  python -c "
import json; d=json.load(open('$O/c2_m$ms.json')); print('config2 merge=$ms', round(d['value'],1), round(d['ms_per_step'],4), 'launches/pivot', d['gpu_launches']/d['steps'])"
done
cat $O/summary.txt
