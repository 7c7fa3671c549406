#!/bin/bash
# r02 session 15: full GPU suite with the refresh as default; config 4 x3 (variance with the pool); deep curve of config 4
set -u
O=gpurun_out/r02s15
mkdir -p $O
( time timeout 1500 python -m pytest tests -q -m gpu --durations=8 ) > $O/tests_gpu.log 2>&1
echo "gpu tests rc=$?" | tee $O/summary.txt
tail -25 $O/tests_gpu.log
for rep in 1 2 3; do
  timeout 600 python bench.py --workload netlib_like --rows 100000 --cols 100000 --steps 3000 --warmup 20 --cpu-baseline-seconds $([ $rep = 1 ] && echo 15 || echo 0) > $O/bench_c4_$rep.json 2> $O/bench_c4_$rep.err
  python -c "
import json; d=json.load(open('$O/bench_c4_$rep.json')); r=d['run_detail']; print('c4 default rep $rep', round(d['value'],1), round(d['ms_per_step'],4), 'refactors', r['refactors_in_region'], 'refac_wall', round(r['refactor_wall_s'],3), 'wall', round(d['e2e']['wall_s'],3), 'parity', (d.get('parity') or {}).get('first_divergence'), (d.get('parity') or {}).get('pivots_compared'))"
done
timeout 300 python scripts/deep_curve.py --workload netlib_like --m 100000 --n 100000 --segment 4000 --max-pivots 400000 --max-seconds 100 > $O/deep_c4.jsonl 2> $O/deep_c4.err
echo "deep c4 rc=$?" | tee -a $O/summary.txt
cut -c1-330 $O/deep_c4.jsonl | tail -30
tail -3 $O/deep_c4.err
cat $O/summary.txt
