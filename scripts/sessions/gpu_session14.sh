#!/bin/bash
set -u
O=gpurun_out/r02s14
mkdir -p $O
for rep in 1 2 3 4; do
  MLP_LU_EVERY=100000000 MLP_REFACTOR_TRACE=2 timeout 600 python bench.py --workload netlib_like --rows 100000 --cols 100000 --steps 3000 --warmup 20 --cpu-baseline-seconds 0 > $O/b_$rep.json 2> $O/b_$rep.err
  python -c "
import json; d=json.load(open('$O/b_$rep.json')); r=d['run_detail']; print('rep $rep', round(d['value'],1), round(d['ms_per_step'],4), 'refac_wall', round(r['refactor_wall_s'],3), 'wall', round(d['e2e']['wall_s'],3))"
  grep -c "refactor event" $O/b_$rep.err
  grep "refactor event" $O/b_$rep.err | sed -e "s/.*ms in //" | sort | uniq -c | sort -rn | head -8
  grep "refactor trace" $O/b_$rep.err
done
