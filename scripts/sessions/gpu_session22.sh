#!/bin/bash
set -u
O=gpurun_out/r02s22
mkdir -p $O
for rep in 1 2 3; do
  MLP_REFACTOR_TRACE=2 timeout 600 python bench.py --workload netlib_like --rows 100000 --cols 100000 --steps 3000 --warmup 20 --cpu-baseline-seconds 0 > $O/b_$rep.json 2> $O/b_$rep.err
  python -c "
import json; d=json.load(open('$O/b_$rep.json')); r=d['run_detail']; print('rep $rep', round(d['value'],1), round(d['ms_per_step'],4), 'refac_wall', round(r['refactor_wall_s'],3), 'wall', round(d['e2e']['wall_s'],3))"
  grep "refactor event" $O/b_$rep.err | awk '{ if ($(NF-3)+0 > 2.0 || 1) print }' | sort -t: -k2 -n | awk '{for(i=1;i<=NF;i++) if ($i=="ms") {v=$(i-1)}; if (v+0 > 3.0) print}' | head -20
done
