#!/bin/bash
set -u
O=gpurun_out/r02s12
mkdir -p $O
for rep in 1 2; do
for le in 0 64 128 512 100000000; do
  MLP_LU_EVERY=$le timeout 600 python bench.py --workload netlib_like --rows 100000 --cols 100000 --steps 3000 --warmup 20 --cpu-baseline-seconds 0 > $O/bench_c4_le${le}_$rep.json 2> $O/bench_c4_le${le}_$rep.err
  python -c "
import json; d=json.load(open('$O/bench_c4_le${le}_$rep.json')); r=d['run_detail']; print('c4 lu_every $le rep $rep', round(d['value'],1), round(d['ms_per_step'],4), 'refactors', r['refactors_in_region'], 'refac_wall', round(r['refactor_wall_s'],3), 'wall', round(d['e2e']['wall_s'],3), 'obj', r['objective_after'], 'clocks', d['clocks']['sm_mhz'], d['clocks']['reasons'])"
done
done
