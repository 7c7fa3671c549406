#!/bin/bash
set -u
O=gpurun_out/r02s13
mkdir -p $O
for rep in 1 2; do
for ms in 0 100 1000; do
for le in 0 100000000; do
  MLP_BENCH_CLOCK_SAMPLE_MS=$ms MLP_LU_EVERY=$le timeout 600 python bench.py --workload netlib_like --rows 100000 --cols 100000 --steps 3000 --warmup 20 --cpu-baseline-seconds 0 > $O/b_${ms}_${le}_$rep.json 2> $O/b_${ms}_${le}_$rep.err
  python -c "
import json; d=json.load(open('$O/b_${ms}_${le}_$rep.json')); r=d['run_detail']; print('sampler $ms ms lu_every $le rep $rep', round(d['value'],1), round(d['ms_per_step'],4), 'refac_wall', round(r['refactor_wall_s'],3), 'wall', round(d['e2e']['wall_s'],3))"
done
done
done
