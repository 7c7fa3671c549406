#!/bin/bash
# r02 session 21: one stream instead of two lanes (MLP_OVERLAP=0) where nothing overlaps anyway: dual loops, sparse, small LPs
set -u
O=gpurun_out/r02s21
mkdir -p $O
for ov in 0 1; do
  MLP_OVERLAP=$ov timeout 600 python bench.py --workload netlib_like --rows 100000 --cols 100000 --steps 3000 --warmup 20 --cpu-baseline-seconds 0 > $O/c4_o$ov.json 2> $O/c4_o$ov.err
  python -c "
import json; d=json.load(open('$O/c4_o$ov.json')); r=d['run_detail']; print('c4 overlap=$ov', round(d['value'],1), round(d['ms_per_step'],4), 'refac_wall', round(r['refactor_wall_s'],3), 'wall', round(d['e2e']['wall_s'],3), 'launches/pivot', round(d['gpu_launches']/d['steps'],2))"
  for kind in 1 2; do
  MLP_OVERLAP=$ov timeout 300 python bench.py --kind $kind --steps 400 --warmup 5 --cpu-baseline-seconds 0 --no-extras > $O/k${kind}_o$ov.json 2> $O/k${kind}_o$ov.err
  python -c "
import json; d=json.load(open('$O/k${kind}_o$ov.json')); print('kind$kind 50k overlap=$ov', round(d['value'],1), round(d['ms_per_step'],4), 'launches/pivot', d['gpu_launches']/d['steps'])"
  done
  MLP_OVERLAP=$ov timeout 300 python bench.py --rows 1000 --cols 1000 --steps 120 --warmup 5 --cpu-baseline-seconds 0 --no-extras > $O/c2_o$ov.json 2> $O/c2_o$ov.err
  python -c "
import json; d=json.load(open('$O/c2_o$ov.json')); print('config2 overlap=$ov', round(d['value'],1), round(d['ms_per_step'],4), 'launches/pivot', d['gpu_launches']/d['steps'])"
done
