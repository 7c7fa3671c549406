#!/bin/bash
set -u
O=gpurun_out/r02s17
mkdir -p $O
MLP_REFACTOR_TRACE=2 timeout 600 python bench.py --workload netlib_like --rows 100000 --cols 100000 --steps 3000 --warmup 20 --cpu-baseline-seconds 0 > $O/c4.json 2> $O/c4.err
python -c "
import json; d=json.load(open('$O/c4.json')); r=d['run_detail']; print('c4', round(d['value'],1), round(d['ms_per_step'],4), 'refac_wall', round(r['refactor_wall_s'],3), 'wall', round(d['e2e']['wall_s'],3))"
grep "refactor trace\] [0-9ah]" $O/c4.err
MLP_REFACTOR_TRACE=2 timeout 300 python bench.py --kind 1 --steps 400 --warmup 5 --cpu-baseline-seconds 0 --no-extras > $O/k1.json 2> $O/k1.err
python -c "
import json; d=json.load(open('$O/k1.json')); print('kind1 50k', round(d['value'],1), round(d['ms_per_step'],4), 'launches/pivot', d['gpu_launches']/d['steps'])"
grep "refactor trace\] [0-9ah]" $O/k1.err
MLP_REFACTOR_TRACE=2 timeout 300 python bench.py --rows 1000 --cols 1000 --steps 120 --warmup 5 --cpu-baseline-seconds 0 --no-extras > $O/c2.json 2> $O/c2.err
python -c "
import json; d=json.load(open('$O/c2.json')); print('config2 1000x1000', round(d['value'],1), round(d['ms_per_step'],4), 'launches/pivot', d['gpu_launches']/d['steps'])"
grep "refactor trace\] [0-9ah]" $O/c2.err
