#!/bin/bash
set -u
O=gpurun_out/r02s16
mkdir -p $O
( time timeout 900 python -m pytest tests/test_refresh_gpu.py tests/test_fullsize_gpu.py -q -m gpu --durations=4 ) > $O/tests.log 2>&1
echo "tests rc=$?" | tee $O/summary.txt
tail -12 $O/tests.log
MLP_REFACTOR_TRACE=1 timeout 600 python bench.py --workload netlib_like --rows 100000 --cols 100000 --steps 3000 --warmup 20 --cpu-baseline-seconds 0 > $O/bench_c4_trace.json 2> $O/bench_c4_trace.err
grep "refactor trace" $O/bench_c4_trace.err
for rep in 1 2; do
  timeout 600 python bench.py --workload netlib_like --rows 100000 --cols 100000 --steps 3000 --warmup 20 --cpu-baseline-seconds $([ $rep = 1 ] && echo 15 || echo 0) > $O/bench_c4_$rep.json 2> $O/bench_c4_$rep.err
  python -c "
import json; d=json.load(open('$O/bench_c4_$rep.json')); r=d['run_detail']; print('c4 default rep $rep', round(d['value'],1), round(d['ms_per_step'],4), 'refactors', r['refactors_in_region'], 'refac_wall', round(r['refactor_wall_s'],3), 'wall', round(d['e2e']['wall_s'],3), 'parity', (d.get('parity') or {}).get('first_divergence'), (d.get('parity') or {}).get('pivots_compared'))"
done
MLP_REFACTOR_TRACE=2 timeout 300 python scripts/deep_curve.py --workload netlib_like --m 100000 --n 100000 --segment 4000 --max-pivots 400000 --max-seconds 40 > $O/deep_c4.jsonl 2> $O/deep_c4.err
echo "deep c4 rc=$?" | tee -a $O/summary.txt
python - <<'PY'
import json
for l in open('gpurun_out/r02s16/deep_c4.jsonl'):
    d=json.loads(l)
    if d.get('summary'): print({k:d[k] for k in ('pivots','device_seconds','objective')})
    else: print(d['pivots_done'], 'k', d['k'], 'K', d['K_end'], 'ms/pivot', round(d['ms_per_pivot'],3), 'refactors', d['refactors'], 'refac ms/pivot', round(d['refactor_wall_ms_per_pivot'],3), 'infeasible rows', d.get('primal_infeasible_rows'))
PY
grep "refactor trace\] [0-9a]" $O/deep_c4.err; grep "refactor event" $O/deep_c4.err | tail -15
cat $O/summary.txt
