#!/bin/bash
# r02 session 7: verification of the final state — suite, smoke, sanitizer on the small solves, the driver's two bench commands
set -u
O=gpurun_out/r02s7
mkdir -p $O
( time timeout 1500 python -m pytest tests -q -m gpu --durations=6 ) > $O/tests_gpu.log 2>&1
echo "gpu tests rc=$?" | tee $O/summary.txt
tail -12 $O/tests_gpu.log
( timeout 300 python -c "import __graft_entry__ as g; g.smoke()" ) > $O/smoke.log 2>&1
echo "smoke rc=$?" | tee -a $O/summary.txt
tail -3 $O/smoke.log
for tool in memcheck racecheck; do
  timeout 900 compute-sanitizer --tool $tool python scripts/sanitize_small.py > $O/san_$tool.log 2>&1
  echo "sanitizer $tool rc=$? : $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' $O/san_$tool.log | tail -1)" | tee -a $O/summary.txt
done
( time timeout 900 python bench.py --steps 20 --warmup 5 ) > $O/bench_default.json 2> $O/bench_default.err
echo "bench default rc=$?" | tee -a $O/summary.txt
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02s7/bench_default.json'))
print('config3', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], 'frac', d['roofline']['frac'], 'parity', d['parity']['first_divergence'], d['parity']['pivots_compared'], 'cpu', d['cpu_baseline']['value'])
for k,v in d.get('extra',{}).items(): print(k, v.get('value'), v.get('ms_per_step'), v.get('error'), v.get('roofline',{}).get('frac'), (v.get('parity') or {}).get('first_divergence'), (v.get('parity') or {}).get('pivots_compared'))
PY
tail -3 $O/bench_default.err
for f in 1 8; do
timeout 600 python bench.py --workload netlib_like --rows 100000 --cols 100000 --steps 3000 --warmup 20 --refactor-factor $f --cpu-baseline-seconds 15 > $O/bench_c4_f$f.json 2> $O/bench_c4_f$f.err
python -c "
import json; d=json.load(open('$O/bench_c4_f$f.json')); print('c4 f$f', d['value'], d['ms_per_step'], 'refactor share', d['run_detail']['refactor_share_of_wall'], 'price ms', d['roofline']['avg_launch_ms'], d['roofline']['achieved'], 'parity', d['parity']['first_divergence'], d['parity']['pivots_compared'], 'cpu', d['cpu_baseline']['value'], 'launches/pivot', d['gpu_launches']/d['steps'])"
done
for kind in 1 2; do
  timeout 600 python bench.py --kind $kind --steps 200 --warmup 5 --cpu-baseline-seconds 12 --no-extras > $O/bench_kind$kind.json 2> $O/bench_kind$kind.err
  python -c "
import json; d=json.load(open('$O/bench_kind$kind.json')); print('kind $kind 50k', d['value'], d['ms_per_step'], 'parity', d['parity']['first_divergence'], d['parity']['pivots_compared'], 'cpu', d['cpu_baseline']['value'])"
done
timeout 300 python tests/tools/config2_kernels.py > $O/config2.json 2> $O/config2.err
cat $O/config2.json
( time timeout 900 python bench.py --impl reference --steps 20 --warmup 5 ) > $O/bench_reference.json 2> $O/bench_reference.err
echo "bench reference rc=$?" | tee -a $O/summary.txt
cat $O/bench_reference.json | cut -c1-600
cat $O/summary.txt
