#!/bin/bash
# r02 session 6: suite (compact core rows in the sparse FTRAN, selection near-ties); config 4 with both cache policies of the
# CSC price-out; config-4 long run towards the optimum
set -u
O=gpurun_out/r02s6
mkdir -p $O
( time timeout 1500 python -m pytest tests -q -m gpu --durations=6 ) > $O/tests_gpu.log 2>&1
echo "gpu tests rc=$?" | tee $O/summary.txt
tail -12 $O/tests_gpu.log
for st in 1 0; do
  MLP_CSC_STREAM=$st timeout 600 python bench.py --workload netlib_like --rows 100000 --cols 100000 --steps 3000 --warmup 20 --cpu-baseline-seconds 10 > $O/c4_stream$st.json 2> $O/c4_stream$st.err
  python -c "
import json; d=json.load(open('$O/c4_stream$st.json')); print('c4 stream=$st', d['value'], d['ms_per_step'], 'refactor share', d['run_detail']['refactor_share_of_wall'], 'price ms', d['roofline']['avg_launch_ms'], d['roofline']['achieved'], 'parity', d['parity']['first_divergence'], d['parity']['pivots_compared'], 'launches/pivot', d['gpu_launches']/d['steps'])"
done
timeout 600 python bench.py --workload netlib_like --rows 100000 --cols 100000 --steps 3000 --warmup 20 --refactor-factor 8 --cpu-baseline-seconds 10 > $O/c4_f8.json 2> $O/c4_f8.err
python -c "
import json; d=json.load(open('$O/c4_f8.json')); print('c4 f8', d['value'], d['ms_per_step'], 'refactor share', d['run_detail']['refactor_share_of_wall'], 'parity', d['parity']['first_divergence'], d['parity']['pivots_compared'])"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 400 -c 900 --csv --log-file $O/launches_c4.csv \
  python bench.py --workload netlib_like --rows 100000 --cols 100000 --steps 45 --warmup 2 --cpu-baseline-seconds 0 > $O/ncu_c4.log 2>&1
echo "ncu launch list c4 rc=$?" | tee -a $O/summary.txt
timeout 420 python scripts/deep_curve.py --workload netlib_like --m 100000 --n 100000 --refactor-factor 8 --segment 4000 --max-pivots 1000000 --max-seconds 300 > $O/deep_c4_f8.jsonl 2> $O/deep_c4_f8.err
echo "deep c4 rc=$?" | tee -a $O/summary.txt
python - <<'PY'
import json
for l in open('gpurun_out/r02s6/deep_c4_f8.jsonl'):
    r=json.loads(l)
    if 'summary' in r: print(r); continue
    print(r['pivots_done'], r['k'], r['K_end'], round(r['ms_per_pivot'],3), r['refactors'], round(r['refactor_wall_ms_per_pivot'],3), round(r['obj'],1), r['primal_infeasible_rows'])
PY
cat $O/summary.txt
