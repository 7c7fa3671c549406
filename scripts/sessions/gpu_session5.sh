#!/bin/bash
# r02 session 5: suite with PDL launches + blocked inverse; divergence diagnosis of the failing sparse case; PDL A/B in the
# latency-bound regimes; launch list of config 4; config-4 long run with the blocked inverse
set -u
O=gpurun_out/r02s5
mkdir -p $O
( time timeout 1500 python -m pytest tests -q -m gpu --durations=6 ) > $O/tests_gpu.log 2>&1
echo "gpu tests rc=$?" | tee $O/summary.txt
tail -12 $O/tests_gpu.log
timeout 300 python tests/tools/sparse_divergence.py netlib_like 500 350 7.0 5 > $O/divergence.json 2> $O/divergence.err
echo "divergence rc=$?" | tee -a $O/summary.txt
cat $O/divergence.json
for pdl in 0 1; do
  MLP_PDL=$pdl timeout 300 python tests/tools/config2_kernels.py > $O/config2_pdl$pdl.json 2> $O/config2_pdl$pdl.err
  python -c "
import json; d=json.load(open('$O/config2_pdl$pdl.json')); print('config2 pdl=$pdl', d['gpu_us'], d['full_solve'])"
  MLP_PDL=$pdl timeout 600 python bench.py --kind 1 --steps 200 --warmup 5 --cpu-baseline-seconds 0 --no-extras > $O/kind1_pdl$pdl.json 2> $O/kind1_pdl$pdl.err
  python -c "
import json; d=json.load(open('$O/kind1_pdl$pdl.json')); print('kind1 50k pdl=$pdl', d['value'], d['ms_per_step'], d['gpu_launches']/d['steps'])"
  MLP_PDL=$pdl timeout 600 python bench.py --workload netlib_like --rows 100000 --cols 100000 --steps 2000 --warmup 20 --cpu-baseline-seconds 0 > $O/c4_pdl$pdl.json 2> $O/c4_pdl$pdl.err
  python -c "
import json; d=json.load(open('$O/c4_pdl$pdl.json')); print('c4 pdl=$pdl', d['value'], d['ms_per_step'], d['run_detail']['refactor_share_of_wall'], d['roofline']['avg_launch_ms'])"
done
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 400 -c 900 --csv --log-file $O/launches_c4.csv \
  python bench.py --workload netlib_like --rows 100000 --cols 100000 --steps 45 --warmup 2 --cpu-baseline-seconds 0 > $O/ncu_c4.log 2>&1
echo "ncu launch list c4 rc=$?" | tee -a $O/summary.txt
timeout 200 python scripts/deep_curve.py --workload netlib_like --m 100000 --n 100000 --refactor-factor 8 --segment 2000 --max-pivots 400000 --max-seconds 75 > $O/deep_c4_f8.jsonl 2> $O/deep_c4_f8.err
echo "deep c4 rc=$?" | tee -a $O/summary.txt
python - <<'PY'
import json
for l in open('gpurun_out/r02s5/deep_c4_f8.jsonl'):
    r=json.loads(l)
    if 'summary' in r: print(r); continue
    print(r['pivots_done'], r['k'], r['K_end'], round(r['ms_per_pivot'],3), r['refactors'], round(r['refactor_wall_ms_per_pivot'],3), round(r['obj'],1))
PY
cat $O/summary.txt
