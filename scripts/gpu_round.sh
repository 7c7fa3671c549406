#!/bin/bash
# One GPU-box visit: new tests first, then the whole GPU suite, then the bench A/B, then the ncu launch list.
# Everything is logged under gpurun_out/ (merged back by gpurun).
set -u
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > $O/gpu.txt 2>&1
( time timeout 600 python -m pytest tests/test_fused_gpu.py -x -q -m gpu ) > $O/t_fused.log 2>&1
echo "fused tests rc=$?" | tee -a $O/summary.txt
( time timeout 900 python -m pytest tests -q -m gpu ) > $O/t_all.log 2>&1
echo "all gpu tests rc=$?" | tee -a $O/summary.txt
tail -5 $O/t_all.log
( timeout 300 python -c "import __graft_entry__ as g; g.smoke()" ) > $O/smoke.log 2>&1
echo "smoke rc=$?" | tee -a $O/summary.txt
timeout 400 python bench.py --steps 200 --warmup 5 > $O/bench_new.json 2> $O/bench_new.err
echo "bench new rc=$?" | tee -a $O/summary.txt
MLP_FUSED=0 timeout 400 python bench.py --steps 200 --warmup 5 --cpu-baseline-seconds 0 > $O/bench_nofuse.json 2> $O/bench_nofuse.err
MLP_FUSED=0 MLP_LANE1_LDG=0 timeout 400 python bench.py --steps 200 --warmup 5 --cpu-baseline-seconds 0 > $O/bench_old.json 2> $O/bench_old.err
MLP_LANE1_LDG=0 timeout 400 python bench.py --steps 200 --warmup 5 --cpu-baseline-seconds 0 > $O/bench_fused_tma1.json 2> $O/bench_fused_tma1.err
for f in new nofuse old fused_tma1; do python - <<PY
import json
try:
    d = json.load(open("$O/bench_$f.json"))
    print("$f", round(d["value"], 2), "piv/s", round(d["ms_per_step"], 4), "ms  e2e", round(d["e2e"]["value"], 2), "frac", round(d["roofline"]["frac"], 4),
          "price_v ms", round(d["roofline"]["avg_launch_ms"], 4), "launches", d["gpu_launches"])
except Exception as e:
    print("$f failed", e)
PY
done | tee -a $O/summary.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file $O/launches_r01d.csv python bench.py --steps 12 --warmup 3 --cpu-baseline-seconds 0 > $O/ncu_bench.log 2>&1
echo "ncu rc=$?" | tee -a $O/summary.txt
timeout 300 python scripts/config2_kernels.py > $O/config2.json 2> $O/config2.err
MLP_FUSED=0 MLP_LANE1_LDG=0 timeout 300 python scripts/config2_kernels.py > $O/config2_old.json 2> $O/config2_old.err
cat $O/summary.txt
