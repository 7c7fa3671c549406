#!/bin/bash
# 1-GPU box: whole GPU suite, smoke, barrier micro-benchmark, sanitizer, final bench + launch list, config 2 / config 4 footnotes
set -u
O=gpurun_out
mkdir -p $O
( time timeout 900 python -m pytest tests -q -m gpu ) > $O/t6.log 2>&1
echo "all gpu tests rc=$?" | tee $O/summary6.txt
tail -3 $O/t6.log
( timeout 300 python -c "import __graft_entry__ as g; g.smoke()" ) > $O/smoke6.log 2>&1
echo "smoke rc=$?" | tee -a $O/summary6.txt
timeout 120 scripts/micro/barrier_bench > $O/barrier.jsonl 2>&1
cat $O/barrier.jsonl
for tool in memcheck racecheck synccheck; do
  timeout 900 compute-sanitizer --tool $tool python scripts/sanitize_small.py > $O/san_$tool.log 2>&1
  echo "sanitizer $tool rc=$? : $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' $O/san_$tool.log | tail -1)" | tee -a $O/summary6.txt
done
timeout 400 python bench.py --steps 200 --warmup 5 > $O/bench6.json 2> $O/bench6.err
echo "bench rc=$?" | tee -a $O/summary6.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_r01e.csv python bench.py --steps 12 --warmup 3 --cpu-baseline-seconds 0 > $O/ncu_bench6.log 2>&1
echo "ncu rc=$?" | tee -a $O/summary6.txt
timeout 300 python scripts/config2_kernels.py > $O/config2_r01e.json 2> $O/config2_r01e.err
for fam in netlib_like sparse_pos; do
  timeout 600 python scripts/sparse_profile.py $fam 30000 30000 30 400 200 > $O/sparse_$fam.json 2> $O/sparse_$fam.err
  cat $O/sparse_$fam.json
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 6000 -c 1500 --csv --log-file $O/launches_sparse_$fam.csv python scripts/sparse_profile.py $fam 30000 30000 30 60 150 > $O/ncu_sparse_$fam.log 2>&1
done
timeout 900 python scripts/sparse_scale.py --pivots 1000 --cpu-seconds 20 > $O/sparse_scale_r01e.json 2> $O/sparse_scale_r01e.err
cat $O/sparse_scale_r01e.json
python - <<PY
import json
d = json.load(open("$O/bench6.json"))
print("bench", round(d["value"], 2), "piv/s", round(d["ms_per_step"], 4), "ms  e2e", round(d["e2e"]["value"], 2), "frac", round(d["roofline"]["frac"], 4), "price_v ms", round(d["roofline"]["avg_launch_ms"], 4), "launches", d["gpu_launches"])
PY
cat $O/summary6.txt
