#!/bin/bash
set -u
O=gpurun_out
mkdir -p $O
( time timeout 900 python -m pytest tests -q -m gpu ) > $O/t10.log 2>&1
echo "all gpu tests rc=$?" | tee $O/summary10.txt
tail -3 $O/t10.log
( timeout 300 python -c "import __graft_entry__ as g; g.smoke()" ) > $O/smoke10.log 2>&1
echo "smoke rc=$?" | tee -a $O/summary10.txt
timeout 400 python bench.py --steps 200 --warmup 5 > $O/bench10.json 2> $O/bench10.err
echo "bench rc=$?" | tee -a $O/summary10.txt
timeout 300 python bench.py --impl reference --steps 12 --warmup 1 --ref-budget-seconds 60 > $O/bench10_ref.json 2> $O/bench10_ref.err
echo "bench ref rc=$?" | tee -a $O/summary10.txt
for fam in netlib_like sparse_pos; do
  timeout 600 python scripts/sparse_profile.py $fam 30000 30000 30 400 200 > $O/sparse10_$fam.json 2> $O/sparse10_$fam.err
  cat $O/sparse10_$fam.json
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 6000 -c 200 --csv --log-file $O/launches10_sparse_refactor.csv python scripts/sparse_profile.py netlib_like 30000 30000 30 60 150 > $O/ncu10_sparse.log 2>&1
timeout 900 python scripts/sparse_scale.py --pivots 1000 --cpu-seconds 20 > $O/sparse_scale10.json 2> $O/sparse_scale10.err
cat $O/sparse_scale10.json
python - <<PY
import json
d = json.load(open("$O/bench10.json"))
print("bench", round(d["value"], 2), "piv/s", round(d["ms_per_step"], 4), "ms  e2e", round(d["e2e"]["value"], 2), "frac", round(d["roofline"]["frac"], 4), "price_v ms", round(d["roofline"]["avg_launch_ms"], 4), "launches", d["gpu_launches"], d["clocks"])
print(open("$O/bench10_ref.json").read()[:400])
PY
cat $O/summary10.txt
