#!/bin/bash
set -u
O=gpurun_out
mkdir -p $O
( time timeout 900 python -m pytest tests -q -m gpu ) > $O/t8.log 2>&1
echo "all gpu tests rc=$?" | tee $O/summary8.txt
tail -3 $O/t8.log
for fam in netlib_like; do
  timeout 600 python scripts/sparse_profile.py $fam 30000 30000 30 400 200 > $O/sparse8_$fam.json 2> $O/sparse8_$fam.err
  cat $O/sparse8_$fam.json
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 6000 -c 200 --csv --log-file $O/launches8_sparse_$fam.csv python scripts/sparse_profile.py $fam 30000 30000 30 60 150 > $O/ncu8_sparse_$fam.log 2>&1
done
timeout 300 python bench.py --steps 200 --warmup 5 --cpu-baseline-seconds 0 > $O/bench8.json 2> $O/bench8.err
python - <<PY
import json
d = json.load(open("$O/bench8.json"))
print("bench", round(d["value"], 2), "piv/s", round(d["ms_per_step"], 4), "ms", "refactor wall", d["e2e"]["refactor_wall_s"], d["config"]["objective_after"])
PY
cat $O/summary8.txt
