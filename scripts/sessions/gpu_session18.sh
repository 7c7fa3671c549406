#!/bin/bash
set -u
O=gpurun_out/r02s18
mkdir -p $O
timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 20000 -c 6000 --csv --log-file $O/launches_c4.csv \
  python bench.py --workload netlib_like --rows 100000 --cols 100000 --steps 1200 --warmup 10 --cpu-baseline-seconds 0 > $O/ncu_c4.log 2>&1
echo "ncu launch list c4 rc=$?" | tee $O/summary.txt
python scripts/summarize_ncu.py launches $O/launches_c4.csv $O/launches_c4_summary.md
head -60 $O/launches_c4_summary.md
