#!/bin/bash
# r02 session 1: GPU suite (new tie accounting, golden full-size traces, f4, row growth), smoke, tie probe, bench line, config-4 footnote
set -u
O=gpurun_out/r02s1
mkdir -p $O
nvidia-smi --query-gpu=name,memory.total --format=csv > $O/gpu.txt 2>&1; nproc >> $O/gpu.txt; free -g >> $O/gpu.txt
( time timeout 1500 python -m pytest tests -q -m gpu -x --durations=15 ) > $O/tests_gpu.log 2>&1
echo "gpu tests rc=$?" | tee $O/summary.txt
tail -25 $O/tests_gpu.log
( timeout 300 python -c "import __graft_entry__ as g; g.smoke()" ) > $O/smoke.log 2>&1
echo "smoke rc=$?" | tee -a $O/summary.txt
timeout 600 python tests/tools/tie_probe.py 60 > $O/tie_probe.json 2> $O/tie_probe.err
echo "tie probe rc=$?" | tee -a $O/summary.txt
timeout 600 python bench.py --steps 20 --warmup 5 > $O/bench.json 2> $O/bench.err
echo "bench rc=$?" | tee -a $O/summary.txt
cat $O/bench.json
timeout 900 python tests/tools/sparse_scale.py --pivots 3000 --cpu-seconds 20 > $O/sparse_scale.json 2> $O/sparse_scale.err
echo "sparse_scale rc=$?" | tee -a $O/summary.txt
cat $O/sparse_scale.json
cat $O/summary.txt
