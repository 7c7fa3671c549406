#!/bin/bash
set -u
O=gpurun_out
mkdir -p $O
( time timeout 600 python -m pytest tests/test_fused_gpu.py tests/test_parity_gpu.py -q -m gpu -x ) > $O/t3.log 2>&1
echo "tests rc=$?" | tee $O/summary3.txt
tail -3 $O/t3.log
timeout 900 python scripts/price_sweep.py > $O/sweep2.jsonl 2> $O/sweep2.err
echo "sweep rc=$?" | tee -a $O/summary3.txt
cat $O/sweep2.jsonl
timeout 400 python bench.py --steps 200 --warmup 5 --cpu-baseline-seconds 0 > $O/bench3.json 2> $O/bench3.err
cat $O/bench3.json
cat $O/summary3.txt
