#!/bin/bash
set -u
O=gpurun_out
mkdir -p $O
( time timeout 600 python -m pytest tests/test_deep_gpu.py -q -m gpu ) > $O/t11.log 2>&1
echo "deep tests rc=$?" | tee $O/summary11.txt
tail -25 $O/t11.log
