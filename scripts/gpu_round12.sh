#!/bin/bash
set -u
O=gpurun_out
mkdir -p $O
( time timeout 600 python -m pytest tests/test_engine_abi_gpu.py -q -m gpu ) > $O/t12.log 2>&1
echo "abi tests rc=$?" | tee $O/summary12.txt
tail -30 $O/t12.log
