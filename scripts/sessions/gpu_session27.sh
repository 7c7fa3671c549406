#!/bin/bash
# r02 session 27: the whole GPU suite on the final commit
set -u
O=gpurun_out/r02s27
mkdir -p $O
( time timeout 150 python -m pytest tests -q -m gpu -x ) > $O/tests_gpu.log 2>&1
echo "gpu tests rc=$?" | tee $O/summary.txt
tail -5 $O/tests_gpu.log
