#!/bin/bash
set -u
O=gpurun_out
mkdir -p $O
( time timeout 900 python -m pytest tests -q -m gpu ) > $O/t13.log 2>&1
echo "all gpu tests rc=$?" | tee $O/summary13.txt
tail -40 $O/t13.log
