#!/bin/bash
set -u
O=gpurun_out
mkdir -p $O
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_lu_panel -s 100 -c 1 -o $O/lupanel_r01e -f python scripts/sparse_profile.py netlib_like 30000 30000 30 60 150 > $O/ncu_lupanel.log 2>&1
echo "ncu rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_core_inverse -s 20 -c 1 -o $O/coreinv_r01e -f python scripts/sparse_profile.py netlib_like 30000 30000 30 60 150 > $O/ncu_coreinv.log 2>&1
echo "ncu rc=$?"
