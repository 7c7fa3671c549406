#!/bin/bash
set -u
O=gpurun_out
mkdir -p $O
( time timeout 600 python -m pytest tests/test_fused_gpu.py tests/test_fullsize_gpu.py -q -m gpu ) > $O/t2_new.log 2>&1
echo "new tests rc=$?" | tee $O/summary2.txt
tail -3 $O/t2_new.log
timeout 600 python scripts/price_sweep.py > $O/sweep.jsonl 2> $O/sweep.err
echo "sweep rc=$?" | tee -a $O/summary2.txt
cat $O/sweep.jsonl
# BASELINE config 5 on ONE GPU (80 GB of A): the N=1 point of the column-sharded runs
timeout 600 python bench.py --m 50000 --n 200000 --steps 100 --warmup 5 --cpu-baseline-seconds 0 > $O/bench_c5_1gpu.json 2> $O/bench_c5_1gpu.err
echo "c5 1gpu rc=$?" | tee -a $O/summary2.txt
cat $O/bench_c5_1gpu.json
# full ncu captures: the fused chain kernel and the price-out kernel
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_chain_primal -s 20 -c 2 -o $O/chain_r01d -f python bench.py --steps 30 --warmup 3 --cpu-baseline-seconds 0 > $O/ncu_chain.log 2>&1
echo "ncu chain rc=$?" | tee -a $O/summary2.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_price_partial_tma -s 20 -c 1 -o $O/price_r01d -f python bench.py --steps 30 --warmup 3 --cpu-baseline-seconds 0 > $O/ncu_price.log 2>&1
echo "ncu price rc=$?" | tee -a $O/summary2.txt
cat $O/summary2.txt
