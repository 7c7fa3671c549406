#!/bin/bash
# 2-GPU box: sharded tests (NCCL + peer memory, fused chain), 50k x 50k and 50k x 200k at N=2 and N=1, wide-tile sweep, ncu
set -u
O=gpurun_out
mkdir -p $O
( time timeout 900 python -m pytest tests/test_sharded_gpu.py tests/test_fused_gpu.py -q -m gpu ) > $O/t4.log 2>&1
echo "tests rc=$?" | tee $O/summary4.txt
tail -3 $O/t4.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 600 $TR bench.py --gpus 2 --steps 200 --warmup 5 > $O/bench4_50k_2gpu.json 2> $O/bench4_50k_2gpu.err
echo "50k 2gpu rc=$?" | tee -a $O/summary4.txt
timeout 900 $TR bench.py --gpus 2 --m 50000 --n 200000 --steps 100 --warmup 5 > $O/bench4_c5_2gpu.json 2> $O/bench4_c5_2gpu.err
echo "c5 2gpu rc=$?" | tee -a $O/summary4.txt
timeout 600 python scripts/price_sweep.py --widths 6250,12500,25000 --tiles 512,1536,2048,2560,3136,4096 --splits 0 --pivots 30 > $O/sweep3.jsonl 2> $O/sweep3.err
echo "sweep rc=$?" | tee -a $O/summary4.txt
timeout 400 python bench.py --steps 200 --warmup 5 > $O/bench4_50k_1gpu.json 2> $O/bench4_50k_1gpu.err
echo "50k 1gpu rc=$?" | tee -a $O/summary4.txt
timeout 600 python bench.py --m 50000 --n 200000 --steps 100 --warmup 5 --cpu-baseline-seconds 0 > $O/bench4_c5_1gpu.json 2> $O/bench4_c5_1gpu.err
echo "c5 1gpu rc=$?" | tee -a $O/summary4.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_price_partial_tma -s 20 -c 1 -o $O/price_r01e -f python bench.py --steps 30 --warmup 3 --cpu-baseline-seconds 0 > $O/ncu_price2.log 2>&1
echo "ncu price rc=$?" | tee -a $O/summary4.txt
for f in bench4_50k_2gpu bench4_c5_2gpu bench4_50k_1gpu bench4_c5_1gpu; do python - <<PY
import json
try:
    d = json.load(open("$O/$f.json"))
    print("$f", round(d["value"], 2), "piv/s", round(d["ms_per_step"], 4), "ms  e2e", round(d["e2e"]["value"], 2), "frac", round(d["roofline"]["frac"], 4),
          "price_v ms", round(d["roofline"]["avg_launch_ms"], 4), "launches", d["gpu_launches"], d["config"].get("parallelism"))
except Exception as e:
    print("$f failed", e)
PY
done | tee -a $O/summary4.txt
cat $O/summary4.txt
