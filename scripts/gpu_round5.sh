#!/bin/bash
# 4-GPU box: BASELINE config 5 (50k x 200k, column-sharded) at N=2 and N=4, the 50k x 50k bench at N=4
set -u
O=gpurun_out
mkdir -p $O
TR2="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521"
TR4="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29522"
timeout 900 $TR4 bench.py --gpus 4 --rows 50000 --cols 200000 --steps 100 --warmup 5 > $O/bench5_c5_4gpu.json 2> $O/bench5_c5_4gpu.err
echo "c5 4gpu rc=$?" | tee $O/summary5.txt
timeout 900 $TR2 bench.py --gpus 2 --rows 50000 --cols 200000 --steps 100 --warmup 5 > $O/bench5_c5_2gpu.json 2> $O/bench5_c5_2gpu.err
echo "c5 2gpu rc=$?" | tee -a $O/summary5.txt
timeout 600 $TR4 bench.py --gpus 4 --steps 200 --warmup 5 > $O/bench5_50k_4gpu.json 2> $O/bench5_50k_4gpu.err
echo "50k 4gpu rc=$?" | tee -a $O/summary5.txt
for f in bench5_c5_4gpu bench5_c5_2gpu bench5_50k_4gpu; do python - <<PY
import json
try:
    d = json.load(open("$O/$f.json"))
    print("$f", round(d["value"], 2), "piv/s", round(d["ms_per_step"], 4), "ms  e2e", round(d["e2e"]["value"], 2), "frac", round(d["roofline"]["frac"], 4),
          "price_v ms", round(d["roofline"]["avg_launch_ms"], 4), "launches", d["gpu_launches"], d["config"].get("parallelism"))
except Exception as e:
    print("$f failed", e)
PY
done | tee -a $O/summary5.txt
tail -3 $O/bench5_c5_4gpu.err
