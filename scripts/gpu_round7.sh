#!/bin/bash
# 8-GPU box: the 50k x 50k bench and BASELINE config 5 (50k x 200k) at N=8
set -u
O=gpurun_out
mkdir -p $O
TR8="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531"
timeout 420 $TR8 bench.py --gpus 8 --steps 200 --warmup 5 > $O/bench7_50k_8gpu.json 2> $O/bench7_50k_8gpu.err
echo "50k 8gpu rc=$?" | tee $O/summary7.txt
timeout 420 $TR8 bench.py --gpus 8 --rows 50000 --cols 200000 --steps 100 --warmup 5 > $O/bench7_c5_8gpu.json 2> $O/bench7_c5_8gpu.err
echo "c5 8gpu rc=$?" | tee -a $O/summary7.txt
for f in bench7_50k_8gpu bench7_c5_8gpu; do python - <<PY
import json
try:
    d = json.load(open("$O/$f.json"))
    print("$f", round(d["value"], 2), "piv/s", round(d["ms_per_step"], 4), "ms  e2e", round(d["e2e"]["value"], 2), "frac", round(d["roofline"]["frac"], 4),
          "price_v ms", round(d["roofline"]["avg_launch_ms"], 4), "launches", d["gpu_launches"], d["config"].get("parallelism"))
except Exception as e:
    print("$f failed", e)
PY
done | tee -a $O/summary7.txt
