#!/bin/bash
# gpurun --gpus N -- 'bash scripts/gpu_scaling.sh N': the 50k x 50k bench (BASELINE config 3) and config 5 (50k x 200k)
# column-sharded over N GPUs, one process per GPU.  torchrun's own parser grabs "--m": use --rows / --cols.
set -u
N=${1:-2}
O=gpurun_out
mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531"
timeout 420 $TR bench.py --gpus $N --steps 200 --warmup 5 > $O/bench_50k_${N}gpu.json 2> $O/bench_50k_${N}gpu.err
echo "50k x 50k at $N GPUs rc=$?"
timeout 600 $TR bench.py --gpus $N --rows 50000 --cols 200000 --steps 100 --warmup 5 > $O/bench_c5_${N}gpu.json 2> $O/bench_c5_${N}gpu.err
echo "50k x 200k at $N GPUs rc=$?"
cat $O/bench_50k_${N}gpu.json $O/bench_c5_${N}gpu.json
