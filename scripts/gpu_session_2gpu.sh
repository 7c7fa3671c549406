#!/bin/bash
# r02, two GPUs of one box (gpurun --gpus 2): the one-process-per-GPU exchange paths against the oracle (peer memory and NCCL
# all-gather; dense and sparse storage), then the bench lines at N = 2
set -u
O=gpurun_out/r02g2
mkdir -p $O
nvidia-smi -L > $O/gpus.txt 2>&1
( time timeout 900 python -m pytest tests/test_sharded_gpu.py -q -k "nccl" --durations=4 ) > $O/tests_nccl.log 2>&1
echo "nccl tests rc=$?" | tee $O/summary.txt
tail -6 $O/tests_nccl.log
cp gpurun_out/nccl_two_process_p2p*.log $O/ 2>/dev/null
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29571 bench.py --gpus 2 --steps 20 --warmup 5 > $O/bench_2gpu.json 2> $O/bench_2gpu.err
echo "bench 2 gpu rc=$?" | tee -a $O/summary.txt
head -c 5000 $O/bench_2gpu.json
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29572 bench.py --gpus 2 --workload netlib_like --rows 100000 --cols 100000 --steps 3000 --warmup 20 > $O/bench_c4_2gpu.json 2> $O/bench_c4_2gpu.err
echo "bench c4 2 gpu rc=$?" | tee -a $O/summary.txt
head -c 3000 $O/bench_c4_2gpu.json
cat $O/summary.txt
