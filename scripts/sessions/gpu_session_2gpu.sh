#!/bin/bash
# r02, two GPUs of one box (gpurun --gpus 2): the one-process-per-GPU exchange paths against the oracle (peer memory and NCCL
# all-gather; dense and sparse storage), then the bench lines at N = 2
set -u
O=gpurun_out/r02g2
mkdir -p $O
nvidia-smi -L > $O/gpus.txt 2>&1
( time timeout 900 python -m pytest tests/test_sharded_gpu.py tests/test_sparse_gpu.py tests/test_fullsize_gpu.py -q --durations=4 ) > $O/tests_2gpu.log 2>&1
echo "sharded + sparse + fullsize tests rc=$?" | tee $O/summary.txt
tail -8 $O/tests_2gpu.log
cp gpurun_out/nccl_two_process_p2p*.log $O/ 2>/dev/null
cat $O/nccl_two_process_p2p1.log 2>/dev/null | tail -12
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29571 bench.py --gpus 2 --steps 20 --warmup 5 > $O/bench_2gpu.json 2> $O/bench_2gpu.err
echo "bench 2 gpu rc=$?" | tee -a $O/summary.txt
python - <<'PY'
import json
try:
    d=json.load(open('gpurun_out/r02g2/bench_2gpu.json'))
    print('N=2 config3', d['value'], d['ms_per_step'], d['run_detail']['parallelism'])
    for k,v in d.get('extra',{}).items(): print(k, v.get('value'), v.get('ms_per_step'), v.get('error'), v.get('run_detail',{}).get('parallelism'))
except Exception as e: print('bench 2gpu parse failed', e)
PY
tail -3 $O/bench_2gpu.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29572 bench.py --gpus 2 --workload netlib_like --rows 100000 --cols 100000 --steps 3000 --warmup 20 > $O/bench_c4_2gpu.json 2> $O/bench_c4_2gpu.err
echo "bench c4 2 gpu rc=$?" | tee -a $O/summary.txt
python - <<'PY'
import json
try:
    d=json.load(open('gpurun_out/r02g2/bench_c4_2gpu.json'))
    print('N=2 config4', d['value'], d['ms_per_step'], d['run_detail']['parallelism'], d['roofline']['avg_launch_ms'])
except Exception as e: print('bench c4 2gpu parse failed', e)
PY
timeout 600 python bench.py --workload netlib_like --rows 100000 --cols 100000 --steps 3000 --warmup 20 --cpu-baseline-seconds 0 > $O/bench_c4_1gpu.json 2> $O/bench_c4_1gpu.err
python -c "
import json; d=json.load(open('$O/bench_c4_1gpu.json')); print('N=1 config4', d['value'], d['ms_per_step'], 'refactor share', d['run_detail']['refactor_share_of_wall'], 'price ms', d['roofline']['avg_launch_ms'])"
cat $O/summary.txt
