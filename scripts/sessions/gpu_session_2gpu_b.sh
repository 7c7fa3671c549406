#!/bin/bash
# r02 final, two GPUs of one box (gpurun --gpus 2): the one-process-per-GPU exchange paths against the oracle (dense and sparse
# storage, the sparse engines with product-form refreshes), then the driver's bench line at N = 2 with its extras
set -u
O=gpurun_out/r02g3
mkdir -p $O
nvidia-smi -L > $O/gpus.txt 2>&1
( time timeout 600 python -m pytest tests/test_sharded_gpu.py -q --durations=3 ) > $O/tests_2gpu.log 2>&1
echo "sharded tests rc=$?" | tee $O/summary.txt
tail -6 $O/tests_2gpu.log
cp gpurun_out/nccl_two_process_p2p*.log $O/ 2>/dev/null
cat $O/nccl_two_process_p2p1.log 2>/dev/null | tail -10
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29571 bench.py --gpus 2 --steps 20 --warmup 5 > $O/bench_2gpu.json 2> $O/bench_2gpu.err
echo "bench 2 gpu rc=$?" | tee -a $O/summary.txt
python - <<'PY'
import json
try:
    d=json.load(open('gpurun_out/r02g3/bench_2gpu.json'))
    print('N=2 config3', d['value'], d['ms_per_step'], d['run_detail']['parallelism'])
    for k,v in d.get('extra',{}).items(): print(k, v.get('value'), v.get('ms_per_step'), v.get('error'), v.get('run_detail',{}).get('parallelism'), v.get('run_detail',{}).get('of_them_product_form_refreshes'), v.get('run_detail',{}).get('refactors_in_region'))
except Exception as e: print('bench 2gpu parse failed', e)
PY
tail -3 $O/bench_2gpu.err
cat $O/summary.txt
