#!/bin/bash
set -u
O=gpurun_out/r02s11
mkdir -p $O
MLP_LU_EVERY=100000000 MLP_REFACTOR_TRACE=2 timeout 600 python bench.py --workload netlib_like --rows 100000 --cols 100000 --steps 3000 --warmup 20 --cpu-baseline-seconds 0 > $O/bench_c4_ev.json 2> $O/bench_c4_ev.err
grep -c "refactor event" $O/bench_c4_ev.err
grep "refactor event" $O/bench_c4_ev.err | awk '{print $NF, $(NF-1), $(NF-2)}' | sort | uniq -c | sort -rn | head
grep "refactor event" $O/bench_c4_ev.err | sed -n '1,12p;100,112p;300,330p'
grep "refactor trace" $O/bench_c4_ev.err
