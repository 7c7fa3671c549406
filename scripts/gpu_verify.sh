#!/bin/bash
# One visit to a 1-GPU B200 box (gpurun -- 'bash scripts/gpu_verify.sh'): the whole GPU suite, smoke(), compute-sanitizer on
# the small solves, the bench line with its reference arm, and the ncu launch list of the same bench command.
# Everything is logged under gpurun_out/ (merged back by gpurun); copy what should be judged into profiles/.
set -u
O=gpurun_out
mkdir -p $O
( time timeout 900 python -m pytest tests -q -m gpu ) > $O/tests_gpu.log 2>&1
echo "gpu tests rc=$?" | tee $O/verify_summary.txt
tail -3 $O/tests_gpu.log
( timeout 300 python -c "import __graft_entry__ as g; g.smoke()" ) > $O/smoke.log 2>&1
echo "smoke rc=$?" | tee -a $O/verify_summary.txt
for tool in memcheck racecheck synccheck; do
  timeout 900 compute-sanitizer --tool $tool python scripts/sanitize_small.py > $O/san_$tool.log 2>&1
  echo "sanitizer $tool rc=$? : $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' $O/san_$tool.log | tail -1)" | tee -a $O/verify_summary.txt
done
timeout 400 python bench.py --steps 200 --warmup 5 > $O/bench.json 2> $O/bench.err
echo "bench rc=$?" | tee -a $O/verify_summary.txt
timeout 300 python bench.py --impl reference --steps 12 --warmup 1 --ref-budget-seconds 60 > $O/bench_reference.json 2> $O/bench_reference.err
echo "bench reference arm rc=$?" | tee -a $O/verify_summary.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches.csv python bench.py --steps 12 --warmup 3 --cpu-baseline-seconds 0 > $O/ncu_bench.log 2>&1
echo "ncu launch list rc=$?" | tee -a $O/verify_summary.txt
cat $O/bench.json
cat $O/verify_summary.txt
