set -u
O=gpurun_out
mkdir -p $O
( timeout 600 python -m pytest tests -q -m gpu ) > $O/tests_gpu.log 2>&1
echo "gpu tests rc=$?" | tee $O/final_summary.txt
tail -2 $O/tests_gpu.log
( timeout 200 python -c "import __graft_entry__ as g; g.smoke()" ) > $O/smoke.log 2>&1
echo "smoke rc=$?" | tee -a $O/final_summary.txt
timeout 300 python bench.py --steps 200 --warmup 5 > $O/bench_final.json 2> $O/bench_final.err
echo "bench rc=$?" | tee -a $O/final_summary.txt
timeout 200 python tests/tools/tie_probe.py > $O/tie_probe.json 2> $O/tie_probe.err
echo "tie probe rc=$?" | tee -a $O/final_summary.txt
python - <<PY
import json
d = json.load(open("$O/bench_final.json"))
print("bench", round(d["value"], 2), "piv/s", round(d["ms_per_step"], 4), "ms frac", round(d["roofline"]["frac"], 4))
t = json.load(open("$O/tie_probe.json")); t.pop("detail"); print(t)
PY
