#!/bin/bash
# round 2, final refresh: measurements + profiles of the validated tree
set -u
mkdir -p gpurun_out
OUT=gpurun_out
export PYTHONUNBUFFERED=1
python -c "import __graft_entry__ as g; g.build()" > $OUT/build_u.log 2>&1; echo "build rc=$?" | tee $OUT/summary_u.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke_u.log 2>&1; echo "smoke rc=$?" | tee -a $OUT/summary_u.txt; tail -2 $OUT/smoke_u.log | tee -a $OUT/summary_u.txt
timeout 200 python tools/fft_microbench.py --chain > $OUT/fu_chain_u.jsonl 2> $OUT/fu_chain_u.err; cat $OUT/fu_chain_u.jsonl | tee -a $OUT/summary_u.txt
timeout 900 python bench.py > $OUT/bench_u.json 2> $OUT/bench_u.err; echo "bench rc=$?" | tee -a $OUT/summary_u.txt
python - <<PY | tee -a $OUT/summary_u.txt
import json
d = json.load(open("$OUT/bench_u.json"))
fu = d["roofline"]["fourier_unit"]
print(round(d["value"],1), "img/s", round(d["ms_per_step"],2), "ms; e2e", round(d["e2e"]["value"],1), "; L", round(d["roofline"]["ms_per_launch"]*1e3,1), "us frac", round(d["roofline"]["frac"],3), "; FU cold", round(fu["ms"]*1e3,1), "us frac", round(fu["frac"],3), fu["per_kernel"], fu["layout"])
print("torch-cuda", json.dumps(d["torch_cuda_baseline"]))
print("fp32 arm", d.get("fp32_arm")); print("cpu", d["cpu_baseline"]); print("clocks", d["clocks"])
PY
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $OUT/bench_u_reference.json 2> $OUT/bench_u_reference.err; echo "reference arm rc=$?" | tee -a $OUT/summary_u.txt
cat $OUT/bench_u_reference.json | cut -c1-400 | tee -a $OUT/summary_u.txt
timeout 300 python bench.py --size 256 --steps 10 --no-cpu-baseline --no-torch-cuda-baseline --no-fp32-arm --io f32 > $OUT/bench_u_256.json 2> $OUT/bench_u_256.err; echo "bench 256 rc=$?" | tee -a $OUT/summary_u.txt
timeout 400 python bench.py --size 1024 --steps 5 --no-cpu-baseline --no-torch-cuda-baseline --no-fp32-arm --io f32 > $OUT/bench_u_1024.json 2> $OUT/bench_u_1024.err; echo "bench 1024 rc=$?" | tee -a $OUT/summary_u.txt
python -c "
import json
for s in ('256','1024'):
    d=json.load(open('$OUT/bench_u_%s.json'%s)); print(s, round(d['value'],1), 'img/s', round(d['ms_per_step'],2), 'ms', d['launches_per_step'])" | tee -a $OUT/summary_u.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file $OUT/launches_bench_cmd_u.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-torch-cuda-baseline --no-fp32-arm --io f32 > $OUT/bench_under_ncu_u.json 2> $OUT/bench_under_ncu_u.err; echo "bench-cmd launch list rc=$?" | tee -a $OUT/summary_u.txt
python tools/summarize_launches.py $OUT/launches_bench_cmd_u.csv > $OUT/launches_bench_cmd_u.txt 2>&1; head -14 $OUT/launches_bench_cmd_u.txt | tee -a $OUT/summary_u.txt
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT/launches_u.csv python tools/profile_step.py bf16x3 32 > $OUT/prof_u.log 2>&1; echo "launch list rc=$?" | tee -a $OUT/summary_u.txt
python tools/summarize_launches.py $OUT/launches_u.csv $OUT/call_order.txt > $OUT/launches_u.txt 2>&1; head -36 $OUT/launches_u.txt | tee -a $OUT/summary_u.txt
ONCE=1 timeout 400 ncu --set full --clock-control none --import-source on --profile-from-start off -f -o $OUT/r02_block_ops_final2 python tools/tc_microbench.py > $OUT/ncu_ops_u.log 2>&1; echo "ncu block ops rc=$?" | tee -a $OUT/summary_u.txt
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 --print-limit 20 python tools/sanitize_step.py > $OUT/sanitizer_u.log 2>&1; echo "memcheck rc=$?" | tee -a $OUT/summary_u.txt
grep -E "ERROR SUMMARY" $OUT/sanitizer_u.log | head -3 | tee -a $OUT/summary_u.txt
timeout 900 python -m pytest tests -m gpu -q --timeout 300 -x > $OUT/pytest_u.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/summary_u.txt
grep -E "^(FAILED|ERROR)|passed|failed" $OUT/pytest_u.log | tail -8 | tee -a $OUT/summary_u.txt
