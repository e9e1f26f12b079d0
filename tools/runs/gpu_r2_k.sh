#!/bin/bash
# round 2, call K: measurements + profiles of the validated tree (after the elect.sync issuer)
set -u
mkdir -p gpurun_out
OUT=gpurun_out
export PYTHONUNBUFFERED=1
python -c "import __graft_entry__ as g; g.build()" > $OUT/build_k.log 2>&1; echo "build rc=$?" | tee $OUT/summary_k.txt
LAMA_B200_FU_LAYOUT=nhwc FFCB_TC_FORCE_IL=1 timeout 300 python tools/tc_microbench.py > $OUT/tc_k_nhwc_forceil.txt 2>&1; echo "tc nhwc force-IL rc=$?" | tee -a $OUT/summary_k.txt
tail -7 $OUT/tc_k_nhwc_forceil.txt | tee -a $OUT/summary_k.txt
timeout 200 python tools/fft_microbench.py --chain > $OUT/fu_chain_k.jsonl 2> $OUT/fu_chain_k.err; cat $OUT/fu_chain_k.jsonl | tee -a $OUT/summary_k.txt
timeout 900 python bench.py > $OUT/bench_k.json 2> $OUT/bench_k.err; echo "bench rc=$?" | tee -a $OUT/summary_k.txt
python - <<PY | tee -a $OUT/summary_k.txt
import json
d = json.load(open("$OUT/bench_k.json"))
fu = d["roofline"]["fourier_unit"]
print(round(d["value"],1), "img/s", round(d["ms_per_step"],2), "ms; e2e", round(d["e2e"]["value"],1), "; L", round(d["roofline"]["ms_per_launch"]*1e3,1), "us frac", round(d["roofline"]["frac"],3), "; FU cold", round(fu["ms"]*1e3,1), "us frac", round(fu["frac"],3), fu["per_kernel"], fu["layout"])
print("torch-cuda", json.dumps(d["torch_cuda_baseline"]))
print("fp32 arm", d.get("fp32_arm")); print("cpu", d["cpu_baseline"]); print("clocks", d["clocks"])
PY
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $OUT/bench_k_reference.json 2> $OUT/bench_k_reference.err; echo "reference arm rc=$?" | tee -a $OUT/summary_k.txt
cat $OUT/bench_k_reference.json | cut -c1-400 | tee -a $OUT/summary_k.txt
timeout 300 python bench.py --size 256 --steps 10 --no-cpu-baseline --no-torch-cuda-baseline --no-fp32-arm --io f32 > $OUT/bench_k_256.json 2> $OUT/bench_k_256.err; echo "bench 256 rc=$?" | tee -a $OUT/summary_k.txt
timeout 400 python bench.py --size 1024 --steps 5 --no-cpu-baseline --no-torch-cuda-baseline --no-fp32-arm --io f32 > $OUT/bench_k_1024.json 2> $OUT/bench_k_1024.err; echo "bench 1024 rc=$?" | tee -a $OUT/summary_k.txt
python -c "
import json
for s in ('256','1024'):
    d=json.load(open('$OUT/bench_k_%s.json'%s)); print(s, round(d['value'],1), 'img/s', round(d['ms_per_step'],2), 'ms', d['launches_per_step'])" | tee -a $OUT/summary_k.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file $OUT/launches_bench_cmd_k.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-torch-cuda-baseline --no-fp32-arm --io f32 > $OUT/bench_under_ncu_k.json 2> $OUT/bench_under_ncu_k.err; echo "bench-cmd launch list rc=$?" | tee -a $OUT/summary_k.txt
python tools/summarize_launches.py $OUT/launches_bench_cmd_k.csv > $OUT/launches_bench_cmd_k.txt 2>&1; head -14 $OUT/launches_bench_cmd_k.txt | tee -a $OUT/summary_k.txt
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT/launches_k.csv python tools/profile_step.py bf16x3 32 > $OUT/prof_k.log 2>&1; echo "launch list rc=$?" | tee -a $OUT/summary_k.txt
python tools/summarize_launches.py $OUT/launches_k.csv $OUT/call_order.txt > $OUT/launches_k.txt 2>&1; head -36 $OUT/launches_k.txt | tee -a $OUT/summary_k.txt
ONCE=1 timeout 400 ncu --set full --clock-control none --import-source on --profile-from-start off -f -o $OUT/r02_block_ops_final python tools/tc_microbench.py > $OUT/ncu_ops_k.log 2>&1; echo "ncu block ops rc=$?" | tee -a $OUT/summary_k.txt
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 --print-limit 20 python tools/sanitize_step.py > $OUT/sanitizer_k.log 2>&1; echo "memcheck rc=$?" | tee -a $OUT/summary_k.txt
grep -E "ERROR SUMMARY" $OUT/sanitizer_k.log | head -3 | tee -a $OUT/summary_k.txt
timeout 900 python -m pytest tests -m gpu -q --timeout 300 -x > $OUT/pytest_k.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/summary_k.txt
grep -E "^(FAILED|ERROR)|passed|failed" $OUT/pytest_k.log | tail -8 | tee -a $OUT/summary_k.txt
