#!/bin/bash
set -u
mkdir -p gpurun_out
OUT=gpurun_out
export PYTHONUNBUFFERED=1
python -m lama_b200.build > $OUT/build_q.log 2>&1; echo "build rc=$?" | tee $OUT/summary_q.txt
timeout 300 python -m pytest tests/test_gpu_parity.py -q --timeout 120 -x -k "generator_golden or stem or head or inpaint or u8 or conv" > $OUT/pytest_q_small.log 2>&1; echo "pytest small rc=$?" | tee -a $OUT/summary_q.txt
grep -E "^(FAILED|ERROR)|passed|failed" $OUT/pytest_q_small.log | tail -4 | tee -a $OUT/summary_q.txt
if grep -q "failed\|rc=124" $OUT/summary_q.txt; then echo "small tests failed: stop" | tee -a $OUT/summary_q.txt; exit 0; fi
for tw in 8 16; do
  FFCB_TC_ROWS_TW=$tw timeout 400 python bench.py --steps 10 --no-cpu-baseline --no-torch-cuda-baseline --no-fp32-arm --io f32 > $OUT/bench_q_tw$tw.json 2> $OUT/bench_q_tw$tw.err
  python -c "import json; d=json.load(open('$OUT/bench_q_tw$tw.json')); print('rows TW $tw', round(d['value'],1), 'img/s', round(d['ms_per_step'],2), d['launches_per_step'], d['clocks'])" | tee -a $OUT/summary_q.txt
  FFCB_TC_ROWS_TW=$tw timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT/launches_q$tw.csv python tools/profile_step.py bf16x3 32 > $OUT/prof_q.log 2>&1
  python tools/summarize_launches.py $OUT/launches_q$tw.csv $OUT/call_order.txt > $OUT/launches_q$tw.txt 2>&1; grep -E "stem 7x7|head 7x7|convT|total" $OUT/launches_q$tw.txt | tee -a $OUT/summary_q.txt
done
timeout 900 python -m pytest tests -m gpu -q --timeout 300 -x > $OUT/pytest_q.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/summary_q.txt
grep -E "^(FAILED|ERROR)|passed|failed" $OUT/pytest_q.log | tail -6 | tee -a $OUT/summary_q.txt
