#!/bin/bash
set -u
mkdir -p gpurun_out
OUT=gpurun_out
export PYTHONUNBUFFERED=1
python -m lama_b200.build > $OUT/build_p.log 2>&1; echo "build rc=$?" | tee $OUT/summary_p.txt
# rows-resident head / stem: first the small cases under a short timeout (a hang must not eat the box)
timeout 300 python -m pytest tests/test_gpu_parity.py -q --timeout 120 -x -k "generator_golden or stem or head or inpaint or u8" > $OUT/pytest_p_small.log 2>&1; echo "pytest small rc=$?" | tee -a $OUT/summary_p.txt
grep -E "^(FAILED|ERROR)|passed|failed" $OUT/pytest_p_small.log | tail -4 | tee -a $OUT/summary_p.txt
if grep -q "failed\|rc=124" $OUT/summary_p.txt; then echo "small tests failed: stop" | tee -a $OUT/summary_p.txt; exit 0; fi
timeout 900 python -m pytest tests -m gpu -q --timeout 300 -x > $OUT/pytest_p.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/summary_p.txt
grep -E "^(FAILED|ERROR)|passed|failed" $OUT/pytest_p.log | tail -6 | tee -a $OUT/summary_p.txt
for rows in 1 0; do
  FFCB_TC_ROWS=$rows timeout 400 python bench.py --steps 10 --no-cpu-baseline --no-torch-cuda-baseline --no-fp32-arm --io f32 > $OUT/bench_p_rows$rows.json 2> $OUT/bench_p_rows$rows.err
  python -c "import json; d=json.load(open('$OUT/bench_p_rows$rows.json')); print('rows-resident $rows', round(d['value'],1), 'img/s', round(d['ms_per_step'],2), d['launches_per_step'], d['clocks'])" | tee -a $OUT/summary_p.txt
done
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT/launches_p.csv python tools/profile_step.py bf16x3 32 > $OUT/prof_p.log 2>&1; echo "launch list rc=$?" | tee -a $OUT/summary_p.txt
python tools/summarize_launches.py $OUT/launches_p.csv $OUT/call_order.txt > $OUT/launches_p.txt 2>&1; sed -n 1,40p $OUT/launches_p.txt | tee -a $OUT/summary_p.txt
