#!/bin/bash
set -u
mkdir -p gpurun_out
OUT=gpurun_out
export PYTHONUNBUFFERED=1
python -m lama_b200.build > $OUT/build_s.log 2>&1; echo "build rc=$?" | tee $OUT/summary_s.txt
timeout 900 python -m pytest tests -m gpu -q --timeout 300 -x > $OUT/pytest_s.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/summary_s.txt
grep -E "^(FAILED|ERROR)|passed|failed" $OUT/pytest_s.log | tail -6 | tee -a $OUT/summary_s.txt
TC_OPS="stem 7x7,head 7x7 rows,convT phase 11,convT phase 00,convl2l|convl2g" timeout 300 python tools/tc_microbench.py > $OUT/tc_s.txt 2>&1; echo "tc rc=$?" | tee -a $OUT/summary_s.txt
tail -6 $OUT/tc_s.txt | cut -c1-130 | tee -a $OUT/summary_s.txt
timeout 300 python tools/tc_microbench.py > $OUT/tc_s_block.txt 2>&1; tail -7 $OUT/tc_s_block.txt | cut -c1-130 | tee -a $OUT/summary_s.txt
timeout 400 python bench.py --steps 10 --no-cpu-baseline --no-torch-cuda-baseline --no-fp32-arm --io f32 > $OUT/bench_s.json 2> $OUT/bench_s.err
python -c "import json; d=json.load(open('$OUT/bench_s.json')); print('bench', round(d['value'],1), 'img/s', round(d['ms_per_step'],2), d['launches_per_step'], d['clocks'])" | tee -a $OUT/summary_s.txt
