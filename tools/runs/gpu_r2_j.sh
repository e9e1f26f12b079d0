#!/bin/bash
set -u
mkdir -p gpurun_out
OUT=gpurun_out
export PYTHONUNBUFFERED=1
python -m lama_b200.build > $OUT/build_j.log 2>&1; echo "build rc=$?" | tee $OUT/summary_j.txt
timeout 900 python -m pytest tests -m gpu -q --timeout 300 -x > $OUT/pytest_j.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/summary_j.txt
grep -E "^(FAILED|ERROR)|passed|failed" $OUT/pytest_j.log | tail -8 | tee -a $OUT/summary_j.txt
for layout in planar nhwc; do
  LAMA_B200_FU_LAYOUT=$layout timeout 300 python tools/tc_microbench.py > $OUT/tc_j_${layout}.txt 2>&1; echo "tc $layout rc=$?" | tee -a $OUT/summary_j.txt
  tail -8 $OUT/tc_j_${layout}.txt | tee -a $OUT/summary_j.txt
done
timeout 400 python bench.py --steps 10 --no-cpu-baseline --no-torch-cuda-baseline --no-fp32-arm --io f32 > $OUT/bench_j.json 2> $OUT/bench_j.err
python -c "import json; d=json.load(open('$OUT/bench_j.json')); print('planar', round(d['value'],1), 'img/s', round(d['ms_per_step'],2), d['clocks'], 'L us', round(d['roofline']['ms_per_launch']*1e3,1))" | tee -a $OUT/summary_j.txt
LAMA_B200_FU_LAYOUT=nhwc timeout 400 python bench.py --steps 10 --no-cpu-baseline --no-torch-cuda-baseline --no-fp32-arm --io f32 > $OUT/bench_j_nhwc.json 2> $OUT/bench_j_nhwc.err
python -c "import json; d=json.load(open('$OUT/bench_j_nhwc.json')); print('nhwc', round(d['value'],1), 'img/s', round(d['ms_per_step'],2))" | tee -a $OUT/summary_j.txt
