#!/bin/bash
set -u
mkdir -p gpurun_out
OUT=gpurun_out
export PYTHONUNBUFFERED=1
python -m lama_b200.build > $OUT/build_o.log 2>&1; echo "build rc=$?" | tee $OUT/summary_o.txt
timeout 900 python -m pytest tests -m gpu -q --timeout 300 -x > $OUT/pytest_o.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/summary_o.txt
grep -E "^(FAILED|ERROR)|passed|failed" $OUT/pytest_o.log | tail -6 | tee -a $OUT/summary_o.txt
for rk in 0 1 0; do
  LAMA_B200_RING_KERNEL=$rk timeout 400 python bench.py --steps 10 --no-cpu-baseline --no-torch-cuda-baseline --no-fp32-arm --io f32 > $OUT/bench_o_rk$rk.json 2> $OUT/bench_o_rk$rk.err
  python -c "import json; d=json.load(open('$OUT/bench_o_rk$rk.json')); print('ring kernel $rk', round(d['value'],1), 'img/s', round(d['ms_per_step'],2), d['launches_per_step'], d['clocks'], 'L us', round(d['roofline']['ms_per_launch']*1e3,1))" | tee -a $OUT/summary_o.txt
done
timeout 300 python tools/tc_microbench.py > $OUT/tc_o.txt 2>&1; tail -7 $OUT/tc_o.txt | cut -c1-110 | tee -a $OUT/summary_o.txt
