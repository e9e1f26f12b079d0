#!/bin/bash
set -u
mkdir -p gpurun_out
OUT=gpurun_out
export PYTHONUNBUFFERED=1
python -m lama_b200.build > $OUT/build_x.log 2>&1; echo "build rc=$?" | tee $OUT/summary_x.txt
timeout 300 python bench.py --size 2048 --batch 8 --steps 3 --no-cpu-baseline --no-torch-cuda-baseline --no-fp32-arm --io f32 > $OUT/bench_x_2048.json 2> $OUT/bench_x_2048.err; echo "bench 2048 rc=$?" | tee -a $OUT/summary_x.txt
python -c "import json; d=json.load(open('$OUT/bench_x_2048.json')); print('2048 bs8', round(d['value'],2), 'img/s', round(d['ms_per_step'],1), 'ms', d['launches_per_step'])" | tee -a $OUT/summary_x.txt
timeout 400 python tools/refine_bench.py --size 1024 > $OUT/refine_x.json 2> $OUT/refine_x.err; echo "refine rc=$?" | tee -a $OUT/summary_x.txt
cat $OUT/refine_x.json | tee -a $OUT/summary_x.txt; tail -3 $OUT/refine_x.err
