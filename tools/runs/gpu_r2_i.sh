#!/bin/bash
set -u
mkdir -p gpurun_out
OUT=gpurun_out
export PYTHONUNBUFFERED=1
python -m lama_b200.build > $OUT/build_i.log 2>&1; echo "build rc=$?" | tee $OUT/summary_i.txt
timeout 900 python -m pytest tests -m gpu -q --timeout 300 > $OUT/pytest_i.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/summary_i.txt
grep -E "^(FAILED|ERROR)|passed|failed" $OUT/pytest_i.log | tail -8 | tee -a $OUT/summary_i.txt
timeout 300 python tools/tc_microbench.py > $OUT/tc_i_planar.txt 2>&1; echo "tc planar rc=$?" | tee -a $OUT/summary_i.txt
tail -8 $OUT/tc_i_planar.txt | tee -a $OUT/summary_i.txt
for chunk in 0 8 16; do
  LAMA_B200_FU_CHUNK=$chunk timeout 400 python bench.py --steps 10 --no-cpu-baseline --no-torch-cuda-baseline --no-fp32-arm --io f32 > $OUT/bench_i_chunk$chunk.json 2> $OUT/bench_i_chunk$chunk.err
  python -c "import json; d=json.load(open('$OUT/bench_i_chunk$chunk.json')); print('chunk $chunk', round(d['value'],1), 'img/s', round(d['ms_per_step'],2), d['clocks'])" | tee -a $OUT/summary_i.txt
done
for layout in planar nhwc; do
  ONCE=1 LAMA_B200_FU_LAYOUT=$layout timeout 400 ncu --set full --clock-control none --import-source on --profile-from-start off -f -o $OUT/r02_block_ops_$layout python tools/tc_microbench.py > $OUT/ncu_ops_$layout.log 2>&1; echo "ncu ops $layout rc=$?" | tee -a $OUT/summary_i.txt
done
