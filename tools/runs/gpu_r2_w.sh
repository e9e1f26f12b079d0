#!/bin/bash
set -u
mkdir -p gpurun_out
OUT=gpurun_out
export PYTHONUNBUFFERED=1
python -m lama_b200.build > $OUT/build_w.log 2>&1; echo "build rc=$?" | tee $OUT/summary_w.txt
FFCB_TC_EPI2=1 timeout 400 python -m pytest tests/test_gpu_parity.py -q --timeout 120 -x -k "conv or generator_golden or stem or head or inpaint or u8 or resnet_block_golden" > $OUT/pytest_w.log 2>&1; echo "pytest epi2 rc=$?" | tee -a $OUT/summary_w.txt
grep -E "^(FAILED|ERROR)|passed|failed" $OUT/pytest_w.log | tail -4 | tee -a $OUT/summary_w.txt
for e2 in 0 1; do
FFCB_TC_EPI2=$e2 TC_OPS="stem 7x7,head 7x7 rows,convT phase 11,convT phase 00,convT phase 01" timeout 300 python tools/tc_microbench.py > $OUT/tc_w$e2.txt 2>&1; echo "tc epi2=$e2 rc=$?" | tee -a $OUT/summary_w.txt
tail -5 $OUT/tc_w$e2.txt | cut -c1-100 | tee -a $OUT/summary_w.txt
FFCB_TC_EPI2=$e2 timeout 400 python bench.py --steps 10 --no-cpu-baseline --no-torch-cuda-baseline --no-fp32-arm --io f32 > $OUT/bench_w$e2.json 2> $OUT/bench_w$e2.err
python -c "import json; d=json.load(open('$OUT/bench_w$e2.json')); print('epi2 $e2', round(d['value'],1), 'img/s', round(d['ms_per_step'],2), d['clocks'])" | tee -a $OUT/summary_w.txt
done
