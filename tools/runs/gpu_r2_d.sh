#!/bin/bash
# round 2, call D: where did the conv_tc kernels lose time since round 1?  (fast-math A/B x layout A/B, per-op microbench)
set -u
mkdir -p gpurun_out
OUT=gpurun_out
export PYTHONUNBUFFERED=1
for fm in nofm fm; do
  if [ $fm = fm ]; then export LAMA_B200_NVCC_EXTRA="--use_fast_math"; else export LAMA_B200_NVCC_EXTRA=""; fi
  python -m lama_b200.build > $OUT/build_$fm.log 2>&1; echo "build $fm rc=$?" | tee -a $OUT/summary_d.txt
  for layout in planar nhwc; do
    LAMA_B200_FU_LAYOUT=$layout timeout 300 python tools/tc_microbench.py > $OUT/tc_${fm}_${layout}.txt 2>&1; echo "tc $fm $layout rc=$?" | tee -a $OUT/summary_d.txt
    cat $OUT/tc_${fm}_${layout}.txt | tee -a $OUT/summary_d.txt
  done
done
export LAMA_B200_NVCC_EXTRA=""
python -m lama_b200.build > /dev/null 2>&1
timeout 900 python -m pytest tests -m gpu -q --timeout 300 -x > $OUT/pytest_d.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/summary_d.txt
tail -8 $OUT/pytest_d.log | tee -a $OUT/summary_d.txt
