#!/bin/bash
set -u
mkdir -p gpurun_out
OUT=gpurun_out
export PYTHONUNBUFFERED=1
python -m lama_b200.build > $OUT/build_r.log 2>&1; echo "build rc=$?" | tee $OUT/summary_r.txt
TC_OPS="stem 7x7,head 7x7 rows,convT phase 11,convT phase 00,convl2l|convl2g" timeout 300 python tools/tc_microbench.py > $OUT/tc_r.txt 2>&1; echo "tc rc=$?" | tee -a $OUT/summary_r.txt
tail -7 $OUT/tc_r.txt | cut -c1-130 | tee -a $OUT/summary_r.txt
