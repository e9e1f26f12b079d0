#!/bin/bash
set -u
mkdir -p gpurun_out
OUT=gpurun_out
export PYTHONUNBUFFERED=1
python -m lama_b200.build > $OUT/build_t.log 2>&1; echo "build rc=$?" | tee $OUT/summary_t.txt
for st in 1 0; do
FFCB_TC_STACK=$st TC_OPS="stem 7x7,head 7x7 rows,convT phase 11,convT phase 00,convT phase 01,convl2l|convl2g" timeout 300 python tools/tc_microbench.py > $OUT/tc_t$st.txt 2>&1; echo "tc stack=$st rc=$?" | tee -a $OUT/summary_t.txt
tail -6 $OUT/tc_t$st.txt | cut -c1-100 | tee -a $OUT/summary_t.txt
done
