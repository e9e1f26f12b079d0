#!/bin/bash
set -u
mkdir -p gpurun_out
OUT=gpurun_out
python -c "import __graft_entry__ as g; g.build()" > $OUT/build.log 2>&1
timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:conv_tc_kernel -s 21 -c 3 -o $OUT/conv_tc_full_v4 -f python tools/profile_step.py bf16x3 32 > $OUT/prof2.log 2>&1
echo "conv_tc full rc=$?"
