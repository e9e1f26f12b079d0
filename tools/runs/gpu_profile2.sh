#!/bin/bash
set -u
mkdir -p gpurun_out
OUT=gpurun_out
python -c "import __graft_entry__ as g; g.build()" > $OUT/build.log 2>&1
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT/launches_v2.csv python tools/profile_step.py bf16x3 32 > $OUT/prof1.log 2>&1
echo "launch list rc=$?"
# conv_tc launches in program order: 1..4 downs(3)+?, then per block: L, conv1, fu, G ... capture a conv1 and a G conv
timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:conv_tc_kernel -s 20 -c 4 -o $OUT/conv_tc_full_v2 -f python tools/profile_step.py bf16x3 32 > $OUT/prof2.log 2>&1
echo "conv_tc full rc=$?"
