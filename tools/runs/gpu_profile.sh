#!/bin/bash
set -u
mkdir -p gpurun_out
OUT=gpurun_out
python -c "import __graft_entry__ as g; g.build()" > $OUT/build.log 2>&1
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT/launches_bf16x3.csv python tools/profile_step.py bf16x3 32 > $OUT/prof1.log 2>&1
echo "launch list rc=$?"
timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:conv_tc_kernel -s 20 -c 2 -o $OUT/conv_tc_full -f python tools/profile_step.py bf16x3 32 > $OUT/prof2.log 2>&1
echo "conv_tc full rc=$?"
timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:"rfft_rows|fft_cols|irfft_rows" -s 8 -c 4 -o $OUT/fft_full -f python tools/profile_step.py bf16x3 32 > $OUT/prof3.log 2>&1
echo "fft full rc=$?"
ls -la $OUT
