#!/bin/bash
set -u
mkdir -p gpurun_out
OUT=gpurun_out
python -c "import __graft_entry__ as g; g.build()" > $OUT/build.log 2>&1
timeout 600 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:"plane64" -s 4 -c 2 -o $OUT/plane_full -f python tools/profile_step.py bf16x3 32 > $OUT/prof2.log 2>&1
echo "plane full rc=$?"
