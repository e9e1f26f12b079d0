#!/bin/bash
# tcgen05 bring-up: graded cases first; full tests + bench only if they pass.
set -u
mkdir -p gpurun_out
OUT=gpurun_out
export PYTHONUNBUFFERED=1
python -c "import __graft_entry__ as g; g.build()" > $OUT/build.log 2>&1
echo "== tc_debug" | tee $OUT/summary.txt
timeout 1200 python tools/tc_debug.py --timeout 100 > $OUT/tc_debug.log 2>&1
cat $OUT/tc_debug.log | cut -c1-1500 | tee -a $OUT/summary.txt
if grep -q '"ok": false' $OUT/tc_debug.log; then
  echo "tc_debug has failures: skipping full tests" | tee -a $OUT/summary.txt
  exit 0
fi
echo "== pytest -m gpu" | tee -a $OUT/summary.txt
timeout 1500 python -m pytest tests -m gpu -q --timeout 300 > $OUT/pytest.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/summary.txt
tail -30 $OUT/pytest.log | tee -a $OUT/summary.txt
echo "== bench bf16x3" | tee -a $OUT/summary.txt
timeout 600 python bench.py --steps 5 --warmup 3 --math bf16x3 --no-cpu-baseline > $OUT/bench_tc.json 2> $OUT/bench_tc.err; echo "bench rc=$?" | tee -a $OUT/summary.txt
cat $OUT/bench_tc.json | tee -a $OUT/summary.txt; tail -5 $OUT/bench_tc.err | tee -a $OUT/summary.txt
if [ "${LAUNCHES:-1}" = "1" ]; then
  timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT/launches.csv python tools/profile_step.py bf16x3 32 > $OUT/prof1.log 2>&1
  echo "launch list rc=$?" | tee -a $OUT/summary.txt
fi
