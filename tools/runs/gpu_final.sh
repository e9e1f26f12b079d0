#!/bin/bash
# Final validation: smoke, memcheck, full GPU tests, default bench (with CPU baseline), reference arm.
set -u
mkdir -p gpurun_out
OUT=gpurun_out
export PYTHONUNBUFFERED=1
python -c "import __graft_entry__ as g; g.build(); g.smoke()" > $OUT/smoke.log 2>&1; echo "smoke rc=$?" | tee $OUT/summary.txt
tail -1 $OUT/smoke.log | tee -a $OUT/summary.txt
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 --print-limit 20 python tools/sanitize_step.py > $OUT/sanitizer.log 2>&1; echo "memcheck rc=$?" | tee -a $OUT/summary.txt
grep -E "ERROR SUMMARY|Invalid|out of bounds|bf16x3|fp32|smoke ok" $OUT/sanitizer.log | head -12 | tee -a $OUT/summary.txt
timeout 900 python -m pytest tests -m gpu -q --timeout 300 > $OUT/pytest.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/summary.txt
tail -3 $OUT/pytest.log | tee -a $OUT/summary.txt
timeout 600 python bench.py > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?" | tee -a $OUT/summary.txt
cat $OUT/bench.json | tee -a $OUT/summary.txt; tail -3 $OUT/bench.err | tee -a $OUT/summary.txt
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/bench_ref.json 2> $OUT/bench_ref.err; echo "ref rc=$?" | tee -a $OUT/summary.txt
cat $OUT/bench_ref.json | tee -a $OUT/summary.txt
