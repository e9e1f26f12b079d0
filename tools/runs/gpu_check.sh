#!/bin/bash
# One gpurun call: smoke, sanitizer on smoke, GPU parity tests, short bench.  Output -> gpurun_out/
set -u
mkdir -p gpurun_out
OUT=gpurun_out
export PYTHONUNBUFFERED=1
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/nvsmi.txt 2>&1
echo "== build+smoke" | tee $OUT/summary.txt
timeout 600 python -c "import __graft_entry__ as g; g.build(); g.smoke()" > $OUT/smoke.log 2>&1; echo "smoke rc=$?" | tee -a $OUT/summary.txt
tail -3 $OUT/smoke.log | tee -a $OUT/summary.txt
if [ "${SANITIZE:-1}" = "1" ]; then
  echo "== compute-sanitizer memcheck (smoke)" | tee -a $OUT/summary.txt
  timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 --print-limit 20 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/sanitizer.log 2>&1; echo "sanitizer rc=$?" | tee -a $OUT/summary.txt
  grep -E "ERROR SUMMARY|Invalid|out of bounds" $OUT/sanitizer.log | head -10 | tee -a $OUT/summary.txt
fi
echo "== pytest -m gpu" | tee -a $OUT/summary.txt
timeout ${PYTEST_TIMEOUT:-1500} python -m pytest tests -m gpu -q -x --timeout 600 ${PYTEST_ARGS:-} > $OUT/pytest.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/summary.txt
tail -40 $OUT/pytest.log | tee -a $OUT/summary.txt
echo "== bench" | tee -a $OUT/summary.txt
timeout 900 python bench.py --steps ${BENCH_STEPS:-3} --warmup 3 ${BENCH_ARGS:-} > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?" | tee -a $OUT/summary.txt
cat $OUT/bench.json | tee -a $OUT/summary.txt
tail -5 $OUT/bench.err | tee -a $OUT/summary.txt
