#!/bin/bash
# pytest (all GPU tests) + bench (bf16x3) + launch list.  Output -> gpurun_out/
set -u
mkdir -p gpurun_out
OUT=gpurun_out
export PYTHONUNBUFFERED=1
python -c "import __graft_entry__ as g; g.build()" > $OUT/build.log 2>&1
echo "== pytest -m gpu" | tee $OUT/summary.txt
timeout 900 python -m pytest tests -m gpu -q --timeout 300 ${PYTEST_ARGS:-} > $OUT/pytest.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/summary.txt
tail -15 $OUT/pytest.log | tee -a $OUT/summary.txt
echo "== bench" | tee -a $OUT/summary.txt
timeout 600 python bench.py --steps 5 --warmup 3 ${BENCH_ARGS:-} > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?" | tee -a $OUT/summary.txt
cat $OUT/bench.json | tee -a $OUT/summary.txt; tail -5 $OUT/bench.err | tee -a $OUT/summary.txt
if [ "${LAUNCHES:-1}" = "1" ]; then
  timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT/launches.csv python tools/profile_step.py bf16x3 32 > $OUT/prof1.log 2>&1
  echo "launch list rc=$?" | tee -a $OUT/summary.txt
fi
