#!/bin/bash
set -u
mkdir -p gpurun_out
OUT=gpurun_out
python -c "import __graft_entry__ as g; g.build(); g.smoke()" > $OUT/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 $OUT/smoke.log
timeout 900 python -m pytest tests -m gpu -q --timeout 600 > $OUT/pytest.log 2>&1; echo "pytest rc=$?"; tail -1 $OUT/pytest.log
timeout 600 python bench.py > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?"; wc -l $OUT/bench.json; head -c 400 $OUT/bench.json; echo
# launch list of the bench command itself (profiling recipe): compare SHARES, not absolutes
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 8000 --csv --log-file $OUT/launches_bench.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $OUT/bench_under_ncu.json 2> $OUT/bench_under_ncu.err
echo "ncu bench rc=$?"; grep -vc "^==" $OUT/launches_bench.csv
