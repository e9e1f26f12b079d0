#!/bin/bash
set -u
mkdir -p gpurun_out
OUT=gpurun_out
export PYTHONUNBUFFERED=1
python -c "import __graft_entry__ as g; g.build()" > $OUT/build_v.log 2>&1; echo "build rc=$?" | tee $OUT/summary_v.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke_v.log 2>&1; echo "smoke rc=$?" | tee -a $OUT/summary_v.txt; tail -1 $OUT/smoke_v.log | tee -a $OUT/summary_v.txt
timeout 900 python -m pytest tests -m gpu -q --timeout 300 > $OUT/pytest_v.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/summary_v.txt
grep -E "^(FAILED|ERROR)|passed|failed" $OUT/pytest_v.log | tail -6 | tee -a $OUT/summary_v.txt
timeout 900 python bench.py > $OUT/bench_v.json 2> $OUT/bench_v.err; echo "bench rc=$?" | tee -a $OUT/summary_v.txt
python - <<PY | tee -a $OUT/summary_v.txt
import json
d = json.load(open("$OUT/bench_v.json"))
r = d["roofline"]; fu = r["fourier_unit"]
print(round(d["value"],1), "img/s", round(d["ms_per_step"],2), "ms; e2e", round(d["e2e"]["value"],1), "; L", round(r["ms_per_launch"]*1e3,1), "us frac", round(r["frac"],3), "hot", round(r["ms_per_launch_hot"]*1e3,1), round(r["frac_hot_vs_sustained_peak"],3), "; FU cold", round(fu["ms"]*1e3,1), "us frac", round(fu["frac"],3), fu["traffic"])
print("clocks", d["clocks"], "launches", d["gpu_launches"], d["launches_per_step"])
PY
