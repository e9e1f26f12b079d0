#!/bin/bash
# round 2, call A: baseline of the round-1 tree + torch-CUDA reference rows (bench.py torch_cuda_baseline leg)
set -u
mkdir -p gpurun_out
OUT=gpurun_out
export PYTHONUNBUFFERED=1
python -c "import __graft_entry__ as g; g.build(); g.smoke()" > $OUT/smoke.log 2>&1; echo "smoke rc=$?" | tee $OUT/summary.txt
timeout 900 python -m pytest tests -m gpu -q --timeout 300 --durations=10 -x > $OUT/pytest.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/summary.txt
tail -15 $OUT/pytest.log | tee -a $OUT/summary.txt
timeout 600 python bench.py > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?" | tee -a $OUT/summary.txt
cat $OUT/bench.json | tee -a $OUT/summary.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file $OUT/launches_bench.csv \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-torch-cuda-baseline --io f32 > $OUT/bench_under_ncu.json 2> $OUT/bench_under_ncu.err
echo "ncu launch list rc=$?" | tee -a $OUT/summary.txt
python tools/summarize_launches.py $OUT/launches_bench.csv > $OUT/launches_bench.txt 2>&1; head -20 $OUT/launches_bench.txt | tee -a $OUT/summary.txt
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"plane64" -c 2 -f -o $OUT/plane_fft_ncu \
  python tools/fft_microbench.py --fu-only > $OUT/plane_ncu.log 2>&1; echo "ncu plane rc=$?" | tee -a $OUT/summary.txt
