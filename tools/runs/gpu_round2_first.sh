#!/bin/bash
# First GPU call of the next round: re-establish the baseline after the end-of-round-1 kernel changes (default inverse
# plane kernel revision 2, mixed-radix FFT, uint8 predict path) that were validated piecewise only:
# smoke + full GPU suite + sanitizer, default bench, launch list of the bench command, ncu --set full of the plane FFT
# kernels and of the dominant contraction.  ~6-8 GPU-minutes.
set -u
mkdir -p gpurun_out
OUT=gpurun_out
export PYTHONUNBUFFERED=1
python -c "import __graft_entry__ as g; g.build(); g.smoke()" > $OUT/smoke.log 2>&1; echo "smoke rc=$?" | tee $OUT/summary.txt
timeout 1200 python -m pytest tests -m gpu -q --timeout 300 --durations=15 > $OUT/pytest.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/summary.txt
tail -25 $OUT/pytest.log | tee -a $OUT/summary.txt
timeout 600 python bench.py > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?" | tee -a $OUT/summary.txt
cat $OUT/bench.json | tee -a $OUT/summary.txt
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 8000 --csv --log-file $OUT/launches_bench.csv \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline --io f32 > $OUT/bench_under_ncu.json 2> $OUT/bench_under_ncu.err
echo "ncu launch list rc=$?" | tee -a $OUT/summary.txt
python tools/summarize_launches.py $OUT/launches_bench.csv > $OUT/launches_bench.txt 2>&1; head -14 $OUT/launches_bench.txt | tee -a $OUT/summary.txt
timeout 200 python tools/fft_microbench.py > $OUT/fft_microbench.jsonl 2> $OUT/fft_microbench.err; echo "fft microbench rc=$?" | tee -a $OUT/summary.txt
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"plane64" -c 2 -f -o $OUT/plane_fft_ncu \
  python tools/fft_microbench.py --fu-only > $OUT/plane_ncu.log 2>&1; echo "ncu plane rc=$?" | tee -a $OUT/summary.txt
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 --print-limit 20 python tools/sanitize_step.py > $OUT/sanitizer.log 2>&1
echo "memcheck rc=$?" | tee -a $OUT/summary.txt
grep -E "ERROR SUMMARY" $OUT/sanitizer.log | head -3 | tee -a $OUT/summary.txt
