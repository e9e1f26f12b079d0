#!/bin/bash
# GPU validation of the uint8 predict path (row f1) + a regression slice of the float path it shares code with.
set -u
mkdir -p gpurun_out
OUT=gpurun_out
timeout 240 python -m pytest tests/test_gpu_parity.py -q --timeout 200 \
  -k "predict_u8 or serving_pipeline or small_generator_golden or errors_are_loud or inpaint_glue" \
  > $OUT/f1_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $OUT/f1_pytest.log
timeout 200 python bench.py --io both --steps 5 --warmup 3 --no-cpu-baseline > $OUT/f1_bench.json 2> $OUT/f1_bench.err
echo "bench rc=$?"; head -c 1500 $OUT/f1_bench.json; echo; tail -3 $OUT/f1_bench.err
