#!/bin/bash
# Second-revision plane FFT kernels: correctness (kernel level, bit-identity with revision 1, generator level under
# the env switches), microbenchmark, and the headline bench with them enabled.
set -u
mkdir -p gpurun_out
OUT=gpurun_out
timeout 120 python -m pytest tests/test_gpu_parity.py -q --timeout 100 \
  -k "second_revision or rfft2_irfft2 or inverse_plane" > $OUT/v2_pytest_a.log 2>&1
echo "pytest A rc=$?"; tail -2 $OUT/v2_pytest_a.log
export FFCB_FFT_PLANE_FWD=2 FFCB_FFT_INV_PLANE=3
timeout 150 python -m pytest tests/test_gpu_parity.py -q --timeout 120 \
  -k "(big_lama_generator_vs_oracle and bf16x3 and (512-1-1 or size2)) or fourier_unit_golden or baseline_config1 or fft_round_trip" \
  > $OUT/v2_pytest_b.log 2>&1
echo "pytest B (v2 via env) rc=$?"; tail -2 $OUT/v2_pytest_b.log
timeout 60 python tools/fft_microbench.py --v2 > $OUT/v2_fft_microbench.jsonl 2> $OUT/v2_fft_microbench.err
echo "microbench rc=$?"; cut -c1-200 $OUT/v2_fft_microbench.jsonl
timeout 120 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --io f32 > $OUT/v2_bench.json 2> $OUT/v2_bench.err
echo "bench (v2) rc=$?"; head -c 420 $OUT/v2_bench.json; echo
