#!/bin/bash
# GPU validation of row f2 (runtime mixed-radix FFT) + plane-kernel variants + FFT microbenchmark + ncu of the FFT kernels
set -u
mkdir -p gpurun_out
OUT=gpurun_out
timeout 150 python -m pytest tests/test_gpu_parity.py -q --timeout 120 \
  -k "fft_lengths_without or inverse_plane or forward_plane or two_pass" > $OUT/f2_pytest_a.log 2>&1
echo "pytest A rc=$?"; tail -2 $OUT/f2_pytest_a.log
FFCB_FFT_MIXED_RADIX=1 timeout 150 python -m pytest tests/test_gpu_parity.py -q --timeout 120 \
  -k "golden or predict_u8_bytes or rfft2_irfft2" > $OUT/f2_pytest_b.log 2>&1
echo "pytest B (mixed radix on) rc=$?"; tail -2 $OUT/f2_pytest_b.log
timeout 120 python tools/fft_microbench.py > $OUT/f2_fft_microbench.jsonl 2> $OUT/f2_fft_microbench.err
echo "microbench rc=$?"; cat $OUT/f2_fft_microbench.jsonl | cut -c1-230
timeout 150 ncu --set full --clock-control none --import-source on -k regex:"fft" -c 3 -f -o $OUT/f2_fft_ncu \
  python tools/fft_microbench.py --fu-only > $OUT/f2_ncu.log 2>&1
echo "ncu rc=$?"; ls -la $OUT/f2_fft_ncu.ncu-rep 2>/dev/null
