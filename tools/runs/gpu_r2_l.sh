#!/bin/bash
set -u
mkdir -p gpurun_out
OUT=gpurun_out
export PYTHONUNBUFFERED=1
python -m lama_b200.build > $OUT/build_l.log 2>&1; echo "build rc=$?" | tee $OUT/summary_l.txt
timeout 600 python -m pytest tests/test_gpu_parity.py -q --timeout 300 -x -k "fft or fourier or planar or spectral or generator" > $OUT/pytest_l.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/summary_l.txt
grep -E "^(FAILED|ERROR)|passed|failed" $OUT/pytest_l.log | tail -4 | tee -a $OUT/summary_l.txt
for pf in 1 0; do
  FFCB_FFT_PREFETCH=$pf timeout 200 python tools/fft_microbench.py --chain > $OUT/fu_chain_l_pf$pf.jsonl 2> $OUT/fu_chain_l_pf$pf.err; echo "prefetch=$pf" | tee -a $OUT/summary_l.txt; grep '"planar": true' $OUT/fu_chain_l_pf$pf.jsonl | tee -a $OUT/summary_l.txt
done
for pf in 1 0; do
  FFCB_FFT_PREFETCH=$pf timeout 400 python bench.py --steps 10 --no-cpu-baseline --no-torch-cuda-baseline --no-fp32-arm --io f32 > $OUT/bench_l_pf$pf.json 2> $OUT/bench_l_pf$pf.err
  python -c "import json; d=json.load(open('$OUT/bench_l_pf$pf.json')); print('prefetch $pf', round(d['value'],1), 'img/s', round(d['ms_per_step'],2), d['clocks'])" | tee -a $OUT/summary_l.txt
done
timeout 600 python tools/config4_bench.py --gpus 1 --steps 3 > $OUT/config4_n1.json 2> $OUT/config4_n1.err; echo "config4 n1 rc=$?" | tee -a $OUT/summary_l.txt
cat $OUT/config4_n1.json | tee -a $OUT/summary_l.txt
