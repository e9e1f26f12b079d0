#!/bin/bash
set -u
mkdir -p gpurun_out
OUT=gpurun_out
export PYTHONUNBUFFERED=1
python -m lama_b200.build > $OUT/build_n.log 2>&1; echo "build rc=$?" | tee $OUT/summary_n.txt
for st in 1 0; do
  FFCB_TC_STACK=$st timeout 600 python -m pytest tests/test_gpu_parity.py -q --timeout 300 -x -k "conv or generator_golden or resnet_block_golden or stem or head" > $OUT/pytest_n_stack$st.log 2>&1; echo "pytest stack=$st rc=$?" | tee -a $OUT/summary_n.txt
  grep -E "^(FAILED|ERROR)|passed|failed" $OUT/pytest_n_stack$st.log | tail -4 | tee -a $OUT/summary_n.txt
done
for st in 0 1 auto; do
  if [ "$st" = "auto" ]; then unset FFCB_TC_STACK; else export FFCB_TC_STACK=$st; fi
  timeout 300 python tools/tc_microbench.py > $OUT/tc_n_stack$st.txt 2>&1; echo "tc stack=$st rc=$?" | tee -a $OUT/summary_n.txt
  tail -7 $OUT/tc_n_stack$st.txt | cut -c1-110 | tee -a $OUT/summary_n.txt
  timeout 400 python bench.py --steps 10 --no-cpu-baseline --no-torch-cuda-baseline --no-fp32-arm --io f32 > $OUT/bench_n_stack$st.json 2> $OUT/bench_n_stack$st.err
  python -c "import json; d=json.load(open('$OUT/bench_n_stack$st.json')); print('stack $st', round(d['value'],1), 'img/s', round(d['ms_per_step'],2), d['clocks'], 'L us', round(d['roofline']['ms_per_launch']*1e3,1))" | tee -a $OUT/summary_n.txt
done
unset FFCB_TC_STACK
timeout 900 python -m pytest tests -m gpu -q --timeout 300 -x > $OUT/pytest_n.log 2>&1; echo "pytest full rc=$?" | tee -a $OUT/summary_n.txt
grep -E "^(FAILED|ERROR)|passed|failed" $OUT/pytest_n.log | tail -6 | tee -a $OUT/summary_n.txt
