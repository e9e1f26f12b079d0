#!/bin/bash
# round 2, call B: bring-up of the channel-group planar FourierUnit chain
set -u
mkdir -p gpurun_out
OUT=gpurun_out
export PYTHONUNBUFFERED=1
python -c "import __graft_entry__ as g; g.build()" > $OUT/build.log 2>&1; echo "build rc=$?" | tee $OUT/summary.txt
timeout 300 python -m pytest tests/test_gpu_parity.py -q --timeout 120 -k "channel_group_planar" > $OUT/pytest_cg.log 2>&1; rc=$?; echo "pytest cg rc=$rc" | tee -a $OUT/summary.txt
tail -40 $OUT/pytest_cg.log | tee -a $OUT/summary.txt
if [ $rc -ne 0 ]; then
  FFCB_TC_DESC_SWAP=1 timeout 300 python -m pytest tests/test_gpu_parity.py -q --timeout 120 -k "conv_tc_channel_group_planar" > $OUT/pytest_cg_swap.log 2>&1; echo "pytest cg (LBO/SBO swapped) rc=$?" | tee -a $OUT/summary.txt
  tail -15 $OUT/pytest_cg_swap.log | tee -a $OUT/summary.txt
fi
timeout 200 python tools/fft_microbench.py --chain > $OUT/fu_chain.jsonl 2> $OUT/fu_chain.err; echo "fu chain rc=$?" | tee -a $OUT/summary.txt
cat $OUT/fu_chain.jsonl | tee -a $OUT/summary.txt; tail -5 $OUT/fu_chain.err | tee -a $OUT/summary.txt
python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1; echo "smoke rc=$?" | tee -a $OUT/summary.txt; tail -3 $OUT/smoke.log | tee -a $OUT/summary.txt
timeout 1200 python -m pytest tests -m gpu -q --timeout 300 --durations=10 -x > $OUT/pytest.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/summary.txt
tail -15 $OUT/pytest.log | tee -a $OUT/summary.txt
timeout 600 python bench.py > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?" | tee -a $OUT/summary.txt
cat $OUT/bench.json | tee -a $OUT/summary.txt; tail -5 $OUT/bench.err
