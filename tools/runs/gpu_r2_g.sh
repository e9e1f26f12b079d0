#!/bin/bash
set -u
mkdir -p gpurun_out
OUT=gpurun_out
export PYTHONUNBUFFERED=1
python -m lama_b200.build > $OUT/build_g.log 2>&1; echo "build rc=$?" | tee $OUT/summary_g.txt
for c in nhwc_to_planar4 spatial_taps_plus_interleaved flat_interleaved_to_planar8 flat_ragged_m; do
  CUDA_LAUNCH_BLOCKING=1 timeout 120 python -m pytest tests/test_gpu_parity.py -q --timeout 100 -k "conv_tc_channel_group_planar_operands and $c" > $OUT/bisect_$c.log 2>&1
  echo "$c rc=$? $(grep -E 'passed|failed' $OUT/bisect_$c.log | tail -1)" | tee -a $OUT/summary_g.txt
  grep -E "^E  " $OUT/bisect_$c.log | head -3 | tee -a $OUT/summary_g.txt
done
timeout 300 python -m pytest tests/test_gpu_parity.py -q --timeout 120 -k "channel_group_planar or planar_chain" > $OUT/pytest_cg_g.log 2>&1; echo "pytest cg rc=$?" | tee -a $OUT/summary_g.txt
grep -E "^(FAILED|ERROR)|passed|failed" $OUT/pytest_cg_g.log | tail -8 | tee -a $OUT/summary_g.txt
for layout in planar nhwc; do
  LAMA_B200_FU_LAYOUT=$layout timeout 300 python tools/tc_microbench.py > $OUT/tc_g_${layout}.txt 2>&1; echo "tc $layout rc=$?" | tee -a $OUT/summary_g.txt
  tail -8 $OUT/tc_g_${layout}.txt | tee -a $OUT/summary_g.txt
done
timeout 200 python tools/fft_microbench.py --chain > $OUT/fu_chain_g.jsonl 2> $OUT/fu_chain_g.err; cat $OUT/fu_chain_g.jsonl | tee -a $OUT/summary_g.txt
timeout 900 python -m pytest tests -m gpu -q --timeout 300 -x > $OUT/pytest_g.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/summary_g.txt
grep -E "^(FAILED|ERROR)|passed|failed" $OUT/pytest_g.log | tail -8 | tee -a $OUT/summary_g.txt
timeout 600 python bench.py --steps 10 > $OUT/bench_g.json 2> $OUT/bench_g.err; echo "bench rc=$?" | tee -a $OUT/summary_g.txt
python - <<PY | tee -a $OUT/summary_g.txt
import json
d = json.load(open("$OUT/bench_g.json"))
fu = d["roofline"]["fourier_unit"]
print(round(d["value"],1), "img/s", round(d["ms_per_step"],2), "ms; L", round(d["roofline"]["ms_per_launch"]*1e3,1), "us; FU cold", round(fu["ms"]*1e3,1), "us", fu["per_kernel"], fu["layout"])
print(json.dumps(d["torch_cuda_baseline"]))
PY
LAMA_B200_FU_LAYOUT=nhwc timeout 400 python bench.py --steps 10 --no-cpu-baseline --no-torch-cuda-baseline --no-fp32-arm --io f32 > $OUT/bench_g_nhwc.json 2> $OUT/bench_g_nhwc.err
python -c "import json; d=json.load(open('$OUT/bench_g_nhwc.json')); print('nhwc', round(d['value'],1), 'img/s', round(d['ms_per_step'],2))" | tee -a $OUT/summary_g.txt
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"plane_cg_kernel" -c 2 -f -o $OUT/r02_plane_cg_v2 python tools/fft_microbench.py --chain-planar-once > $OUT/ncu_plane_g.log 2>&1; echo "ncu plane rc=$?" | tee -a $OUT/summary_g.txt
