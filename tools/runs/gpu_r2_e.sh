#!/bin/bash
set -u
mkdir -p gpurun_out
OUT=gpurun_out
export PYTHONUNBUFFERED=1
python -m lama_b200.build > $OUT/build_e.log 2>&1; echo "build rc=$?" | tee $OUT/summary_e.txt
timeout 300 python -m pytest tests/test_gpu_parity.py -q --timeout 120 -k "channel_group_planar or planar_chain" > $OUT/pytest_cg_e.log 2>&1; echo "pytest cg rc=$?" | tee -a $OUT/summary_e.txt
tail -5 $OUT/pytest_cg_e.log | tee -a $OUT/summary_e.txt
for layout in planar nhwc; do
  LAMA_B200_FU_LAYOUT=$layout timeout 300 python tools/tc_microbench.py > $OUT/tc_e_${layout}.txt 2>&1; echo "tc $layout rc=$?" | tee -a $OUT/summary_e.txt
  cat $OUT/tc_e_${layout}.txt | tee -a $OUT/summary_e.txt
done
timeout 200 python tools/fft_microbench.py --chain > $OUT/fu_chain_e.jsonl 2> $OUT/fu_chain_e.err; cat $OUT/fu_chain_e.jsonl | tee -a $OUT/summary_e.txt
timeout 900 python -m pytest tests -m gpu -q --timeout 300 -x > $OUT/pytest_e.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/summary_e.txt
tail -8 $OUT/pytest_e.log | tee -a $OUT/summary_e.txt
timeout 600 python bench.py --steps 10 > $OUT/bench_e.json 2> $OUT/bench_e.err; echo "bench rc=$?" | tee -a $OUT/summary_e.txt
python - <<PY | tee -a $OUT/summary_e.txt
import json
d = json.load(open("$OUT/bench_e.json"))
fu = d["roofline"]["fourier_unit"]
print(round(d["value"],1), "img/s", round(d["ms_per_step"],2), "ms; L", round(d["roofline"]["ms_per_launch"]*1e3,1), "us; FU cold", round(fu["ms"]*1e3,1), "us", fu["per_kernel"], fu["layout"])
print(json.dumps(d["torch_cuda_baseline"]))
PY
LAMA_B200_FU_LAYOUT=nhwc timeout 400 python bench.py --steps 10 --no-cpu-baseline --no-torch-cuda-baseline --no-fp32-arm --io f32 > $OUT/bench_e_nhwc.json 2> $OUT/bench_e_nhwc.err
python -c "import json; d=json.load(open('$OUT/bench_e_nhwc.json')); print('nhwc', round(d['value'],1), 'img/s', round(d['ms_per_step'],2))" | tee -a $OUT/summary_e.txt
