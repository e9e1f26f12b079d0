#!/bin/bash
set -u
mkdir -p gpurun_out
OUT=gpurun_out
export PYTHONUNBUFFERED=1
python -m lama_b200.build > $OUT/build_m.log 2>&1; echo "build rc=$?" | tee $OUT/summary_m.txt
timeout 900 python -m pytest tests -m gpu -q --timeout 300 -x > $OUT/pytest_m.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/summary_m.txt
grep -E "^(FAILED|ERROR)|passed|failed" $OUT/pytest_m.log | tail -6 | tee -a $OUT/summary_m.txt
timeout 200 python tools/fft_microbench.py --chain > $OUT/fu_chain_m.jsonl 2> $OUT/fu_chain_m.err; grep '"planar": true' $OUT/fu_chain_m.jsonl | tee -a $OUT/summary_m.txt
timeout 900 python bench.py > $OUT/bench_m.json 2> $OUT/bench_m.err; echo "bench rc=$?" | tee -a $OUT/summary_m.txt
python - <<PY | tee -a $OUT/summary_m.txt
import json
d = json.load(open("$OUT/bench_m.json"))
fu = d["roofline"]["fourier_unit"]
print(round(d["value"],1), "img/s", round(d["ms_per_step"],2), "ms; e2e", round(d["e2e"]["value"],1), "; L", round(d["roofline"]["ms_per_launch"]*1e3,1), "us frac", round(d["roofline"]["frac"],3), "; FU cold", round(fu["ms"]*1e3,1), "us frac", round(fu["frac"],3), fu["traffic"])
print("torch-cuda", json.dumps(d["torch_cuda_baseline"]))
print("clocks", d["clocks"])
PY
LAMA_B200_POOL=0 timeout 400 python bench.py --steps 10 --no-cpu-baseline --no-torch-cuda-baseline --no-fp32-arm --io f32 > $OUT/bench_m_nopool.json 2> $OUT/bench_m_nopool.err
python -c "import json; d=json.load(open('$OUT/bench_m_nopool.json')); print('no pooling', round(d['value'],1), 'img/s', round(d['ms_per_step'],2), d['clocks'])" | tee -a $OUT/summary_m.txt
timeout 600 python tools/config4_bench.py --gpus 1 --steps 3 > $OUT/config4_n1.json 2> $OUT/config4_n1.err; echo "config4 n1 rc=$?" | tee -a $OUT/summary_m.txt
cat $OUT/config4_n1.json | tee -a $OUT/summary_m.txt; tail -3 $OUT/config4_n1.err | tee -a $OUT/summary_m.txt
nvidia-smi --query-gpu=memory.used,memory.total --format=csv | tee -a $OUT/summary_m.txt
