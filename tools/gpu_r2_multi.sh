#!/bin/bash
# usage: gpu_r2_multi.sh N   (run under `gpurun --gpus N`)
set -u
N=$1
mkdir -p gpurun_out
OUT=gpurun_out
export PYTHONUNBUFFERED=1
python -m lama_b200.build > $OUT/build_multi$N.log 2>&1; echo "build rc=$?" | tee $OUT/summary_multi$N.txt
nvidia-smi --query-gpu=index,name --format=csv,noheader | tee -a $OUT/summary_multi$N.txt
if [ "$N" = "2" ]; then
  timeout 600 python -m pytest tests/test_gpu_multi.py -q --timeout 500 > $OUT/pytest_multi.log 2>&1; echo "pytest multi rc=$?" | tee -a $OUT/summary_multi$N.txt
  grep -E "^(FAILED|ERROR)|passed|failed" $OUT/pytest_multi.log | tail -4 | tee -a $OUT/summary_multi$N.txt
fi
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tools/config4_bench.py --gpus $N --steps 3 > $OUT/config4_n$N.json 2> $OUT/config4_n$N.err; echo "config4 rc=$?" | tee -a $OUT/summary_multi$N.txt
cat $OUT/config4_n$N.json | tee -a $OUT/summary_multi$N.txt; tail -3 $OUT/config4_n$N.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 10 --warmup 3 > $OUT/bench_n$N.json 2> $OUT/bench_n$N.err; echo "bench rc=$?" | tee -a $OUT/summary_multi$N.txt
python -c "
import json
d=json.load(open('$OUT/bench_n$N.json')); print('bench N=$N', round(d['value'],1), 'img/s', round(d['ms_per_step'],2), 'ms e2e', round(d['e2e']['value'],1), d['clocks'])" | tee -a $OUT/summary_multi$N.txt
