#!/bin/bash
mkdir -p gpurun_out
python -m lama_b200.build > gpurun_out/build_y.log 2>&1
timeout 300 python tools/refine_bench.py --size 1024 > gpurun_out/refine_y.json 2> gpurun_out/refine_y.err; echo "refine rc=$?"
cat gpurun_out/refine_y.json
