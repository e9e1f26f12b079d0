#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
for c in 0 6 8 12 16; do
  LAMA_B200_FU_CHUNK=$c timeout 300 python bench.py --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/bench_chunk_$c.json 2> gpurun_out/bench_chunk_$c.err
  python -c "
import json; d=json.load(open('gpurun_out/bench_chunk_$c.json')); print('chunk', $c, 'img/s', round(d['value'],1), 'ms', round(d['ms_per_step'],2), 'launches', d['launches_per_step'], 'fu_ms', d['roofline']['fourier_unit']['ms'], d['clocks']['sm_mhz'])"
done
