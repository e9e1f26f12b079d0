#!/bin/bash
# round 2, call C: measurements + profiles of the validated tree
set -u
mkdir -p gpurun_out
OUT=gpurun_out
export PYTHONUNBUFFERED=1
python -c "import __graft_entry__ as g; g.build()" > $OUT/build.log 2>&1; echo "build rc=$?" | tee $OUT/summary_c.txt
timeout 300 python tools/fft_microbench.py --gemm-knobs > $OUT/gemm_knobs.jsonl 2> $OUT/gemm_knobs.err; echo "gemm knobs rc=$?" | tee -a $OUT/summary_c.txt
cat $OUT/gemm_knobs.jsonl | tee -a $OUT/summary_c.txt
timeout 600 python bench.py --steps 10 > $OUT/bench_planar.json 2> $OUT/bench_planar.err; echo "bench planar rc=$?" | tee -a $OUT/summary_c.txt
LAMA_B200_FU_LAYOUT=nhwc timeout 400 python bench.py --steps 10 --no-cpu-baseline --no-torch-cuda-baseline --no-fp32-arm --io f32 > $OUT/bench_nhwc.json 2> $OUT/bench_nhwc.err; echo "bench nhwc rc=$?" | tee -a $OUT/summary_c.txt
FFCB_L2_HINTS=0 timeout 400 python bench.py --steps 10 --no-cpu-baseline --no-torch-cuda-baseline --no-fp32-arm --io f32 > $OUT/bench_nohints.json 2> $OUT/bench_nohints.err; echo "bench nohints rc=$?" | tee -a $OUT/summary_c.txt
for ch in 8 16; do
LAMA_B200_FU_CHUNK=$ch timeout 400 python bench.py --steps 10 --no-cpu-baseline --no-torch-cuda-baseline --no-fp32-arm --io f32 > $OUT/bench_chunk$ch.json 2> $OUT/bench_chunk$ch.err; echo "bench chunk $ch rc=$?" | tee -a $OUT/summary_c.txt
done
for f in bench_planar bench_nhwc bench_nohints bench_chunk8 bench_chunk16; do python - <<PY | tee -a $OUT/summary_c.txt
import json
try:
    d = json.load(open("$OUT/$f.json"))
    fu = (d.get("roofline") or {}).get("fourier_unit") or {}
    print("$f", round(d["value"],1), "img/s", round(d["ms_per_step"],2), "ms", d["launches_per_step"], "launches; FU cold", fu.get("ms"), "warm", (fu.get("warm") or {}).get("ms"), fu.get("per_kernel"))
    if d.get("torch_cuda_baseline"): print("torch-cuda", json.dumps(d["torch_cuda_baseline"]))
    if d.get("fp32_arm"): print("fp32 arm", d["fp32_arm"])
except Exception as e: print("$f", "failed", e)
PY
done
timeout 300 python bench.py --size 256 --steps 10 --no-cpu-baseline --no-torch-cuda-baseline --no-fp32-arm --io f32 > $OUT/bench_256.json 2> $OUT/bench_256.err; echo "bench 256 rc=$?" | tee -a $OUT/summary_c.txt
timeout 400 python bench.py --size 1024 --steps 5 --no-cpu-baseline --no-torch-cuda-baseline --no-fp32-arm --io f32 > $OUT/bench_1024.json 2> $OUT/bench_1024.err; echo "bench 1024 rc=$?" | tee -a $OUT/summary_c.txt
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT/launches.csv python tools/profile_step.py bf16x3 32 > $OUT/prof1.log 2>&1; echo "launch list rc=$?" | tee -a $OUT/summary_c.txt
python tools/summarize_launches.py $OUT/launches.csv $OUT/call_order.txt > $OUT/launches.txt 2>&1; head -32 $OUT/launches.txt | tee -a $OUT/summary_c.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file $OUT/launches_bench_cmd.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-torch-cuda-baseline --no-fp32-arm --io f32 > $OUT/bench_under_ncu.json 2> $OUT/bench_under_ncu.err; echo "bench-cmd launch list rc=$?" | tee -a $OUT/summary_c.txt
python tools/summarize_launches.py $OUT/launches_bench_cmd.csv > $OUT/launches_bench_cmd.txt 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"plane_cg_kernel" -c 2 -f -o $OUT/r02_plane_cg python tools/fft_microbench.py --chain-planar-once > $OUT/ncu_plane.log 2>&1; echo "ncu plane rc=$?" | tee -a $OUT/summary_c.txt
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"conv_tc_kernel" -c 1 -f -o $OUT/r02_fu_gemm python tools/fft_microbench.py --chain-planar-once > $OUT/ncu_gemm.log 2>&1; echo "ncu fu gemm rc=$?" | tee -a $OUT/summary_c.txt
timeout 600 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:conv_tc_kernel -s 12 -c 5 -f -o $OUT/r02_conv_tc python tools/profile_step.py bf16x3 32 > $OUT/ncu_conv.log 2>&1; echo "ncu conv rc=$?" | tee -a $OUT/summary_c.txt
python tools/ncu_traffic.py $OUT/r02_traffic_raw.json $OUT/r02_plane_cg.ncu-rep $OUT/r02_fu_gemm.ncu-rep $OUT/r02_conv_tc.ncu-rep 2>&1 | tee -a $OUT/summary_c.txt
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 --print-limit 20 python tools/sanitize_step.py > $OUT/sanitizer.log 2>&1; echo "memcheck rc=$?" | tee -a $OUT/summary_c.txt
grep -E "ERROR SUMMARY" $OUT/sanitizer.log | head -3 | tee -a $OUT/summary_c.txt
