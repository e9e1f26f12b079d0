#!/bin/bash
mkdir -p gpurun_out
timeout 150 python tools/grad_check.py > gpurun_out/grad_check.json 2> gpurun_out/grad_check.err; echo "rc=$?"; cat gpurun_out/grad_check.json; tail -3 gpurun_out/grad_check.err
