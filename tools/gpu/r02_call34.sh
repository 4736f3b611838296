#!/bin/bash
mkdir -p gpurun_out
timeout 900 python tools/gpu_soak.py 420 0 > gpurun_out/r02_gpu_soak.json 2> gpurun_out/r02_gpu_soak.err; echo "rc=$?"; cat gpurun_out/r02_gpu_soak.json; tail -5 gpurun_out/r02_gpu_soak.err
