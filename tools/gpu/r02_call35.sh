#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python tools/gpu_soak.py 780 1195 > gpurun_out/r02_gpu_soak2.json 2> gpurun_out/r02_gpu_soak2.err; echo "rc=$?"; cat gpurun_out/r02_gpu_soak2.json; tail -5 gpurun_out/r02_gpu_soak2.err
