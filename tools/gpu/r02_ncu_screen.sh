#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_spec_async -c 1 -o $O/r02_spec_screen -f python bench.py --steps 1 --warmup 0 --no-e2e --no-cpu-baseline > $O/ncu_screen.log 2>&1; tail -2 $O/ncu_screen.log | cut -c1-300
