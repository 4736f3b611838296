#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'k_swing_finish$' -c 1 -o $O/r02_swing_finish -f python bench.py --steps 1 --warmup 0 --no-e2e --no-cpu-baseline > $O/ncu_swing.log 2>&1; tail -2 $O/ncu_swing.log | cut -c1-300
