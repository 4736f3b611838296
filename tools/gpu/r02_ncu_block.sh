#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_macaque_block -s 2 -c 2 -o $O/r02_macaque_block -f python bench.py --config cfg3 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline > $O/ncu_block.log 2>&1; tail -2 $O/ncu_block.log
