#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_grid_tile_tma -s 1 -c 1 -o $O/r02_grid_tile_tma -f python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline > $O/ncu_tma.log 2>&1; tail -2 $O/ncu_tma.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_grid_tile -s 1 -c 1 -o $O/r02_grid_tile_plain -f python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --option grid_plain_stores=1 > $O/ncu_plain.log 2>&1; tail -2 $O/ncu_plain.log
