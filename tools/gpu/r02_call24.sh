#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
echo "== grid tests"; timeout 1200 python -m pytest tests/test_gpu_round2.py tests/test_gpu_parity.py tests/test_golden_fixtures.py tests/test_gpu_operators.py -m gpu -x -q 2>&1 | tail -3
B="python bench.py --steps 3 --warmup 2 --no-e2e --no-cpu-baseline"
run() { name=$1; shift; timeout 300 $B "$@" > $O/r02z_$name.json 2> $O/r02z_$name.err; echo "$name rc=$?"; python tools/bench_brief.py $name < $O/r02z_$name.json 2>/dev/null || tail -3 $O/r02z_$name.err; }
run cfg2
run cfg5 --config cfg5
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_grid_tile_search -c 1 -o $O/r02_grid_tile_search -f python bench.py --steps 1 --warmup 0 --no-e2e --no-cpu-baseline > $O/ncu_tile_search.log 2>&1; tail -1 $O/ncu_tile_search.log | cut -c1-200
