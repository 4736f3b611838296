#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
echo "== tests"; timeout 900 python -m pytest tests/test_gpu_round2.py tests/test_gpu_parity.py tests/test_golden_fixtures.py -m gpu -x -q -k "block or macaque or config1 or full_size or golden" 2>&1 | tail -3
B="python bench.py --steps 3 --warmup 2 --no-e2e --no-cpu-baseline --config cfg3"
run() { name=$1; shift; timeout 300 $B "$@" > $O/r02l_$name.json 2> $O/r02l_$name.err; echo "$name rc=$?"; python tools/bench_brief.py $name < $O/r02l_$name.json 2>/dev/null || tail -3 $O/r02l_$name.err; }
for w in 2 4 8; do run w$w --option block_row_warps=$w; done
