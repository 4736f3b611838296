#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
echo "== fit engine tests"; timeout 900 python -m pytest tests/test_gpu_fit_engines.py tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -3
B="python bench.py --steps 3 --warmup 2 --no-e2e --no-cpu-baseline"
run() { name=$1; shift; timeout 300 $B "$@" > $O/r02r_$name.json 2> $O/r02r_$name.err; echo "$name rc=$?"; python tools/bench_brief.py $name < $O/r02r_$name.json 2>/dev/null || tail -3 $O/r02r_$name.err; }
run cfg2_auto
run cfg4_auto --config cfg4
run cfg5_screen --config cfg5 --fit-engine 5
