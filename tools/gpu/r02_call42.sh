#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
echo "== tests"; timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_golden_fixtures.py tests/test_gpu_round2.py tests/test_gpu_fit_engines.py -m gpu -x -q 2>&1 | tail -2
B="python bench.py --steps 3 --warmup 2 --no-e2e --no-cpu-baseline"
run() { name=$1; shift; timeout 300 $B "$@" > $O/r02ai_$name.json 2> $O/r02ai_$name.err; echo "$name rc=$?"; python tools/bench_brief.py $name < $O/r02ai_$name.json 2>/dev/null || tail -3 $O/r02ai_$name.err; }
run cfg3 --config cfg3
