#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
echo "== fit engine tests"; timeout 900 python -m pytest tests/test_gpu_fit_engines.py tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -3
B="python bench.py --steps 3 --warmup 2 --no-e2e --no-cpu-baseline"
run() { name=$1; shift; timeout 300 $B "$@" > $O/r02u_$name.json 2> $O/r02u_$name.err; echo "$name rc=$?"; python tools/bench_brief.py $name < $O/r02u_$name.json 2>/dev/null || tail -3 $O/r02u_$name.err; }
run cfg2_auto
run cfg4_auto --config cfg4
run cfg5_screen --config cfg5 --fit-engine 5
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_spec_async -c 1 -o $O/r02_spec_screen3 -f python bench.py --steps 1 --warmup 0 --no-e2e --no-cpu-baseline > $O/ncu_screen3.log 2>&1; tail -1 $O/ncu_screen2.log | cut -c1-200
