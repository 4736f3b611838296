#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
echo "== suite"; timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
B="python bench.py --steps 3 --warmup 2 --no-e2e --no-cpu-baseline"
run() { name=$1; shift; timeout 300 $B "$@" > $O/r02n_$name.json 2> $O/r02n_$name.err; echo "$name rc=$?"; python tools/bench_brief.py $name < $O/r02n_$name.json 2>/dev/null || tail -3 $O/r02n_$name.err; }
run cfg2
run cfg4 --config cfg4
run cfg3 --config cfg3
run cfg5 --config cfg5
run cfg2_lanes --fit-engine 4
run homog --sine 100:100,5:15,1000:1000
run homog_lanes --sine 100:100,5:15,1000:1000 --fit-engine 4
