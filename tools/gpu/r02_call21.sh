#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
echo "== suite"; timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
B="python bench.py --steps 3 --warmup 2 --no-e2e --no-cpu-baseline"
run() { name=$1; shift; timeout 300 $B "$@" > $O/r02w_$name.json 2> $O/r02w_$name.err; echo "$name rc=$?"; python tools/bench_brief.py $name < $O/r02w_$name.json 2>/dev/null || tail -3 $O/r02w_$name.err; }
run cfg2
run cfg3 --config cfg3
run cfg4 --config cfg4
run cfg5 --config cfg5
MODELARDB_CUDA_LIB=$PWD/modelardb_rs_b200/libmodelardb_cuda_b3.so run cfg2_b3
echo "== default line"; timeout 900 python bench.py > $O/r02w_default.json 2> $O/r02w_default.err; python tools/bench_brief.py default < $O/r02w_default.json
